"""Quick device-resident timing of the vocabulary transform (development aid; bench.py is the contract)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from orb_slam2_with_comment_b200 import capi, synth, vocabulary

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
voc = synth.vocabulary_tree_full(10, 6, seed=7)
desc = synth.vocabulary_descriptors_fast(voc, F * 2000, seed=31)
kp_off = (np.arange(F + 1) * 2000).astype(np.int32)
v = vocabulary.ORBVocabulary().from_records(voc)
L = vocabulary._lib()
dev = torch.device("cuda:0")
n = F * 2000
d_off, d_desc = torch.from_numpy(kp_off).to(dev), torch.from_numpy(desc).to(dev)
i32, f64 = torch.int32, torch.float64
outs = [torch.empty(F + 1, dtype=i32, device=dev), torch.empty(n, dtype=i32, device=dev), torch.empty(n, dtype=f64, device=dev),
        torch.empty(F + 1, dtype=i32, device=dev), torch.empty(n, dtype=i32, device=dev), torch.empty(n + 1, dtype=i32, device=dev),
        torch.empty(n, dtype=i32, device=dev)]
torch.cuda.synchronize()
sp = C.c_void_p()
capi.check(L.orbgpu_vocabulary_stream(v._h, C.byref(sp)))
st = torch.cuda.ExternalStream(sp.value, device=dev)
step = lambda: capi.check(L.orbgpu_bow_transform_dev(v._h, F, d_off.data_ptr(), n, 2000, d_desc.data_ptr(), 4, *[C.c_void_p(t.data_ptr()) for t in outs], None, None))
for _ in range(2):
    step()
capi.check(L.orbgpu_vocabulary_sync(v._h))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    step()
e1.record(st)
capi.check(L.orbgpu_vocabulary_sync(v._h))
ms = e0.elapsed_time(e1) / steps
print(f"F={F} {ms:.3f} ms/batch {F / ms * 1e3:.0f} frames/s words/frame {int(outs[0][-1]) / F:.0f}")

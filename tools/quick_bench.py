"""Quick device-resident timing of the extraction path (development aid; bench.py is the contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from orb_slam2_with_comment_b200 import ORBextractor, synth

W, H, NF = 1241, 376, 2000
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ND = int(sys.argv[3]) if len(sys.argv) > 3 else 64   # distinct frames
base = np.stack([synth.g_rects(W, H, s) for s in range(min(ND, B))])
imgs = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B].copy()
ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
d_img = torch.from_numpy(imgs).cuda()
d_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8, device="cuda")
d_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8, device="cuda")
d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ex.stream())
for _ in range(3):
    ex.extract_batch_dev(d_img.data_ptr(), B, W, H, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
ex.sync()
with torch.cuda.stream(st):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        ex.extract_batch_dev(d_img.data_ptr(), B, W, H, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
    e1.record(st)
ex.sync()
ms = e0.elapsed_time(e1) / steps
print(f"B={B} {ms:.3f} ms/batch  {B / ms * 1e3:.0f} frames/s  kp/frame {d_cnt.float().mean().item():.0f} launches {ex.last_launches()}")
t = time.time(); kp, desc, cnt = ex.extract_batch(imgs); dt = time.time() - t
print(f"host path: {dt*1e3:.1f} ms/batch {B/dt:.0f} frames/s")

"""Per-stage device times of the extraction path (CUDA events on the launching stream, orbgpu_extractor_stage_ms): development aid.
    python tools/stage_times.py [batch] [reps] [distinct frames]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from orb_slam2_with_comment_b200 import ORBextractor, synth

W, H, NF = 1241, 376, 2000
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ND = int(sys.argv[3]) if len(sys.argv) > 3 else 64
import os
cache = os.environ.get("ORBGPU_FRAMES_CACHE")   # development: reuse the synthetic frames across runs of one session
if cache and os.path.exists(cache) and len(np.load(cache, mmap_mode="r")) >= min(ND, B):
    base = np.load(cache)[:min(ND, B)]
else:
    base = np.stack([synth.g_rects(W, H, s) for s in range(min(ND, B))])
    if cache:
        np.save(cache, base)
imgs = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B].copy()
ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
d_img = torch.from_numpy(imgs).cuda()
d_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8, device="cuda")
d_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8, device="cuda")
d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
ex.set_profiling(True)
acc = np.zeros(5)
for i in range(reps + 2):
    ex.extract_batch_dev(d_img.data_ptr(), B, W, H, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
    ex.sync()
    if i >= 2:
        acc += np.array([ex.stage_ms()[k] for k in ex.STAGES])
acc /= reps
names = ("pyramid", "fast_cells", "octree", "blur", "orient_desc")
print(f"B={B} " + "  ".join(f"{n} {v:.3f}" for n, v in zip(names, acc)) + f"  sum {acc.sum():.3f} ms")

"""End-to-end (host pointers in/out) extraction time for the current ORBGPU_CHUNK / ORBGPU_STREAMS setting: 1024 KITTI frames."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from orb_slam2_with_comment_b200 import ORBextractor, capi
B = 1024
frames = bench.make_frames(B)
ex = ORBextractor(bench.NFEATURES, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, max_width=bench.W, max_height=bench.H, max_batch=B)
h_img = torch.from_numpy(frames).pin_memory()
h_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8).pin_memory()
h_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8).pin_memory()
h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
step = lambda: capi.check(ex._lib.orbgpu_extract_batch(ex._h, h_img.data_ptr(), B, bench.W, bench.H, bench.W, bench.W * bench.H, h_kp.data_ptr(), h_desc.data_ptr(), ex.kp_cap, h_cnt.data_ptr()))
for _ in range(3):
    step()
t0 = time.perf_counter()
for _ in range(10):
    step()
dt = (time.perf_counter() - t0) / 10
print(f"chunk {os.environ.get('ORBGPU_CHUNK', 'default')} streams {os.environ.get('ORBGPU_STREAMS', 'default')}: {dt * 1e3:.3f} ms  {B / dt:.0f} frames/s", flush=True)

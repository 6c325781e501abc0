#!/usr/bin/env python3
"""End-to-end extraction through the single-process multi-GPU dispatcher (orbgpu_multi_extract_batch): pinned host images in,
host key points / descriptors out, 1024 KITTI frames per device per step.  One JSON line per device count.
    python tools/multi_bench.py [--gpus 1,2,4,8] [--steps 5]"""
import argparse, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from orb_slam2_with_comment_b200.extractor import MultiGpuExtractor

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--per-gpu", type=int, default=1024)
ap.add_argument("--chunk", type=int, default=1024, help="max_batch_per_device")
args = ap.parse_args()
ndev = torch.cuda.device_count()
counts = [int(x) for x in args.gpus.split(",") if x] or [n for n in (1, 2, 4, 8) if n <= ndev]
base = bench.make_frames(min(256, args.per_gpu))
for n in counts:
    B = n * args.per_gpu
    h_img = torch.empty((B, bench.H, bench.W), dtype=torch.uint8).pin_memory()
    hv = h_img.numpy()
    for i in range(0, B, len(base)):
        hv[i:i + len(base)] = base[:min(len(base), B - i)]
    me = MultiGpuExtractor(list(range(n)), bench.NFEATURES, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, max_width=bench.W,
                           max_height=bench.H, max_batch_per_device=args.chunk)
    from orb_slam2_with_comment_b200.capi import KP_DTYPE
    h_kp = torch.empty(B * me.kp_cap * 28, dtype=torch.uint8).pin_memory()
    h_desc = torch.empty(B * me.kp_cap * 32, dtype=torch.uint8).pin_memory()
    h_cnt = torch.empty(B, dtype=torch.int32).pin_memory()
    kp = h_kp.numpy().view(KP_DTYPE).reshape(B, me.kp_cap)
    desc = h_desc.numpy().reshape(B, me.kp_cap, 32)
    cnt = h_cnt.numpy()
    for _ in range(2):
        me.extract_batch(hv, kp, desc, cnt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        me.extract_batch(hv, kp, desc, cnt)
    dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"gpus": n, "frames_per_step": B, "ms_per_step": dt * 1e3, "frames_per_s": B / dt, "kp_per_frame": float(cnt.mean()),
                      "path": "orbgpu_multi_extract_batch (one process, one host thread per device, pinned host buffers)"}), flush=True)
    me.close()
    del h_img, h_kp, h_desc, h_cnt

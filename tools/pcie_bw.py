#!/usr/bin/env python3
"""Host<->device copy ceiling of this box for the bench's transfer sizes: pinned memory, one copy stream per direction and
device, H2D of the images and D2H of the results running at the same time on 1, 2, 4, ... N devices CONCURRENTLY (one host
thread per device, as the multi-GPU extraction dispatcher does).  The end-to-end extraction rate cannot exceed
frames_per_s_ceiling = aggregate H2D bytes/s / bytes per frame; bench.py reports e2e next to this ceiling.

    python tools/pcie_bw.py [--gpus 1,2,4,8] [--frames 1024] [--reps 8] [--write-combined]
Prints one JSON line per device count.
"""
import argparse
import json
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="")
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--write-combined", action="store_true", help="host input buffers allocated cudaHostAllocWriteCombined")
args = ap.parse_args()
ndev = torch.cuda.device_count()
counts = [int(x) for x in args.gpus.split(",") if x] or [n for n in (1, 2, 4, 8) if n <= ndev]
W, H, KP_CAP = 1241, 376, 2024
n_in, n_out = args.frames * W * H, args.frames * KP_CAP * 60


def host_buffer(nbytes, wc):
    if not wc:
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), 0x04 | 0x01) == 0   # write-combined | portable
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


bufs = []
for d in range(max(counts)):
    dev = torch.device("cuda", d)
    bufs.append(dict(dev=dev, h_in=host_buffer(n_in, args.write_combined), d_in=torch.empty(n_in, dtype=torch.uint8, device=dev),
                     h_out=torch.empty(n_out, dtype=torch.uint8).pin_memory(), d_out=torch.empty(n_out, dtype=torch.uint8, device=dev),
                     s1=torch.cuda.Stream(dev), s2=torch.cuda.Stream(dev)))


def worker(b, h2d, d2h, reps, barrier, out, i):
    torch.cuda.set_device(b["dev"])
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(b["s1"]):
                b["d_in"].copy_(b["h_in"], non_blocking=True)
        if d2h:
            with torch.cuda.stream(b["s2"]):
                b["h_out"].copy_(b["d_out"], non_blocking=True)
    b["s1"].synchronize()
    b["s2"].synchronize()
    out[i] = time.perf_counter() - t0


def run(n, h2d, d2h, reps):
    barrier, out = threading.Barrier(n), [0.0] * n
    th = [threading.Thread(target=worker, args=(bufs[i], h2d, d2h, reps, barrier, out, i)) for i in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return max(out) / reps


for n in counts:
    run(n, True, True, 2)
    a, b, c = run(n, True, False, args.reps), run(n, False, True, args.reps), run(n, True, True, args.reps)
    print(json.dumps({"gpus": n, "frames_per_gpu": args.frames, "write_combined": args.write_combined,
                      "h2d_only_GBps_aggregate": n * n_in / a / 1e9, "d2h_only_GBps_aggregate": n * n_out / b / 1e9,
                      "both_ms_per_step": c * 1e3, "both_h2d_GBps_aggregate": n * n_in / c / 1e9, "both_d2h_GBps_aggregate": n * n_out / c / 1e9,
                      "frames_per_s_ceiling": n * args.frames / c}), flush=True)

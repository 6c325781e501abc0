#!/usr/bin/env python3
"""BASELINE.json configs #4 and #5 at their stated sizes over the GPUs of one box, from ONE process (plain per-device streams,
one host thread per device, host gather; no collective — the path shards by frame / frame pair):

  #4  EuRoC-shaped 752x480 batch of 8192 synthetic frames, nFeatures=1200, frame-sharded across N GPUs
      (a) end to end through orbgpu_multi_extract_batch: pinned host images in, host key points / descriptors out
      (b) device-resident: every device extracts its share from HBM-resident images
  #5  brute-force 2000x2000 256-bit keyframe pairs (SearchByBoW(KF,KF) rule, one node), 65,536 pairs sharded across N GPUs,
      pairs formed from 256 distinct descriptor sets per device (BASELINE allows 4096 distinct sets; stated here)

    python tools/bench_configs_multi.py [--gpus 1,2,4,8]
One JSON line per (config, device count).  Results are checked: (a) is compared frame by frame with a one-device run of the same
frames (first 64 frames of every device's range), #5's match counts are equal on all devices (same pair list per device)."""
import argparse, json, sys, threading, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from orb_slam2_with_comment_b200 import ORBextractor, synth
from orb_slam2_with_comment_b200.capi import KP_DTYPE
from orb_slam2_with_comment_b200.extractor import MultiGpuExtractor
from orb_slam2_with_comment_b200.matcher import ORBmatcher

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="")
ap.add_argument("--frames", type=int, default=8192)
ap.add_argument("--pairs", type=int, default=65536)
args = ap.parse_args()
ndev = torch.cuda.device_count()
counts = [int(x) for x in args.gpus.split(",") if x] or [n for n in (1, 2, 4, 8) if n <= ndev]
W, H, NF = 752, 480, 1200


def out(d):
    print(json.dumps(d), flush=True)


from concurrent.futures import ProcessPoolExecutor


def _euroc(s):
    return synth.g_rects(W, H, s)


with ProcessPoolExecutor(max_workers=16) as pool:
    base = np.stack(list(pool.map(_euroc, range(256), chunksize=8)))
N = args.frames
h_img = torch.empty((N, H, W), dtype=torch.uint8).pin_memory()
hv = h_img.numpy()
for i in range(0, N, 256):
    hv[i:i + 256] = base[:min(256, N - i)]

# one-device reference of the first 64 frames of every possible range start
ref_ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=64)
ref_cache = {}


def ref_of(f0):
    if f0 not in ref_cache:
        kp, d, c = ref_ex.extract_batch(np.ascontiguousarray(hv[f0:f0 + 64]))
        ref_cache[f0] = (kp.copy(), d.copy(), c.copy())
    return ref_cache[f0]


for n in counts:
    # ---- #4 (a) end to end through the dispatcher
    me = MultiGpuExtractor(list(range(n)), NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch_per_device=1024)
    cap = me.kp_cap
    h_kp = torch.empty(N * cap * 28, dtype=torch.uint8).pin_memory()
    h_desc = torch.empty(N * cap * 32, dtype=torch.uint8).pin_memory()
    h_cnt = torch.empty(N, dtype=torch.int32).pin_memory()
    kp, desc, cnt = h_kp.numpy().view(KP_DTYPE).reshape(N, cap), h_desc.numpy().reshape(N, cap, 32), h_cnt.numpy()
    me.extract_batch(hv, kp, desc, cnt)
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        me.extract_batch(hv, kp, desc, cnt)
    dt = (time.perf_counter() - t0) / reps
    checked = 0
    for g in range(n):
        f0, f1 = me.frame_range(N, g)
        rk, rd, rc = ref_of(f0)
        m = min(64, f1 - f0)
        assert np.array_equal(cnt[f0:f0 + m], rc[:m]), (n, g)
        for f in range(m):
            assert kp[f0 + f, :rc[f]].tobytes() == rk[f, :rc[f]].tobytes() and np.array_equal(desc[f0 + f, :rc[f]], rd[f, :rc[f]]), (n, g, f)
        checked += m
    out({"config": 4, "gpus": n, "path": "end to end (orbgpu_multi_extract_batch, pinned host buffers)", "frames": N, "ms": dt * 1e3,
         "frames_per_s": N / dt, "keypoints_per_frame": float(cnt.mean()), "frames_checked_against_one_device_run": checked})
    me.close()
    del h_kp, h_desc, h_cnt

    # ---- #4 (b) device resident, one host thread per device
    per = N // n
    B = 512
    res = [None] * n

    def work(g):
        dev = torch.device("cuda", g)
        torch.cuda.set_device(dev)
        ex = ORBextractor(NF, 1.2, 8, 20, 7, device=g, max_width=W, max_height=H, max_batch=B)
        d_img = torch.from_numpy(hv[:B]).to(dev)
        d_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8, device=dev)
        d_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
        fn = lambda: ex.extract_batch_dev(d_img.data_ptr(), B, W, H, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
        for _ in range(2):
            fn()
        ex.sync()
        bar.wait()
        t0 = time.perf_counter()
        for _ in range(per // B):
            fn()
        ex.sync()
        res[g] = time.perf_counter() - t0
        ex.close()

    bar = threading.Barrier(n)
    th = [threading.Thread(target=work, args=(g,)) for g in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    out({"config": 4, "gpus": n, "path": "device resident (images in HBM), one host thread per device", "frames": per // B * B * n, "ms": max(res) * 1e3,
         "frames_per_s": per // B * B * n / max(res)})

    # ---- #5: 65,536 brute-force pairs over n devices
    P_total = args.pairs
    per_p = P_total // n
    CH = 4096
    mres, mcount = [None] * n, [None] * n

    def mwork(g):
        dev = torch.device("cuda", g)
        torch.cuda.set_device(dev)
        sA, sB, i1, i2, off, total = bench.make_match_workload(CH)
        m = ORBmatcher(0.75, True, device=g)
        hA, hB = m.upload(sA), m.upload(sB)
        d12 = torch.empty(total, dtype=torch.int32, device=dev)
        dd = torch.empty(total, dtype=torch.int32, device=dev)
        dn = torch.empty(CH, dtype=torch.int32, device=dev)
        fn = lambda: m.search_by_bow_dev(hA, hB, i1, i2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
        fn(); m.sync()
        bar2.wait()
        t0 = time.perf_counter()
        for _ in range(per_p // CH):
            fn()
        m.sync()
        mres[g] = time.perf_counter() - t0
        mcount[g] = int(dn.sum().item()) * (per_p // CH)
        m.release(hA); m.release(hB); m.close()

    bar2 = threading.Barrier(n)
    th = [threading.Thread(target=mwork, args=(g,)) for g in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert len(set(mcount)) == 1, mcount
    pairs_done = per_p // CH * CH * n
    sec = max(mres)
    peak, _ = bench.popc_peak()
    out({"config": 5, "gpus": n, "pairs": pairs_done, "distinct_descriptor_sets_per_device": min(bench.MATCH_SETS, CH), "ms": sec * 1e3,
         "pairs_per_s": pairs_done / sec, "distance_evals_per_s": pairs_done * 4e6 / sec, "matches_per_s": sum(mcount) / sec,
         "frac_of_popc_peak_8_per_distance": pairs_done * 4e6 * 8 / sec / (peak * n)})
ref_ex.close()

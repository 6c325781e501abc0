"""Quick device-resident timing of the brute-force (config #5) Hamming scan: tools/quick_match_bench.py [pairs] [sets]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orb_slam2_with_comment_b200 import synth
from orb_slam2_with_comment_b200.matcher import ORBmatcher, FrameSet, match_offsets

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_sets = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = 2000
A, B, angA, angB = synth.bruteforce_sets(n_sets, n, 900)
kp_off = np.arange(n_sets + 1, dtype=np.int32) * n
kA = np.zeros(n_sets * n, synth.KP_DTYPE); kA["angle"] = angA.ravel()
kB = np.zeros(n_sets * n, synth.KP_DTYPE); kB["angle"] = angB.ravel()
fl = np.ones(n_sets * n, np.uint8)
sA = FrameSet.single_node(kp_off, kA, A.reshape(-1, 32), kp_flags=fl)
sB = FrameSet.single_node(kp_off, kB, B.reshape(-1, 32), kp_flags=fl)
i1 = (np.arange(pairs) % n_sets).astype(np.int32)
i2 = i1.copy()   # B_i is the perturbed permutation of A_i
off, total = match_offsets(sA, i1)
dev = torch.device("cuda", 0)
d12 = torch.empty(total, dtype=torch.int32, device=dev)
dd = torch.empty(total, dtype=torch.int32, device=dev)
dn = torch.empty(pairs, dtype=torch.int32, device=dev)
for cfg in (dict(queries_per_thread=4), dict(queries_per_thread=8), dict(min_queries=0)):
    m = ORBmatcher(0.75, True)
    m.configure(**cfg)
    hA, hB = m.upload(sA), m.upload(sB)
    for _ in range(2):
        m.search_by_bow_dev(hA, hB, i1, i2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
    m.sync()
    t0 = time.time()
    reps = 3
    ms_tot = 0
    for _ in range(reps):
        m.search_by_bow_dev(hA, hB, i1, i2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
        ms, ev = m.last_stats()
        ms_tot += ms
    wall = (time.time() - t0) / reps
    ms = ms_tot / reps
    print(f"cfg {cfg}: {ms:.2f} ms kernels ({wall*1e3:.2f} ms wall) for {pairs} pairs, {ev/1e9:.2f} G evals -> {ev/ms/1e6:.1f} G evals/s, "
          f"{8*ev/ms/1e9:.2f} T popc/s, matches {int(dn.sum())} -> {int(dn.sum())/ms/1e3:.2f} M matches/s, launches {m.last_launches()}")
    m.release(hA); m.release(hB); m.close()

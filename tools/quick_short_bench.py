"""One pass of the short-list matcher kernels at the sizes of BASELINE configs #2 / #3 (development aid for ncu captures):
SearchByProjection (256 TUM frames x 5000 map points), SearchForTriangulation (1024 pairs of ~2000 key points), isInFrustum,
ComputeStereoMatches (256 KITTI pairs).   python tools/quick_short_bench.py [reps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import match_cases as mc
from orb_slam2_with_comment_b200 import ORBextractor, synth
from orb_slam2_with_comment_b200.matcher import ORBmatcher, match_offsets

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)


def timed(name, stream, fn, sync, work, unit):
    st = torch.cuda.ExternalStream(stream, device=dev)
    fn(); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    sync()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.3f} ms per call, {work / ms * 1e3:.0f} {unit}/s", flush=True)


m = ORBmatcher(0.8, True)
# SearchByProjection: 256 frames x 5000 map points
fs, mps, sf, th = mc.sbp_case(7, n_frames=256, n_lo=950, n_hi=1050, n_mp=5000, th=1.0)
hf, hm = m.upload(fs), m.upload_mappoints(mps, fs.n_frames)
d_o = [torch.zeros(n, dtype=torch.int32, device=dev) for n in (int(fs.kp_off[-1]), mps.n, mps.n, mps.n, fs.n_frames)]
timed("SearchByProjection 256 frames x 5000 map points", m.stream(), lambda: m.search_by_projection_dev(hf, hm, sf, 1.0, *[t.data_ptr() for t in d_o]), m.sync, 256, "frames")
_, ev = m.last_stats(); print("  distance evaluations per call:", ev)
# SearchForTriangulation: 1024 pairs
s1, s2, i1, i2, F12, epi, sf2, s2t = mc.tri_case(8, n_frames=129, n_lo=1900, n_hi=2100)
i1 = np.tile(i1, 8); i2 = np.tile(i2, 8); F12 = np.tile(F12, (8, 1)); epi = np.tile(epi, (8, 1))
h = m.upload(s1)
off, total = match_offsets(s1, i1)
d12, dd, dn = (torch.zeros(total, dtype=torch.int32, device=dev), torch.zeros(total, dtype=torch.int32, device=dev), torch.zeros(len(i1), dtype=torch.int32, device=dev))
timed("SearchForTriangulation 1024 pairs", m.stream(), lambda: m.search_for_triangulation_dev(h, h, i1, i2, F12, epi, sf2, s2t, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr()), m.sync, len(i1), "pairs")
_, ev = m.last_stats(); print("  distance evaluations per call:", ev)
# isInFrustum
args = mc.frustum_case(3, n_frames=256, n_mp=6000)
t0 = time.perf_counter(); m.isInFrustum(*args); print(f"isInFrustum (host call, {int(args[4][-1])} points): {(time.perf_counter() - t0) * 1e3:.2f} ms")
m.close()
# ComputeStereoMatches: 256 KITTI pairs
W, H, P = 1241, 376, 256
exL = ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=P)
exR = ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=P)
pairs = [synth.stereo_pair(W, H, s) for s in range(16)]
left = np.ascontiguousarray(np.stack([pairs[i % 16][0] for i in range(P)])); right = np.ascontiguousarray(np.stack([pairs[i % 16][1] for i in range(P)]))
for ex, im in ((exL, left), (exR, right)):
    d_img = torch.from_numpy(im).to(dev)
    ex._keep = [d_img, torch.zeros(P * ex.kp_cap * 28, dtype=torch.uint8, device=dev), torch.zeros(P * ex.kp_cap * 32, dtype=torch.uint8, device=dev), torch.zeros(P, dtype=torch.int32, device=dev)]
    ex.extract_batch_dev(d_img.data_ptr(), P, W, H, ex._keep[1].data_ptr(), ex._keep[2].data_ptr(), ex._keep[3].data_ptr()); ex.sync()
d_u = torch.zeros(P * exL.kp_cap, dtype=torch.float32, device=dev); d_d = torch.zeros(P * exL.kp_cap, dtype=torch.float32, device=dev)
timed("ComputeStereoMatches 256 KITTI pairs", exL.stream(), lambda: exL.stereo_matches_dev(exR, 0.537, 386.1448, d_u.data_ptr(), d_d.data_ptr()), exL.sync, P, "pairs")

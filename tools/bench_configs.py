#!/usr/bin/env python3
"""All five BASELINE.json configs on ONE GPU (development aid; bench.py is the contract).  Prints one JSON line per config.

  #1 single KITTI frame latency                     #2 TUM extraction + SearchByProjection vs 5k-point local maps
  #3 KITTI stereo extraction + SearchForTriangulation between consecutive keyframes
  #4 EuRoC batch extraction                          #5 brute-force 2000x2000 keyframe pairs (same as bench.py's matching leg)
Sizes are scaled by --scale (1.0 = the BASELINE sizes: 256 TUM frames, 1024 stereo pairs, 8192 EuRoC frames, 8192 pairs per GPU).
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orb_slam2_with_comment_b200 import ORBextractor, synth
from orb_slam2_with_comment_b200.matcher import ORBmatcher, FrameSet, MapPointSet, match_offsets

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)


def ev_time(stream_ptr, fn, reps, sync):
    st = torch.cuda.ExternalStream(stream_ptr, device=dev)
    for _ in range(2):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    sync()
    return e0.elapsed_time(e1) / reps


def frames(w, h, n, distinct=64, drift=0):
    """n frames from `distinct` scenes; with drift > 0 consecutive frames of a scene are the same scene shifted by `drift` px
    (so the previous frame's key points reappear), 8 frames per scene."""
    if not drift:
        base = [synth.g_rects(w, h, s) for s in range(min(distinct, n))]
        return np.ascontiguousarray(np.stack([base[i % len(base)] for i in range(n)]))
    wide = [synth.g_rects(w + 8 * drift, h, s) for s in range(min(distinct, (n + 7) // 8))]
    return np.ascontiguousarray(np.stack([wide[(i // 8) % len(wide)][:, drift * (i % 8):drift * (i % 8) + w] for i in range(n)]))


sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as ol   # parity assertions of the configs (the checker, never the thing measured)
PORT = ol.load_port()
SLAMREF = ol.load_slam_ref()


def extract_dev(ex, imgs, w, h):
    B = len(imgs)
    d_img = torch.from_numpy(imgs).to(dev)
    d_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8, device=dev)
    d_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    fn = lambda: ex.extract_batch_dev(d_img.data_ptr(), B, w, h, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
    return fn, d_kp, d_desc, d_cnt


def out(line):
    print(json.dumps(line), flush=True)


# ---- #1 -------------------------------------------------------------------------------------------------------------
W, H = 1241, 376
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=1)
img = frames(W, H, 1)
fn, *_ = extract_dev(ex, img, W, H)
ms_dev = ev_time(ex.stream(), fn, 50, ex.sync)
t0 = time.time()
for _ in range(50):
    kp, desc = ex(img[0])
ms_host = (time.time() - t0) / 50 * 1e3
out({"config": 1, "what": "single 1241x376 frame, 2000 features", "ms_per_frame_device": ms_dev, "ms_per_frame_host_call": ms_host, "keypoints": len(kp)})
ex.close()

# ---- #2 -------------------------------------------------------------------------------------------------------------
W, H, NF = 640, 480, 1000
B = max(8, int(256 * args.scale))
ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
imgs = frames(W, H, B, drift=3)
fn, d_kp, d_desc, d_cnt = extract_dev(ex, imgs, W, H)
ms_ex = ev_time(ex.stream(), fn, args.reps, ex.sync)
kp, desc, cnt = ex.extract_batch(imgs)
kp_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
keys = np.concatenate([kp[f, :cnt[f]] for f in range(B)])
descs = np.concatenate([desc[f, :cnt[f]] for f in range(B)])
parts, mp_off = [], [0]
for f in range(B):
    p = f - 1 if f else 0   # the local map of frame f is built from the PREVIOUS frame's real extraction
    parts.append(synth.local_map(kp[p, :cnt[p]], desc[p, :cnt[p]], 5000, W, H, 7000 + f))
    mp_off.append(mp_off[-1] + 5000)
cat = {k: np.concatenate([q[k] for q in parts]) for k in parts[0]}
fs = FrameSet(kp_off, keys, descs, grid=np.tile(synth.frame_grid(W, H), (B, 1)))
mps = MapPointSet(np.array(mp_off, np.int32), cat["proj_x"], cat["proj_y"], cat["view_cos"], cat["level"], cat["flags"], cat["desc"])
sf, _ = synth.scale_tables()
m = ORBmatcher(0.8, True)
hf, hm = m.upload(fs), m.upload_mappoints(mps, B)
d_o = [torch.zeros(n, dtype=torch.int32, device=dev) for n in (len(keys), mps.n, mps.n, mps.n, B)]
fn = lambda: m.search_by_projection_dev(hf, hm, sf, 1.0, *[t.data_ptr() for t in d_o])
ms_m = ev_time(m.stream(), fn, args.reps, m.sync)
_, evals = m.last_stats()
r = m.SearchByProjection(fs, mps, sf, 1.0)
t0 = time.time(); r = m.SearchByProjection(fs, mps, sf, 1.0); ms_host = (time.time() - t0) * 1e3
# parity of this very workload: every frame against the CPU port, the first frames against the reference's own ORBmatcher.cc
exp = ol.MatcherOracle(PORT, 0.8, True).SearchByProjection(fs, mps, sf, 1.0)
for k in ("nmatches", "kp_match", "mp_best_idx", "mp_best_dist", "mp_second_dist"):
    assert np.array_equal(r[k], exp[k]), f"config 2: {k} differs from the port"
nref = 0
if SLAMREF is not None:
    nref = min(B, 8)
    fs8 = FrameSet(kp_off[:nref + 1], keys[:kp_off[nref]], descs[:kp_off[nref]], grid=np.tile(synth.frame_grid(W, H), (nref, 1)))
    sl = slice(0, mp_off[nref])
    mps8 = MapPointSet(np.array(mp_off[:nref + 1], np.int32), cat["proj_x"][sl], cat["proj_y"][sl], cat["view_cos"][sl], cat["level"][sl],
                       cat["flags"][sl], cat["desc"][sl])
    e8 = ol.MatcherRef(SLAMREF, 0.8, True).SearchByProjection(fs8, mps8, sf, 1.0)
    assert np.array_equal(r["nmatches"][:nref], e8["nmatches"]) and np.array_equal(r["kp_match"][:kp_off[nref]], e8["kp_match"]), "config 2 differs from the reference"
out({"config": 2, "what": f"{B} TUM 640x480 frames, 1000 features; SearchByProjection vs 5000 map points per frame built from the previous frame's extraction (th=1, nnratio 0.8)",
     "parity": f"{B} frames bit-exact vs the CPU port, {nref} frames vs the reference's ORBmatcher.cc (libslamref.so)",
     "extract_frames_per_s": B / ms_ex * 1e3, "match_frames_per_s": B / ms_m * 1e3, "match_ms": ms_m, "match_host_call_ms": ms_host,
     "matches_per_frame": float(r["nmatches"].mean()), "distance_evals": int(evals), "matches_per_s": float(r["nmatches"].sum()) / ms_m * 1e3})
m.release(hf); m.release_mappoints(hm); m.close(); ex.close()

# ---- #3 -------------------------------------------------------------------------------------------------------------
W, H, NF = 1241, 376, 2000
P = max(8, int(1024 * args.scale))
ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=2 * (P + 1))
# consecutive key frames see the same scene 8 px further along (16 frames per scene), so triangulation finds real matches
wide = [synth.g_rects(W + 128, H, s) for s in range(8)]
left = np.ascontiguousarray(np.stack([wide[(f // 16) % 8][:, 8 * (f % 16):8 * (f % 16) + W] for f in range(P + 1)]))
right = np.ascontiguousarray(np.roll(left, -12, axis=2))      # constant-disparity right images
both = np.ascontiguousarray(np.concatenate([left, right]))
fn, *_ = extract_dev(ex, both, W, H)
ms_ex = ev_time(ex.stream(), fn, max(2, args.reps // 2), ex.sync)
kp, desc, cnt = ex.extract_batch(both)
NB = 2 * (P + 1)
kp_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
keys = np.concatenate([kp[f, :cnt[f]] for f in range(NB)])
descs = np.concatenate([desc[f, :cnt[f]] for f in range(NB)])
fv = synth.pack_feature_vectors(kp_off, descs, synth.synth_vocabulary())
fs = FrameSet(kp_off, keys, descs, fv_node_off=fv[0], fv_node_id=fv[1], fv_feat_off=fv[2], fv_feat=fv[3])
K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]])   # Examples/Stereo/KITTI00-02.yaml
F, ep = synth.fundamental_and_epipole(K, np.eye(3), np.array([0.5, 0.0, 0.001]))   # (almost) pure x translation: horizontal epipolar lines
i1, i2 = np.arange(0, P, dtype=np.int32), np.arange(P + 1, 2 * P + 1, dtype=np.int32)   # left image p against its right image
F12, EP = np.tile(F, (P, 1)), np.tile(ep, (P, 1))
sf, s2 = synth.scale_tables()
m = ORBmatcher(0.6, False)
h = m.upload(fs)
off, total = match_offsets(fs, i1)
d12, dd, dn = (torch.zeros(total, dtype=torch.int32, device=dev), torch.zeros(total, dtype=torch.int32, device=dev),
               torch.zeros(P, dtype=torch.int32, device=dev))
fn = lambda: m.search_for_triangulation_dev(h, h, i1, i2, F12, EP, sf, s2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
ms_m = ev_time(m.stream(), fn, args.reps, m.sync)
_, evals = m.last_stats()
# parity: the first pairs against the CPU port and against the reference's own SearchForTriangulation
ncheck = min(P, 12)
got = m.SearchForTriangulation(fs, fs, i1[:ncheck], i2[:ncheck], F12[:ncheck], EP[:ncheck], sf, s2)
exp = ol.MatcherOracle(PORT, 0.6, False).SearchForTriangulation(fs, fs, i1[:ncheck], i2[:ncheck], F12[:ncheck], EP[:ncheck], sf, s2)
for k in ("nmatches", "match12", "match_dist"):
    assert np.array_equal(got[k], exp[k]), f"config 3: {k} differs from the port"
if SLAMREF is not None:
    e3 = ol.MatcherRef(SLAMREF, 0.6, False).SearchForTriangulation(fs, fs, i1[:ncheck], i2[:ncheck], F12[:ncheck], EP[:ncheck], sf, s2)
    assert np.array_equal(got["nmatches"], e3["nmatches"]) and np.array_equal(got["match12"], e3["match12"]), "config 3 differs from the reference"
out({"config": 3, "parity": f"{ncheck} pairs bit-exact vs the CPU port" + (" and the reference's ORBmatcher.cc" if SLAMREF is not None else ""),
     "what": f"{P} KITTI stereo pairs: extraction of {2 * (P + 1)} images, SearchForTriangulation left vs right image of every pair "
     "(synthetic 10x10 vocabulary, no MapPoints, mono, checkOri off)", "extract_images_per_s": 2 * (P + 1) / ms_ex * 1e3, "pairs_per_s": P / ms_m * 1e3,
     "match_ms": ms_m, "distance_evals": int(evals), "matches_per_pair": float(dn.float().mean().item())})
m.release(h); m.close(); ex.close()
# stereo matching of the same pairs (Frame::ComputeStereoMatches), two extractor instances as in the stereo Frame constructor
exL = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=P)
exR = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=P)
fnL, *_ = extract_dev(exL, left[:P], W, H)
fnR, *_ = extract_dev(exR, right[:P], W, H)
fnL(); fnR(); exL.sync(); exR.sync()
d_u = torch.zeros(P * exL.kp_cap, dtype=torch.float32, device=dev)
d_d = torch.zeros(P * exL.kp_cap, dtype=torch.float32, device=dev)
fn = lambda: exL.stereo_matches_dev(exR, 0.537, 386.1448, d_u.data_ptr(), d_d.data_ptr())
ms_s = ev_time(exL.stream(), fn, args.reps, exL.sync)
out({"config": "3-stereo", "what": f"Frame::ComputeStereoMatches for {P} KITTI stereo pairs (device-resident key points, descriptors, pyramids)",
     "pairs_per_s": P / ms_s * 1e3, "ms": ms_s, "stereo_matches_per_pair": float((d_u >= 0).sum().item()) / P})
# the whole of config #3 chained on the device: extraction of both images of every key frame, ComputeStereoMatches, FeatureVectors by the
# device vocabulary transform (10 x 10 two-level tree, levelsup 0), frame set built in HBM, SearchForTriangulation between consecutive
# key frames with the stereo coordinates — the host sees only the per-frame counts
from orb_slam2_with_comment_b200.vocabulary import ORBVocabulary
voc = ORBVocabulary().from_records(synth.vocabulary_tree(k=10, L=2, seed=12345))
fnL, *_ = extract_dev(exL, left[:P], W, H)
fnR, *_ = extract_dev(exR, right[:P], W, H)
m = ORBmatcher(0.6, False)
i1c, i2c = np.arange(1, P, dtype=np.int32), np.arange(0, P - 1, dtype=np.int32)
F12c, EPc = np.tile(F, (P - 1, 1)), np.tile(ep, (P - 1, 1))
cap = exL.kp_cap
d12c = torch.zeros((P - 1) * cap, dtype=torch.int32, device=dev)
ddc = torch.zeros((P - 1) * cap, dtype=torch.int32, device=dev)
dnc = torch.zeros(P - 1, dtype=torch.int32, device=dev)


def chain():
    fnL(); fnR()
    exR.sync()   # the stereo kernels run on the left extractor's stream
    exL.stereo_matches_dev(exR, 0.537, 386.1448, d_u.data_ptr(), d_d.data_ptr())
    hc = m.frame_set_from_extraction(exL, voc, levelsup=0, kp_flag=0, u_right_ptr=d_u.data_ptr(), u_right_stride=cap)
    kp_off_c, _, _ = m.frame_set_info(hc)
    offc = kp_off_c[i1c].astype(np.int64) - kp_off_c[1]     # match slots of pair p start where frame i1[p]'s key points start
    m.search_for_triangulation_dev(hc, hc, i1c, i2c, F12c, EPc, sf, s2, offc, d12c.data_ptr(), ddc.data_ptr(), dnc.data_ptr())
    m.sync()
    m.release(hc)


chain(); chain()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(args.reps):
    chain()
torch.cuda.synchronize()
ms_c = (time.time() - t0) / args.reps * 1e3
out({"config": "3-chained", "what": f"{P} KITTI stereo key frames per step, everything on the device: extraction of {2 * P} images, ComputeStereoMatches, "
     "vocabulary transform, frame set in HBM, SearchForTriangulation (stereo) between consecutive key frames; wall clock incl. the host's "
     "count read-backs", "ms_per_step": ms_c, "keyframes_per_s": P / ms_c * 1e3, "images_per_s": 2 * P / ms_c * 1e3,
     "matches_per_pair": float(dnc.float().mean().item())})
m.close(); voc.close()
exL.close(); exR.close()

# ---- #4 -------------------------------------------------------------------------------------------------------------
W, H, NF = 752, 480, 1200
N = max(512, int(8192 * args.scale))
B = 512
ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
imgs = frames(W, H, B)
fn, d_kp, d_desc, d_cnt = extract_dev(ex, imgs, W, H)
ms = ev_time(ex.stream(), fn, max(1, N // B), ex.sync)
h_img = torch.from_numpy(imgs).pin_memory()
h_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8).pin_memory()
h_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8).pin_memory()
h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
from orb_slam2_with_comment_b200 import capi
step = lambda: capi.check(ex._lib.orbgpu_extract_batch(ex._h, h_img.data_ptr(), B, W, H, W, W * H, h_kp.data_ptr(), h_desc.data_ptr(), ex.kp_cap, h_cnt.data_ptr()))
step()
t0 = time.time()
for _ in range(max(1, N // B)):
    step()
e2e = (time.time() - t0) / max(1, N // B) * 1e3
out({"config": 4, "what": f"EuRoC 752x480, 1200 features: {N} frames in batches of {B} on one GPU", "frames_per_s_device": B / ms * 1e3,
     "frames_per_s_host_call": B / e2e * 1e3, "keypoints_per_frame": float(d_cnt.float().mean().item())})
ex.close()

"""Parity of an ORBmatcher implementation — the CPU port (tests/test_ref_matcher.py) or the CUDA path through the C ABI
(tests/test_gpu_ref_matcher.py) — against the REFERENCE ITSELF: oracle/_ref/libslamref.so holds the reference's own ORBmatcher.cc,
Frame.cc, KeyFrame.cc and MapPoint.cc compiled from /root/reference and runs its member functions on real Frame / KeyFrame /
MapPoint objects built from the same flat views (oracle/slam_ref.cc).

`impl(nnratio, checkOri)` returns an object with the method names of orb_slam2_with_comment_b200.matcher.ORBmatcher;
`ref(nnratio, checkOri)` returns an oracle_lib.MatcherRef.  Everything the reference function makes observable is compared
bit for bit (indices, counts); its internal best / second distances are not observable and stay pinned through the port.
"""
import numpy as np

import match_cases as mc
from orb_slam2_with_comment_b200 import synth
from orb_slam2_with_comment_b200.matcher import FrameSet, MapPointSet, WindowQuerySet

MBF = 37.5


def same(a, b, keys, what):
    for k in keys:
        assert np.array_equal(a[k], b[k]), f"{what}: {k} differs at {np.nonzero(a[k] != b[k])[0][:8]}"


def descriptor_distance(impl, ref):
    rs = np.random.RandomState(0)
    a = rs.randint(0, 256, (3000, 32)).astype(np.uint8)
    b = rs.randint(0, 256, (3000, 32)).astype(np.uint8)
    b[:30] = a[:30]
    b[30:40] = ~a[30:40]
    assert np.array_equal(impl().hamming_pairs(a, b), ref().hamming_pairs(a, b))


def search_by_projection(impl, ref, sizes=(4, 400, 900, 2500)):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:59-155."""
    nf, lo, hi, nmp = sizes
    total = 0
    for seed, stereo, th in ((501, 0.0, 1.0), (502, 0.0, 3.0), (503, 0.5, 3.0), (504, 0.3, 15.0)):
        fs, mps, sf, th_ = mc.sbp_case(seed, n_frames=nf, n_lo=lo, n_hi=hi, n_mp=nmp, stereo_frac=stereo, th=th)
        got, exp = impl(0.8, True).SearchByProjection(fs, mps, sf, th_), ref(0.8, True).SearchByProjection(fs, mps, sf, th_)
        same(got, exp, ("nmatches", "kp_match"), f"SearchByProjection seed {seed}")
        total += int(exp["nmatches"].sum())
    assert total > 400
    # edge cases: a frame without key points, a frame without map points, projections far outside the image
    keys = synth.synth_keypoints(300, 640, 480, 7)
    desc = synth.random_descriptors(300, 8)
    grid = np.tile(synth.frame_grid(640, 480), (3, 1))
    fs = FrameSet([0, 0, 300, 300], keys, desc, grid=grid)
    lm = synth.local_map(keys, desc, 200, 640, 480, 9)
    lm["proj_x"][:20] = -500
    lm["proj_y"][20:40] = 5000
    mps = MapPointSet([0, 50, 200, 200], lm["proj_x"], lm["proj_y"], lm["view_cos"], lm["level"], lm["flags"], lm["desc"])
    sf, _ = synth.scale_tables()
    same(impl(0.8, True).SearchByProjection(fs, mps, sf, 3.0), ref(0.8, True).SearchByProjection(fs, mps, sf, 3.0), ("nmatches", "kp_match"), "edge")


def search_by_bow(impl, ref, sizes=(5, 200, 600)):
    """SearchByBoW(KeyFrame*, KeyFrame*, ...) ORBmatcher.cc:635-768 and SearchByBoW(KeyFrame*, Frame&, ...) :211-344."""
    nf, lo, hi = sizes
    total = 0
    for kf_frame in (False, True):
        for seed, ratio, ori in ((101, 0.75, True), (102, 0.9, False), (103, 0.6, True)):
            s1, s2, i1, i2 = mc.bow_case(seed, n_frames=nf, n_lo=lo, n_hi=hi, all_pairs=True)
            got, exp = impl(ratio, ori).SearchByBoW(s1, s2, i1, i2, kf_frame=kf_frame), ref(ratio, ori).SearchByBoW(s1, s2, i1, i2, kf_frame=kf_frame)
            same(got, exp, ("nmatches", "match12"), f"SearchByBoW seed {seed} kf_frame {kf_frame}")
            total += int(exp["nmatches"].sum())
        # brute force (config #5 shape): one node holding every key point
        s1, s2, i1, i2 = mc.bow_case(201, n_frames=3, n_lo=hi, n_hi=hi + 60, single_node=True, flag_density=0.9)
        same(impl(0.75, True).SearchByBoW(s1, s2, i1, i2, kf_frame=kf_frame), ref(0.75, True).SearchByBoW(s1, s2, i1, i2, kf_frame=kf_frame),
             ("nmatches", "match12"), f"brute force kf_frame {kf_frame}")
    assert total > 300
    # heavy contention: near-identical descriptors on both sides
    rs = np.random.RandomState(5)
    base = rs.randint(0, 256, 32).astype(np.uint8)
    n1, n2 = 300, 350
    d1 = synth.flip_bits(np.tile(base, (n1, 1)), rs, 0.01)
    d2 = synth.flip_bits(np.tile(base, (n2, 1)), rs, 0.02)
    k1, k2 = synth.synth_keypoints(n1, 640, 480, 1), synth.synth_keypoints(n2, 640, 480, 2)
    s1 = FrameSet.single_node([0, n1], k1, d1, kp_flags=np.ones(n1, np.uint8))
    s2 = FrameSet.single_node([0, n2], k2, d2, kp_flags=np.ones(n2, np.uint8))
    for kf_frame in (False, True):
        same(impl(0.99, False).SearchByBoW(s1, s2, [0], [0], kf_frame=kf_frame), ref(0.99, False).SearchByBoW(s1, s2, [0], [0], kf_frame=kf_frame),
             ("nmatches", "match12"), "contention")


def search_for_triangulation(impl, ref, sizes=(6, 500, 1200)):
    """ORBmatcher::SearchForTriangulation, ORBmatcher.cc:783-975 (incl. CheckDistEpipolarLine :173-196 and the epipole rule)."""
    nf, lo, hi = sizes
    total = 0
    for seed, ori, only, stereo in ((401, False, False, 0.0), (402, True, False, 0.0), (403, False, True, 0.5), (404, True, False, 0.5)):
        s1, s2, i1, i2, F12, epi, sf, s2t = mc.tri_case(seed, n_frames=nf, n_lo=lo, n_hi=hi, stereo_frac=stereo)
        got = impl(0.6, ori).SearchForTriangulation(s1, s2, i1, i2, F12, epi, sf, s2t, bOnlyStereo=only)
        exp = ref(0.6, ori).SearchForTriangulation(s1, s2, i1, i2, F12, epi, sf, s2t, bOnlyStereo=only)
        same(got, exp, ("nmatches", "match12"), f"SearchForTriangulation seed {seed}")
        total += int(exp["nmatches"].sum())
    assert total > 100


def search_windowed(impl, ref, sizes=(6, 500, 1000, 800)):
    """The search loops of SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.cc:1540-1685) and
    SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist) (:1711-1849)."""
    nf, lo, hi, nq = sizes
    sf, _ = synth.scale_tables()
    total = 0
    for mode, stereo, th, th_dist, skip_any, ori in (("frame", 0.0, 15.0, 100, False, True), ("frame", 0.5, 7.0, 100, False, True),
                                                     ("keyframe", 0.0, 10.0, 64, True, True), ("frame", 0.3, 40.0, 100, False, False),
                                                     ("keyframe", 0.0, 15.0, 100, True, False)):
        for seed in (701, 702):
            fs, qs = mc.win_case(seed, n_frames=nf, n_lo=lo, n_hi=hi, n_q=nq, stereo_frac=stereo, mode=mode, th=th, mbf=MBF)
            got = impl(0.9, ori).SearchWindowed(fs, qs, th_dist, skip_any)
            exp = ref(0.9, ori).SearchWindowed(fs, qs, sf, th, MBF, th_dist, skip_any)
            # a key point reset to NULL by the rotation check that was NULL before is invisible in the reference's output
            km = got["kp_match"].copy()
            km[(km == -2) & (fs.kp_flags == 0)] = -1
            assert np.array_equal(km, exp["kp_match"]), f"windowed {mode} seed {seed}: kp_match differs at {np.nonzero(km != exp['kp_match'])[0][:8]}"
            assert np.array_equal(got["nmatches"], exp["nmatches"]), f"windowed {mode} seed {seed}: nmatches"
            total += int(exp["nmatches"].sum())
    assert total > 1000


def search_for_initialization(impl, ref, sizes=(3, 600, 1200)):
    """ORBmatcher::SearchForInitialization, ORBmatcher.cc:493-632."""
    nf, lo, hi = sizes
    total = 0
    for ratio, ori in ((0.9, True), (0.9, False), (0.6, True)):
        for seed in (901, 902):
            fs2, qs = mc.init_case(seed, n_frames=nf, n_lo=lo, n_hi=hi)
            got, exp = impl(ratio, ori).SearchForInitialization(fs2, qs), ref(ratio, ori).SearchForInitialization(fs2, qs)
            same(got, exp, ("nmatches", "match12"), f"SearchForInitialization seed {seed}")
            total += int(exp["nmatches"].sum())
    assert total > 300


def fuse_best(impl, ref, sizes=(3, 500, 1000, 500)):
    """The candidate loop of Fuse(KeyFrame*, const vector<MapPoint*>&, th) (ORBmatcher.cc:977-1137) with its chi-square gate."""
    nf, lo, hi, nq = sizes
    sf, s2 = synth.scale_tables()
    inv_s2 = (np.float32(1.0) / s2).astype(np.float32)
    fused = 0
    for seed, stereo in ((811, 0.0), (812, 0.5)):
        th = 12.0
        fs, qs = mc.win_case(seed, n_frames=nf, n_lo=lo, n_hi=hi, n_q=nq, stereo_frac=stereo, mode="keyframe", th=th, mbf=MBF)
        ur = qs.ur if qs.ur is not None else (qs.u - np.float32(MBF)).astype(np.float32)
        q2 = WindowQuerySet(qs.q_off, qs.u, qs.v, qs.radius, qs.min_level, qs.min_level + 1, qs.flags, qs.desc, ur=ur, angle=qs.angle)
        # Fuse reads no MapPoint occupancy in its candidate loop
        fs0 = FrameSet(fs.kp_off, fs.keys_un, fs.desc, u_right=fs.u_right, grid=fs.grid)
        got = impl(0.9, False).SearchWindowBest(fs0, q2, inv_s2, False)
        exp = ref(0.9, False).FuseBest(fs0, q2, sf, s2, th, MBF)
        want = np.where(got["q_best_dist"] <= 50, got["q_best_idx"], -1)
        assert np.array_equal(want, exp["q_best_idx"]), f"Fuse seed {seed}: differs at {np.nonzero(want != exp['q_best_idx'])[0][:8]}"
        fused += int((exp["q_best_idx"] >= 0).sum())
    assert fused > 200


# ---- committed fixtures of the reference build (tests/golden/reference_golden.npz) -------------------------------------------
GOLDEN_SIZES = {"search_by_projection": (3, 300, 600, 1200), "search_by_bow": (4, 150, 400), "search_for_triangulation": (4, 300, 700),
                "search_windowed": (3, 300, 600, 500), "search_for_initialization": (2, 400, 700), "fuse_best": (3, 300, 600, 400)}
SUITE = {"search_by_projection": search_by_projection, "search_by_bow": search_by_bow, "search_for_triangulation": search_for_triangulation,
         "search_windowed": search_windowed, "search_for_initialization": search_for_initialization, "fuse_best": fuse_best}


class TapeRef:
    """A MatcherRef stand-in that records (store given, inner given) or replays (inner None) the reference's results call by
    call.  Replaying needs neither /root/reference nor libslamref.so: the fixtures were produced by the reference's own
    ORBmatcher.cc (tests/golden/gen_reference_golden.py).  Every record carries a CRC of its inputs so drifting synthetic
    inputs are noticed."""

    def __init__(self, store, prefix, inner_factory=None):
        self.store, self.prefix, self.inner_factory, self.n = store, prefix, inner_factory, 0

    def __call__(self, nnratio=0.6, checkOri=True):
        tape = self

        class _One:
            def __getattr__(self, name):
                def call(*args, **kw):
                    import zlib
                    key = f"{tape.prefix}/{tape.n}/{name}"
                    tape.n += 1
                    crc = 0
                    for a in args:
                        for attr in ("desc", "keys_un", "u", "proj_x"):
                            v = getattr(a, attr, None)
                            if isinstance(v, np.ndarray):
                                crc = zlib.crc32(np.ascontiguousarray(v).tobytes(), crc)
                    crc = np.array([crc, int(round(nnratio * 1000)), int(checkOri)], np.int64)
                    if tape.inner_factory is not None:
                        res = getattr(tape.inner_factory(nnratio, checkOri), name)(*args, **kw)
                        for k, v in res.items():
                            tape.store[f"{key}/{k}"] = np.asarray(v)
                        tape.store[f"{key}/_crc"] = crc
                        return res
                    assert np.array_equal(tape.store[f"{key}/_crc"], crc), f"{key}: inputs or parameters drifted from the fixture"
                    pre = key + "/"
                    return {k[len(pre):]: tape.store[k] for k in tape.store if k.startswith(pre) and not k.endswith("/_crc")}
                return call
        return _One()

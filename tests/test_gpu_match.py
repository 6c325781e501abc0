"""Parity of the CUDA matcher path (through the C ABI) against the CPU oracle and the committed golden fixtures.

Bar: bit-exact match indices, distances and counts (integer work; the float geometry of SearchForTriangulation and
the window arithmetic of SearchByProjection are plain IEEE mul/add on both sides).
"""
import os

import numpy as np
import pytest

import match_cases as mc
import oracle_lib as ol
import test_oracle_matcher as tom
from orb_slam2_with_comment_b200 import synth
from orb_slam2_with_comment_b200.matcher import FrameSet, MapPointSet, match_offsets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gm():
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    made = []

    def make(nnratio=0.6, checkOri=True, **cfg):
        m = ORBmatcher(nnratio, checkOri)
        if cfg:
            m.configure(**cfg)
        made.append(m)
        return m
    yield make
    for m in made:
        m.close()


@pytest.fixture(scope="module")
def mo(oracle):
    return lambda nnratio=0.6, checkOri=True: ol.MatcherOracle(oracle, nnratio, checkOri)


def same(a, b, keys, what):
    for k in keys:
        assert np.array_equal(a[k], b[k]), f"{what}: {k} differs at {np.nonzero(a[k] != b[k])[0][:8]}"


BOW_KEYS = ("nmatches", "match12", "match_dist")
SBP_KEYS = ("nmatches", "kp_match", "mp_best_idx", "mp_best_dist", "mp_second_dist")


def test_hamming_pairs(gm):
    rs = np.random.RandomState(0)
    a = rs.randint(0, 256, (10007, 32)).astype(np.uint8)
    b = rs.randint(0, 256, (10007, 32)).astype(np.uint8)
    b[:50] = a[:50]
    b[50:60] = ~a[50:60]
    m = gm()
    assert np.array_equal(m.hamming_pairs(a, b), np.unpackbits(a ^ b, axis=1).sum(1))
    assert m.DescriptorDistance(a[70], b[70]) == int(np.unpackbits(a[70] ^ b[70]).sum())
    assert len(m.hamming_pairs(a[:0], b[:0])) == 0


# the hand-checkable rule cases of the oracle tests, run through the CUDA path
def test_micro_bow(gm): tom.test_bow_micro_rules(gm)
def test_micro_histogram(gm): tom.test_rotation_histogram_keeps_three_bins(gm)
def test_micro_triangulation(gm): tom.test_triangulation_micro_rules(gm)
def test_micro_projection(gm): tom.test_projection_micro_rules(gm)


def test_golden_fixtures(gm):
    tom.test_port_reproduces_golden_fixtures(gm)


@pytest.mark.parametrize("cfg", [dict(min_queries=0), dict(min_queries=1, min_candidates=1, queries_per_thread=4),
                                 dict(min_queries=1, min_candidates=1, queries_per_thread=8), dict()])
@pytest.mark.parametrize("kf_frame", [False, True])
def test_bow_nodes_vs_oracle(gm, mo, cfg, kf_frame):
    for seed, ratio, ori in ((101, 0.75, True), (102, 0.9, False)):
        s1, s2, i1, i2 = mc.bow_case(seed, n_frames=7, n_lo=300, n_hi=900, all_pairs=True)
        got = gm(ratio, ori, **cfg).SearchByBoW(s1, s2, i1, i2, kf_frame=kf_frame)
        exp = mo(ratio, ori).SearchByBoW(s1, s2, i1, i2, kf_frame=kf_frame)
        same(got, exp, BOW_KEYS, f"bow nodes seed {seed} cfg {cfg}")
        assert exp["nmatches"].sum() > 100


@pytest.mark.parametrize("cfg", [dict(min_queries=0), dict(queries_per_thread=4), dict(queries_per_thread=8),
                                 dict(min_queries=1, min_candidates=1, queries_per_thread=8)])
def test_bow_bruteforce_vs_oracle(gm, mo, cfg):
    # config #5 shape: one node holding every keypoint; sizes straddle the tile (256 * RQ) and chunk (256) boundaries
    for seed, n_lo, n_hi, dens in ((201, 1000, 1100, 1.0), (202, 2040, 2060, 0.9), (203, 250, 260, 0.5), (204, 2000, 2000, 1.0)):
        s1, s2, i1, i2 = mc.bow_case(seed, n_frames=4, n_lo=n_lo, n_hi=n_hi, single_node=True, flag_density=dens)
        got = gm(0.75, True, **cfg).SearchByBoW(s1, s2, i1, i2)
        exp = mo(0.75, True).SearchByBoW(s1, s2, i1, i2)
        same(got, exp, BOW_KEYS, f"brute force seed {seed} cfg {cfg}")
        assert exp["nmatches"].sum() > 100


def test_bow_heavy_contention_forces_rescans(gm, mo):
    """Many near-identical queries compete for a handful of candidates: the top-K lists run dry and phase B must
    rescan with the taken mask."""
    rs = np.random.RandomState(5)
    base = rs.randint(0, 256, 32).astype(np.uint8)
    n1, n2 = 600, 700
    d1 = synth.flip_bits(np.tile(base, (n1, 1)), rs, 0.01)
    d2 = synth.flip_bits(np.tile(base, (n2, 1)), rs, 0.02)
    k1, k2 = synth.synth_keypoints(n1, 640, 480, 1), synth.synth_keypoints(n2, 640, 480, 2)
    s1 = FrameSet.single_node([0, n1], k1, d1, kp_flags=np.ones(n1, np.uint8))
    s2 = FrameSet.single_node([0, n2], k2, d2, kp_flags=np.ones(n2, np.uint8))
    for cfg in (dict(min_queries=0), dict(min_queries=1, min_candidates=1)):
        got = gm(0.99, False, **cfg).SearchByBoW(s1, s2, [0], [0], kf_frame=True)
        exp = mo(0.99, False).SearchByBoW(s1, s2, [0], [0], kf_frame=True)
        same(got, exp, BOW_KEYS, f"contention cfg {cfg}")
    assert exp["nmatches"][0] > 20


def test_bow_edge_cases(gm, mo):
    m, o = gm(0.75, True), mo(0.75, True)
    # zero pairs
    s1, s2, i1, i2 = mc.bow_case(301, n_frames=3, n_lo=50, n_hi=80)
    r = m.SearchByBoW(s1, s2, i1[:0], i2[:0])
    assert len(r["match12"]) == 0 and len(r["nmatches"]) == 0
    # an empty frame on either side, and frames with no common node
    kp_off = np.array([0, 0, 40, 80], np.int32)
    keys = synth.synth_keypoints(80, 640, 480, 3)
    desc = synth.random_descriptors(80, 4)
    fl = np.ones(80, np.uint8)
    fs = FrameSet(kp_off, keys, desc, kp_flags=fl, fv_node_off=[0, 0, 1, 2], fv_node_id=[3, 9], fv_feat_off=[0, 40, 80],
                  fv_feat=np.concatenate([np.arange(40), np.arange(40)]).astype(np.int32))
    pairs1, pairs2 = np.array([0, 1, 1, 2, 1], np.int32), np.array([1, 0, 2, 1, 1], np.int32)
    got, exp = m.SearchByBoW(fs, fs, pairs1, pairs2), o.SearchByBoW(fs, fs, pairs1, pairs2)
    same(got, exp, BOW_KEYS, "edge")
    assert got["nmatches"].tolist()[:4] == [0, 0, 0, 0] and got["nmatches"][4] > 0   # only the self pair shares a node
    # no query / no candidate carries a MapPoint
    s1, s2, i1, i2 = mc.bow_case(302, n_frames=3, n_lo=60, n_hi=90, flag_density=0.0)
    assert m.SearchByBoW(s1, s2, i1, i2)["nmatches"].sum() == 0


@pytest.mark.parametrize("stereo", [0.0, 0.5])
def test_triangulation_vs_oracle(gm, mo, stereo):
    for seed, ori, only in ((401, False, False), (402, True, False), (403, False, True)):
        if only and stereo == 0.0:
            continue
        s1, s2, i1, i2, F12, epi, sf, s2t = mc.tri_case(seed, n_frames=8, n_lo=800, n_hi=2000, stereo_frac=stereo)
        got = gm(0.6, ori).SearchForTriangulation(s1, s2, i1, i2, F12, epi, sf, s2t, bOnlyStereo=only)
        exp = mo(0.6, ori).SearchForTriangulation(s1, s2, i1, i2, F12, epi, sf, s2t, bOnlyStereo=only)
        same(got, exp, BOW_KEYS, f"triangulation seed {seed}")
        assert exp["nmatches"].sum() > 50


@pytest.mark.parametrize("stereo,th", [(0.0, 1.0), (0.0, 3.0), (0.5, 3.0), (0.3, 15.0)])
def test_projection_vs_oracle(gm, mo, stereo, th):
    for seed in (501, 502):
        fs, mps, sf, th_ = mc.sbp_case(seed, n_frames=5, n_lo=600, n_hi=1400, n_mp=4000, stereo_frac=stereo, th=th)
        got = gm(0.8, True).SearchByProjection(fs, mps, sf, th_)
        exp = mo(0.8, True).SearchByProjection(fs, mps, sf, th_)
        same(got, exp, SBP_KEYS, f"projection seed {seed} th {th}")
        assert exp["nmatches"].sum() > 200


def test_projection_edge_cases(gm, mo):
    m, o = gm(0.8, True), mo(0.8, True)
    sf, _ = synth.scale_tables()
    # a frame without keypoints, a frame without map points, projections far outside the image
    keys = synth.synth_keypoints(300, 640, 480, 7)
    desc = synth.random_descriptors(300, 8)
    grid = np.tile(synth.frame_grid(640, 480), (3, 1))
    fs = FrameSet([0, 0, 300, 300], keys, desc, grid=grid)
    lm = synth.local_map(keys, desc, 200, 640, 480, 9)
    lm["proj_x"][:20] = -500
    lm["proj_y"][20:40] = 5000
    mps = MapPointSet([0, 50, 200, 200], lm["proj_x"], lm["proj_y"], lm["view_cos"], lm["level"], lm["flags"], lm["desc"])
    got, exp = m.SearchByProjection(fs, mps, sf, 3.0), o.SearchByProjection(fs, mps, sf, 3.0)
    same(got, exp, SBP_KEYS, "projection edge")
    assert got["nmatches"][0] == 0 and got["nmatches"][2] == 0 and got["nmatches"][1] > 0
    # all keypoints in one grid cell (clustered): exercises the per-cell ordering
    keys2 = keys.copy()
    keys2["x"] = 321 + (np.arange(300) % 7)
    keys2["y"] = 241 + (np.arange(300) % 5)
    keys2["octave"] = np.arange(300) % 3
    fs2 = FrameSet([0, 300], keys2, desc, grid=grid[:1])
    lm2 = synth.local_map(keys2, desc, 400, 640, 480, 10)
    lm2["level"][:] = np.arange(400) % 3
    mps2 = MapPointSet([0, 400], lm2["proj_x"], lm2["proj_y"], lm2["view_cos"], lm2["level"], lm2["flags"], lm2["desc"])
    same(m.SearchByProjection(fs2, mps2, sf, 3.0), o.SearchByProjection(fs2, mps2, sf, 3.0), SBP_KEYS, "clustered")


def test_device_resident_variant_equals_host_variant(gm):
    import torch
    m = gm(0.75, True)
    s1, s2, i1, i2 = mc.bow_case(601, n_frames=5, n_lo=900, n_hi=1100, single_node=True, flag_density=1.0)
    host = m.SearchByBoW(s1, s2, i1, i2)
    h1 = m.upload(s1)
    off, total = match_offsets(s1, i1)
    dev = torch.device("cuda", 0)
    d12 = torch.full((total,), -7, dtype=torch.int32, device=dev)
    dd = torch.full((total,), -7, dtype=torch.int32, device=dev)
    dn = torch.zeros(len(i1), dtype=torch.int32, device=dev)
    for _ in range(2):   # twice: scratch reuse must not change results
        m.search_by_bow_dev(h1, h1, i1, i2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
        m.sync()
        assert np.array_equal(d12.cpu().numpy(), host["match12"]) and np.array_equal(dd.cpu().numpy(), host["match_dist"])
        assert np.array_equal(dn.cpu().numpy(), host["nmatches"])
    ms, evals = m.last_stats()
    n = np.array([s1.n_kp(int(f)) for f in range(s1.n_frames)])
    assert evals >= int(sum(n[a] * n[b] for a, b in zip(i1, i2))) and ms > 0
    m.release(h1)


def test_full_size_bruteforce_properties(gm):
    """Config #5 at full size (2000 x 2000) — size-independent properties instead of the (slow) CPU oracle:
    injectivity of the greedy assignment, distances recomputed from the descriptors, planted matches recovered."""
    A, B, angA, angB = synth.bruteforce_sets(6, 2000, 900)
    n_sets, n = A.shape[:2]
    kp_off = np.arange(n_sets + 1, dtype=np.int32) * n
    kA = np.zeros(n_sets * n, synth.KP_DTYPE); kA["angle"] = angA.ravel()
    kB = np.zeros(n_sets * n, synth.KP_DTYPE); kB["angle"] = angB.ravel()
    fl = np.ones(n_sets * n, np.uint8)
    sA = FrameSet.single_node(kp_off, kA, A.reshape(-1, 32), kp_flags=fl)
    sB = FrameSet.single_node(kp_off, kB, B.reshape(-1, 32), kp_flags=fl)
    idx = np.arange(n_sets, dtype=np.int32)
    m = gm(0.75, False)
    r = m.SearchByBoW(sA, sB, idx, idx)
    m12 = r["match12"].reshape(n_sets, n)
    for s in range(n_sets):
        hit = np.nonzero(m12[s] >= 0)[0]
        assert len(hit) == r["nmatches"][s] and len(np.unique(m12[s][hit])) == len(hit)          # injective
        d = np.unpackbits(A[s][hit] ^ B[s][m12[s][hit]], axis=1).sum(1)
        assert np.array_equal(d, r["match_dist"].reshape(n_sets, n)[s][hit]) and d.max() < 50     # TH_LOW exclusive
        assert len(hit) > 0.4 * n                                                                  # ~half the rows were planted
    # the same search is idempotent and independent of the kernel choice
    for cfg in (dict(min_queries=0), dict(queries_per_thread=8)):
        r2 = gm(0.75, False, **cfg).SearchByBoW(sA, sB, idx, idx)
        same(r, r2, BOW_KEYS, f"cfg {cfg}")


@pytest.mark.parametrize("mode,stereo,th,th_dist,skip_any,ori", [("frame", 0.0, 15.0, 100, False, True), ("frame", 0.5, 7.0, 100, False, True),
                                                                 ("keyframe", 0.0, 10.0, 64, True, True), ("frame", 0.3, 40.0, 100, False, False)])
def test_windowed_search_vs_oracle(gm, mo, mode, stereo, th, th_dist, skip_any, ori):
    """The search loop of SearchByProjection(CurrentFrame, LastFrame) / (CurrentFrame, KeyFrame) (ORBmatcher.cc:1540, :1711)."""
    for seed in (701, 702):
        fs, qs = mc.win_case(seed, n_frames=6, n_lo=700, n_hi=1500, n_q=1200, stereo_frac=stereo, mode=mode, th=th)
        got = gm(0.9, ori).SearchWindowed(fs, qs, th_dist, skip_any)
        exp = mo(0.9, ori).SearchWindowed(fs, qs, th_dist, skip_any)
        same(got, exp, ("nmatches", "kp_match", "q_best_idx", "q_best_dist"), f"windowed seed {seed} {mode}")
        assert exp["nmatches"].sum() > 500


def test_is_in_frustum_matches_oracle():
    """Frame::isInFrustum + PredictScale (Frame.cc:274-342): visibility and mTrack* fields, float results bit for bit."""
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    args = mc.frustum_case(3)
    exp = ol.is_in_frustum(ol.load_port(), *args)
    m = ORBmatcher()
    got = m.isInFrustum(*args)
    assert 0.15 < exp["in_view"].mean() < 0.85          # every exit of the function is exercised
    assert np.array_equal(got["in_view"], exp["in_view"])
    for k in ("proj_x", "proj_y", "proj_xr", "view_cos"):
        assert got[k].tobytes() == exp[k].tobytes(), k
    # predicted level: glibc's logf vs a correctly rounded logf can differ when the quotient sits on an integer
    diff = np.nonzero(got["level"] != exp["level"])[0]
    assert len(diff) <= 2 and np.all(np.abs(got["level"][diff] - exp["level"][diff]) == 1), diff
    assert len(np.unique(exp["level"][exp["in_view"] == 1])) == 8
    m.close()


def test_distinctive_descriptors_match_oracle():
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-316): least-median row per map point, first wins."""
    from orb_slam2_with_comment_b200 import capi
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    off, desc = mc.distinctive_case(11)
    ei, em = ol.distinctive_descriptors(ol.load_port(), off, desc)
    m = ORBmatcher()
    gi, gm = m.ComputeDistinctiveDescriptors(off, desc)
    assert np.array_equal(gi, ei) and np.array_equal(gm[ei >= 0], em[ei >= 0])
    assert ei[0] == -1 and ei[1] == 0 and (ei > 0).mean() > 0.5
    with pytest.raises(capi.OrbGpuError):
        m.ComputeDistinctiveDescriptors(np.array([0, 257], np.int32), np.zeros((257, 32), np.uint8))
    m.close()


@pytest.mark.parametrize("stereo,gate,skip", [(0.0, True, False), (0.5, True, False), (0.5, False, False), (0.0, False, True)])
def test_window_best_vs_oracle(gm, mo, stereo, gate, skip):
    """The candidate loops of Fuse (ORBmatcher.cc:1051-1112 with the chi-square gate, :1211-1246 without) and SearchBySim3
    (:1363-1401: candidates already matched are skipped): independent queries, best candidate only."""
    from orb_slam2_with_comment_b200.matcher import WindowQuerySet
    _, s2 = synth.scale_tables()
    inv_s2 = (np.float32(1.0) / s2).astype(np.float32) if gate else None
    for seed in (811, 812):
        fs, qs = mc.win_case(seed, n_frames=5, n_lo=700, n_hi=1500, n_q=1200, stereo_frac=stereo, mode="keyframe", th=12.0)
        # Fuse / SearchBySim3 search levels [l-1, l]
        q2 = WindowQuerySet(qs.q_off, qs.u, qs.v, qs.radius, qs.min_level, qs.min_level + 1, qs.flags, qs.desc,
                            ur=qs.ur if qs.ur is not None else (qs.u - 5.0).astype(np.float32), angle=qs.angle)
        got = gm(0.9, False).SearchWindowBest(fs, q2, inv_s2, skip)
        exp = mo(0.9, False).SearchWindowBest(fs, q2, inv_s2, skip)
        same(got, exp, ("q_best_idx", "q_best_dist"), f"window best seed {seed}")
        found = (exp["q_best_idx"] >= 0).sum()
        assert found > 300
        if gate:   # the gate must reject some candidates a plain search accepts
            plain = mo(0.9, False).SearchWindowBest(fs, q2, None, skip)
            assert (plain["q_best_idx"] != exp["q_best_idx"]).sum() > 50


@pytest.mark.parametrize("ratio,ori", [(0.9, True), (0.9, False), (0.6, True)])
def test_search_for_initialization_vs_oracle(gm, mo, ratio, ori):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:493-632) incl. the stealing rule and the stale histogram entries."""
    for seed in (901, 902):
        fs2, qs = mc.init_case(seed)
        got = gm(ratio, ori).SearchForInitialization(fs2, qs)
        exp = mo(ratio, ori).SearchForInitialization(fs2, qs)
        same(got, exp, ("nmatches", "match12"), f"init seed {seed}")
        assert exp["nmatches"].sum() > 150


def test_search_local_points_chain_on_device():
    """Tracking::SearchLocalPoints (Tracking.cc:1150-1200): isInFrustum on the device straight into a device-resident map-point
    set, then SearchByProjection on it — against oracle isInFrustum -> host MapPointSet -> oracle SearchByProjection."""
    import torch
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    fs, sf, cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref, fl, dd = mc.local_map_case(5)
    fr = ol.is_in_frustum(ol.load_port(), cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref)
    mps = MapPointSet(off, fr["proj_x"], fr["proj_y"], fr["view_cos"], fr["level"], (fl & 0xFE) | fr["in_view"], dd, proj_xr=fr["proj_xr"])
    exp = ol.MatcherOracle(ol.load_port(), 0.8, True).SearchByProjection(fs, mps, sf, 3.0)
    m = ORBmatcher(0.8, True)
    hf = m.upload(fs)
    hm = m.project_mappoints(cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref, fl, dd)
    dev = torch.device("cuda:0")
    nkp, nmp, nf = int(fs.kp_off[-1]), int(off[-1]), fs.n_frames
    o = [torch.zeros(nkp, dtype=torch.int32, device=dev)] + [torch.zeros(nmp, dtype=torch.int32, device=dev) for _ in range(3)]
    o.append(torch.zeros(nf, dtype=torch.int32, device=dev))
    m.search_by_projection_dev(hf, hm, sf, 3.0, *[t.data_ptr() for t in o])
    m.sync()
    assert np.array_equal(o[4].cpu().numpy(), exp["nmatches"]) and exp["nmatches"].sum() > 600
    assert np.array_equal(o[0].cpu().numpy(), exp["kp_match"])
    assert np.array_equal(o[1].cpu().numpy(), exp["mp_best_idx"]) and np.array_equal(o[2].cpu().numpy(), exp["mp_best_dist"])
    assert 0.4 < fr["in_view"].mean() < 0.95
    m.release_mappoints(hm); m.release(hf); m.close()

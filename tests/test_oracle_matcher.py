"""The matcher oracle (oracle/match_oracle.cc, the C++ restatement of ORBmatcher.cc) against
  (a) an independent popcount for DescriptorDistance,
  (b) hand-checkable micro-cases of the rules that are easy to get wrong (tie-breaking, thresholds, greedy masks),
  (c) the committed golden fixtures produced by the independent Python restatement (tests/golden/gen_matcher_golden.py).
The reference ships no tests for this path, so this is all the pinning there is ("parity unpinned by the reference").
"""
import os

import numpy as np
import pytest

import match_cases as mc
import oracle_lib as ol
from orb_slam2_with_comment_b200 import synth
from orb_slam2_with_comment_b200.matcher import FrameSet, MapPointSet

GOLD = os.path.join(os.path.dirname(__file__), "golden", "matcher_golden.npz")


@pytest.fixture(scope="module")
def mo(oracle):
    return lambda nnratio=0.6, checkOri=True: ol.MatcherOracle(oracle, nnratio, checkOri)


def test_descriptor_distance_is_popcount(mo):
    rs = np.random.RandomState(0)
    a = rs.randint(0, 256, (500, 32)).astype(np.uint8)
    b = rs.randint(0, 256, (500, 32)).astype(np.uint8)
    b[:50] = a[:50]
    b[50:60] = ~a[50:60]
    exp = np.unpackbits(a ^ b, axis=1).sum(1)
    assert np.array_equal(mo().hamming_pairs(a, b), exp)
    assert exp[:50].max() == 0 and exp[50:60].min() == 256


def _desc_at_distance(base, d):
    """A copy of `base` with exactly the first d bits flipped."""
    bits = np.unpackbits(base)
    bits[:d] ^= 1
    return np.packbits(bits)


def _kps(n, angle=0.0):
    k = np.zeros(n, synth.KP_DTYPE)
    k["x"] = 100 + 10 * np.arange(n)
    k["y"] = 100
    k["angle"] = angle
    k["class_id"] = -1
    return k


def _single(descs, flags=None, angles=None):
    k = _kps(len(descs))
    if angles is not None:
        k["angle"] = angles
    return FrameSet.single_node([0, len(descs)], k, np.stack(descs), kp_flags=np.ones(len(descs), np.uint8) if flags is None else flags)


def test_bow_micro_rules(mo):
    base = np.zeros(32, np.uint8)
    q = [base.copy()]
    # best 10, second 40: 10 < 0.6*40 accepted
    s2 = _single([_desc_at_distance(base, 40), _desc_at_distance(base, 10)])
    r = mo(0.6, False).SearchByBoW(_single(q), s2, [0], [0])
    assert r["match12"].tolist() == [1] and r["match_dist"].tolist() == [10] and r["nmatches"].tolist() == [1]
    # first of two equal minima wins (strict <); a ratio above 1 lets the tie through
    s2 = _single([_desc_at_distance(base, 40), _desc_at_distance(base, 10), _desc_at_distance(base, 10)])
    assert mo(1.5, False).SearchByBoW(_single(q), s2, [0], [0])["match12"].tolist() == [1]
    # with two candidates at the same minimum the second best equals the best, so the ratio test rejects
    r = mo(0.6, False).SearchByBoW(_single(q), _single([_desc_at_distance(base, 10), _desc_at_distance(base, 10)]), [0], [0])
    assert r["match12"].tolist() == [-1]           # 10 < 0.6*10 is false
    # TH_LOW: KeyFrame-KeyFrame is exclusive (<50, :711), KeyFrame-Frame inclusive (<=50, :284)
    far = [_desc_at_distance(base, 50), _desc_at_distance(base, 200)]
    assert mo(0.9, False).SearchByBoW(_single(q), _single(far), [0], [0], kf_frame=False)["match12"].tolist() == [-1]
    assert mo(0.9, False).SearchByBoW(_single(q), _single(far), [0], [0], kf_frame=True)["match12"].tolist() == [0]
    # greedy mask: two identical queries, the second must take the runner-up
    r = mo(0.9, False).SearchByBoW(_single([base, base]), _single([_desc_at_distance(base, 3), _desc_at_distance(base, 30),
                                                                   _desc_at_distance(base, 200)]), [0], [0])
    assert r["match12"].tolist() == [0, 1] and r["match_dist"].tolist() == [3, 30]
    # candidates of the KeyFrame-KeyFrame variant need a MapPoint (flag), queries always do
    fl = np.array([0, 1, 1], np.uint8)
    r = mo(0.9, False).SearchByBoW(_single([base]), _single([_desc_at_distance(base, 3), _desc_at_distance(base, 30),
                                                             _desc_at_distance(base, 200)], flags=fl), [0], [0])
    assert r["match12"].tolist() == [1]
    r = mo(0.9, False).SearchByBoW(_single([base], flags=np.zeros(1, np.uint8)), _single([_desc_at_distance(base, 3)]), [0], [0])
    assert r["match12"].tolist() == [-1] and r["nmatches"].tolist() == [0]


def test_rotation_histogram_keeps_three_bins(mo):
    base = np.zeros(32, np.uint8)
    rs = np.random.RandomState(1)
    n = 40
    descs = [rs.randint(0, 256, 32).astype(np.uint8) for _ in range(n)]
    # rotation differences: 30 matches at ~0 deg (bin 0), 6 at 60 (bin 2), 3 at 120 (bin 4), 1 at 200 (bin 7: dropped, < third)
    rot = np.array([0.0] * 30 + [60.0] * 6 + [120.0] * 3 + [200.0])
    s1 = _single(descs, angles=rot.astype(np.float32))
    s2 = _single([d.copy() for d in descs], angles=np.zeros(n, np.float32))
    r = mo(0.9, True).SearchByBoW(s1, s2, [0], [0])
    assert r["nmatches"].tolist() == [39]
    assert r["match12"][:39].tolist() == list(range(39)) and r["match12"][39] == -1
    r = mo(0.9, False).SearchByBoW(s1, s2, [0], [0])
    assert r["nmatches"].tolist() == [40]


def test_triangulation_micro_rules(mo):
    base = np.zeros(32, np.uint8)
    sf, s2t = synth.scale_tables()
    F = np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32)   # x1^T F = (0, 1, -y1): horizontal epipolar lines
    far_epipole = np.array([[-1e6, -1e6]], np.float32)
    k1 = _kps(1)
    k2 = _kps(3)
    k2["y"] = [100, 100, 100]
    # two candidates at distance 20: the LAST one wins (`dist>bestDist` -> continue, :882)
    d2 = [_desc_at_distance(base, 20), _desc_at_distance(base, 20), _desc_at_distance(base, 60)]
    s1 = FrameSet.single_node([0, 1], k1, np.stack([base]))
    s2 = FrameSet.single_node([0, 3], k2, np.stack(d2))
    r = mo(0.6, False).SearchForTriangulation(s1, s2, [0], [0], F, far_epipole, sf, s2t)
    assert r["match12"].tolist() == [1] and r["match_dist"].tolist() == [20]
    # off the epipolar line by more than sqrt(3.84): rejected; the other one remains
    k2b = k2.copy()
    k2b["y"] = [100, 103, 100]
    r = mo(0.6, False).SearchForTriangulation(s1, FrameSet.single_node([0, 3], k2b, np.stack(d2)), [0], [0], F, far_epipole, sf, s2t)
    assert r["match12"].tolist() == [0]
    # candidate closer than 10 px (level 0: 100*1.0) to the epipole is skipped in mono pairs
    ep = np.array([[k2["x"][1] + 3, 100.0]], np.float32)
    r = mo(0.6, False).SearchForTriangulation(s1, s2, [0], [0], F, ep, sf, s2t)
    assert r["match12"].tolist() == [0]
    # keypoints that already have a MapPoint are skipped on both sides
    s2f = FrameSet.single_node([0, 3], k2, np.stack(d2), kp_flags=np.array([0, 1, 0], np.uint8))
    r = mo(0.6, False).SearchForTriangulation(s1, s2f, [0], [0], F, far_epipole, sf, s2t)
    assert r["match12"].tolist() == [0]
    # distance 50 is accepted (bestDist starts at TH_LOW, `dist>TH_LOW` skips), 51 is not
    for d, exp in ((50, 0), (51, -1)):
        s2d = FrameSet.single_node([0, 1], k2[:1], np.stack([_desc_at_distance(base, d)]))
        assert mo(0.6, False).SearchForTriangulation(s1, s2d, [0], [0], F, far_epipole, sf, s2t)["match12"].tolist() == [exp]


def test_projection_micro_rules(mo):
    base = np.zeros(32, np.uint8)
    sf, _ = synth.scale_tables()
    k = _kps(3)
    k["x"] = [200, 203, 400]
    k["y"] = [200, 200, 200]
    k["octave"] = [1, 1, 1]
    desc = np.stack([_desc_at_distance(base, 30), _desc_at_distance(base, 35), _desc_at_distance(base, 5)])
    grid = synth.frame_grid(640, 480).reshape(1, 4)

    def run(flags_kp, mp_flags, nn=0.8, n_mp=1, th=1.0, level=1):
        fs = FrameSet([0, 3], k, desc, kp_flags=np.array(flags_kp, np.uint8), grid=grid)
        mps = MapPointSet([0, n_mp], [201.0] * n_mp, [200.0] * n_mp, [0.9999] * n_mp, [level] * n_mp, mp_flags, np.stack([base] * n_mp))
        return mo(nn, True).SearchByProjection(fs, mps, sf, th)

    # window 2.5*1.2 = 3 px: keypoints 0 and 1 are candidates (same level): 30 > 0.8*35 -> ratio test rejects
    r = run([0, 0, 0], [5])
    assert r["mp_best_idx"].tolist() == [0] and r["mp_best_dist"].tolist() == [30] and r["mp_second_dist"].tolist() == [35]
    assert r["nmatches"].tolist() == [0] and r["kp_match"].tolist() == [-1, -1, -1]
    # a keypoint that already holds an observed MapPoint is skipped -> single candidate, accepted
    r = run([0, 1, 0], [5])
    assert r["mp_best_idx"].tolist() == [0] and r["mp_second_dist"].tolist() == [256] and r["kp_match"].tolist() == [0, -1, -1]
    # second identical map point: the first match (with observations) blocks keypoint 0, so it falls to keypoint 1
    r = run([0, 0, 0], [5, 5], nn=0.9, n_mp=2)
    assert r["mp_best_idx"].tolist() == [0, 1] and r["kp_match"].tolist() == [0, 1, -1] and r["nmatches"].tolist() == [2]
    # ... unless the first map point has no observations: then both write keypoint 0 and the later one stays
    r = run([0, 0, 0], [1, 5], nn=0.9, n_mp=2)
    assert r["mp_best_idx"].tolist() == [0, 0] and r["kp_match"].tolist() == [1, -1, -1] and r["nmatches"].tolist() == [2]
    # not in view / bad map points are skipped
    assert run([0, 0, 0], [4], nn=0.9)["nmatches"].tolist() == [0]
    assert run([0, 0, 0], [7], nn=0.9)["nmatches"].tolist() == [0]
    # level filter: predicted level 3 only admits octaves 2..3
    assert run([0, 0, 0], [5], nn=0.9, level=3)["mp_best_idx"].tolist() == [-1]


def _check(res, G, name, keys):
    for k in keys:
        assert np.array_equal(res[k], G[f"{name}__{k}"]), f"{name}: {k} differs from the golden fixture"


def test_port_reproduces_golden_fixtures(mo):
    from golden import gen_matcher_golden as gen
    G = np.load(GOLD)
    assert set(str(c) for c in G["cases"]) == set(gen.CASES)
    for name, (kind, g, mk) in gen.CASES.items():
        if kind == "bow":
            s1, s2, i1, i2 = mc.bow_case(**g)
            assert np.array_equal(gen.fs_checksum(s1), G[f"{name}__input_crc"]), "synthetic inputs drifted from the fixture"
            res = mo(mk["nnratio"], mk["check_ori"]).SearchByBoW(s1, s2, i1, i2, kf_frame=mk["kf_frame"])
            _check(res, G, name, ("nmatches", "match12", "match_dist"))
        elif kind == "tri":
            s1, s2, i1, i2, F12, epi, sf, s2t = mc.tri_case(**g)
            assert np.array_equal(gen.fs_checksum(s1), G[f"{name}__input_crc"])
            res = mo(0.6, mk["check_ori"]).SearchForTriangulation(s1, s2, i1, i2, F12, epi, sf, s2t, bOnlyStereo=mk["only_stereo"])
            _check(res, G, name, ("nmatches", "match12", "match_dist"))
        elif kind == "best":
            fs, qs, inv = gen.best_inputs(g)
            res = mo(0.6, False).SearchWindowBest(fs, qs, inv if mk["gate"] else None, mk["skip"])
            _check(res, G, name, ("q_best_idx", "q_best_dist"))
            res["nmatches"] = np.array([(res["q_best_idx"] >= 0).sum()])
        elif kind == "init":
            fs2, qs = mc.init_case(**g)
            res = mo(mk["nnratio"], mk["check_ori"]).SearchForInitialization(fs2, qs)
            _check(res, G, name, ("nmatches", "match12"))
        elif kind == "distinctive":
            off, desc = mc.distinctive_case(**g)
            idx, med = ol.distinctive_descriptors(ol.load_port(), off, desc)
            assert np.array_equal(idx, G[f"{name}__best_idx"]) and np.array_equal(med, G[f"{name}__best_median"])
            res = {"nmatches": np.array([(idx > 0).sum()])}
        elif kind == "frustum":
            args = mc.frustum_case(**g)
            res = ol.is_in_frustum(ol.load_port(), *args)
            for k in ("in_view", "proj_x", "proj_y", "proj_xr", "view_cos"):
                assert res[k].tobytes() == G[f"{name}__{k}"].tobytes(), f"{name}: {k}"
            # glibc logf (C oracle) vs correctly rounded log (Python restatement): a level may differ by one on an exact boundary
            diff = np.nonzero(res["level"] != G[f"{name}__level"])[0]
            assert len(diff) <= 1 and np.all(np.abs(res["level"][diff] - G[f"{name}__level"][diff]) == 1)
            res["nmatches"] = np.array([res["in_view"].sum()])
        elif kind == "win":
            fs, qs = mc.win_case(**g)
            res = mo(0.6, mk["check_ori"]).SearchWindowed(fs, qs, mk["th_dist"], mk["skip_any"])
            _check(res, G, name, ("nmatches", "kp_match", "q_best_idx", "q_best_dist"))
            if mk["check_ori"]:
                assert (res["kp_match"] == -2).any(), f"{name}: the rotation check removed nothing"
        else:
            fs, mps, sf, th = mc.sbp_case(**g)
            res = mo(mk["nnratio"], True).SearchByProjection(fs, mps, sf, th)
            _check(res, G, name, ("nmatches", "kp_match", "mp_best_idx", "mp_best_dist", "mp_second_dist"))
        assert int(res["nmatches"].sum()) > 0, f"{name}: degenerate case (no matches)"

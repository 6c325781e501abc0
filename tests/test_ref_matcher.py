"""The CPU port (oracle/match_oracle.cc, stereo_oracle.cc) against the REFERENCE ITSELF: oracle/_ref/libslamref.so = the reference's
ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc, ORBextractor.cc and DBoW2 compiled from /root/reference where they lie
(oracle/Makefile), driven through real Frame / KeyFrame / MapPoint objects (oracle/slam_ref.cc).  This is what pins the port; the
-m gpu twin of this file (tests/test_gpu_ref_matcher.py) compares the CUDA path with the same library directly."""
import numpy as np
import pytest

import match_cases as mc
import oracle_lib as ol
import ref_parity as rp
from orb_slam2_with_comment_b200 import synth


@pytest.fixture(scope="module")
def slamref():
    lib = ol.load_slam_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libslamref.so not built (needs /root/reference at build time)")
    return lib


@pytest.fixture(scope="module")
def ref(slamref):
    return lambda nnratio=0.6, checkOri=True: ol.MatcherRef(slamref, nnratio, checkOri)


@pytest.fixture(scope="module")
def port(oracle):
    return lambda nnratio=0.6, checkOri=True: ol.MatcherOracle(oracle, nnratio, checkOri)


def test_descriptor_distance(port, ref): rp.descriptor_distance(port, ref)
def test_search_by_projection(port, ref): rp.search_by_projection(port, ref)
def test_search_by_bow(port, ref): rp.search_by_bow(port, ref)
def test_search_for_triangulation(port, ref): rp.search_for_triangulation(port, ref)
def test_search_windowed(port, ref): rp.search_windowed(port, ref)
def test_search_for_initialization(port, ref): rp.search_for_initialization(port, ref)
def test_fuse_candidate_loop(port, ref): rp.fuse_best(port, ref)


def test_is_in_frustum(oracle, slamref):
    """Frame::isInFrustum + MapPoint::PredictScale (Frame.cc:274-342, MapPoint.cc:421-436): every output bit for bit."""
    args = mc.frustum_case(3, raw=True)
    cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref, rmin = args
    got = ol.is_in_frustum(oracle, cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref)
    exp = ol.ref_is_in_frustum(slamref, cam, lsf, nl, cosl, off, P, Nn, rmin, dref)
    assert 0.15 < exp["in_view"].mean() < 0.85
    assert np.array_equal(got["in_view"], exp["in_view"])
    for k in ("proj_x", "proj_y", "proj_xr", "view_cos"):
        assert got[k].tobytes() == exp[k].tobytes(), k
    assert np.array_equal(got["level"], exp["level"])


def test_distinctive_descriptors(oracle, slamref):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-316) on real observation maps."""
    off, desc = mc.distinctive_case(11, n_points=600)
    idx, _ = ol.distinctive_descriptors(oracle, off, desc)
    has, best = ol.ref_distinctive_descriptors(slamref, off, desc)
    assert np.array_equal(has.astype(bool), idx >= 0)
    sel = idx >= 0
    assert np.array_equal(desc[off[:-1][sel] + idx[sel]], best[sel])
    assert sel.sum() > 500


def test_features_in_area(oracle, slamref):
    """Frame::AssignFeaturesToGrid + GetFeaturesInArea (Frame.cc:232-247, :353-410): candidate lists in the reference's order, seen
    through SearchByProjection with a single map point and all-zero descriptors (every candidate ties, the first one wins)."""
    from orb_slam2_with_comment_b200.matcher import FrameSet
    keys = synth.synth_keypoints(900, 640, 480, 3)
    fs = FrameSet([0, 900], keys, np.zeros((900, 32), np.uint8), grid=synth.frame_grid(640, 480)[None])
    rs = np.random.RandomState(1)
    for _ in range(60):
        x, y, r = rs.uniform(-20, 660), rs.uniform(-20, 500), rs.uniform(1, 60)
        lo, hi = sorted(rs.randint(-1, 8, 2))
        got = ol.ref_features_in_area(slamref, fs, 0, x, y, r, int(lo), int(hi))
        k = keys
        inside = (np.abs(k["x"] - np.float32(x)) < np.float32(r)) & (np.abs(k["y"] - np.float32(y)) < np.float32(r))
        if lo > 0 or hi >= 0:   # bCheckLevels (Frame.cc:380)
            inside &= k["octave"] >= lo
            if hi >= 0:
                inside &= k["octave"] <= hi
        assert set(got.tolist()) == set(np.nonzero(inside)[0].tolist())


def test_stereo_matches(oracle, slamref):
    """Frame::ComputeStereoMatches (Frame.cc:501-675) through the reference's stereo Frame constructor against the stereo port,
    bit for bit.  The reference reads Frame::mb before its constructor sets it (Frame.cc:88 vs :113; see oracle/slam_ref.cc), so
    the value is supplied: the intended mbf / fx, and a small one (maxD large: the disparity window stops gating)."""
    mbf, fx = np.float32(386.1448), np.float32(718.856)
    for seed, (w, h, nf), mb in ((0, (1241, 376, 2000), mbf / fx), (1, (1241, 376, 2000), np.float32(0.05)), (2, (752, 480, 1200), mbf / fx)):
        L, R = synth.stereo_pair(w, h, seed)
        kp, desc, ur, dp = ol.ref_stereo_frame(slamref, L, R, nf, mbf, mb, fx=fx, fy=fx)
        S = ol.stereo_inputs(oracle, L, R, nf)
        assert np.array_equal(S["kpL"], kp) and np.array_equal(S["descL"], desc)
        pur, pdp, _ = ol.stereo_matches(oracle, S, mb, mbf)
        assert pur.tobytes() == ur.tobytes() and pdp.tobytes() == dp.tobytes()
        assert (ur > 0).sum() > 300

"""The CUDA path (through the C ABI) against the REFERENCE ITSELF on the GPU box: oracle/_ref/libslamref.so (the reference's
ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc, ORBextractor.cc compiled from /root/reference in the build container; it travels
with the repo snapshot) runs the reference's member functions on real Frame / KeyFrame / MapPoint objects, liborbgpu.so runs the same
flat views on the device — no port in between.  Same suite as tests/test_ref_matcher.py (tests/ref_parity.py), larger sizes."""
import numpy as np
import pytest

import match_cases as mc
import oracle_lib as ol
import ref_parity as rp
from orb_slam2_with_comment_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def slamref():
    lib = ol.load_slam_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libslamref.so not shipped")
    return lib


@pytest.fixture(scope="module")
def ref(slamref):
    return lambda nnratio=0.6, checkOri=True: ol.MatcherRef(slamref, nnratio, checkOri)


@pytest.fixture(scope="module")
def gpu():
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    made = []

    def make(nnratio=0.6, checkOri=True):
        m = ORBmatcher(nnratio, checkOri)
        made.append(m)
        return m
    yield make
    for m in made:
        m.close()


def test_descriptor_distance(gpu, ref): rp.descriptor_distance(gpu, ref)
def test_search_by_projection(gpu, ref): rp.search_by_projection(gpu, ref, sizes=(5, 600, 1400, 4000))
def test_search_by_bow(gpu, ref): rp.search_by_bow(gpu, ref, sizes=(6, 300, 900))
def test_search_for_triangulation(gpu, ref): rp.search_for_triangulation(gpu, ref, sizes=(8, 800, 2000))
def test_search_windowed(gpu, ref): rp.search_windowed(gpu, ref, sizes=(6, 700, 1500, 1200))
def test_search_for_initialization(gpu, ref): rp.search_for_initialization(gpu, ref, sizes=(3, 800, 1600))
def test_fuse_candidate_loop(gpu, ref): rp.fuse_best(gpu, ref, sizes=(4, 700, 1500, 800))


def test_is_in_frustum(slamref):
    """Frame::isInFrustum + MapPoint::PredictScale on the device vs the reference's member function."""
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref, rmin = mc.frustum_case(3, raw=True)
    exp = ol.ref_is_in_frustum(slamref, cam, lsf, nl, cosl, off, P, Nn, rmin, dref)
    m = ORBmatcher()
    got = m.isInFrustum(cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref)
    assert np.array_equal(got["in_view"], exp["in_view"]) and 0.15 < exp["in_view"].mean() < 0.85
    for k in ("proj_x", "proj_y", "proj_xr", "view_cos"):
        assert got[k].tobytes() == exp[k].tobytes(), k
    # predicted level: the device uses a correctly rounded logf, the reference glibc's — they may differ by one level only where
    # log(ratio)/log(scaleFactor) sits within an ulp of an integer
    diff = np.nonzero(got["level"] != exp["level"])[0]
    assert len(diff) <= 2 and np.all(np.abs(got["level"][diff] - exp["level"][diff]) == 1), diff
    m.close()


def test_distinctive_descriptors(slamref):
    """MapPoint::ComputeDistinctiveDescriptors on the device vs the reference's member function on real observation maps."""
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    off, desc = mc.distinctive_case(11, n_points=800)
    has, best = ol.ref_distinctive_descriptors(slamref, off, desc)
    m = ORBmatcher()
    idx, _ = m.ComputeDistinctiveDescriptors(off, desc)
    assert np.array_equal(has.astype(bool), idx >= 0)
    sel = idx >= 0
    assert np.array_equal(desc[off[:-1][sel] + idx[sel]], best[sel]) and sel.sum() > 700
    m.close()


def test_stereo_matches_vs_reference_frame_constructor(slamref):
    """Frame::ComputeStereoMatches: the device (two extractors + orbgpu_stereo_matches) vs the reference's own stereo Frame
    constructor (extraction of both images on two threads + ComputeStereoMatches), bit for bit in mvuRight / mvDepth."""
    from orb_slam2_with_comment_b200 import ORBextractor
    mbf, fx = np.float32(386.1448), np.float32(718.856)
    for seed, (w, h, nf), mb in ((0, (1241, 376, 2000), mbf / fx), (1, (1241, 376, 2000), np.float32(0.05)), (2, (752, 480, 1200), mbf / fx)):
        L, R = synth.stereo_pair(w, h, seed)
        kp, desc, ur, dp = ol.ref_stereo_frame(slamref, L, R, nf, mbf, mb, fx=fx, fy=fx)
        exl = ORBextractor(nf, 1.2, 8, 20, 7, device=0, max_width=w, max_height=h, max_batch=1)
        exr = ORBextractor(nf, 1.2, 8, 20, 7, device=0, max_width=w, max_height=h, max_batch=1)
        gk, gd = exl(L)
        exr(R)
        gur, gdp = exl.stereo_matches(exr, float(mb), float(mbf))
        n = len(kp)
        assert len(gk) == n
        for f in ("x", "y", "octave", "response", "size"):
            assert np.array_equal(gk[f], kp[f]), f
        assert gur[0, :n].tobytes() == ur.tobytes() and gdp[0, :n].tobytes() == dp.tobytes()
        assert (ur > 0).sum() > 300
        exl.close(); exr.close()


def test_extraction_vs_reference_extractor(slamref):
    """ORBextractor::operator() on the device vs the reference's ORBextractor.cc directly (not via the port), on the GPU box."""
    from orb_slam2_with_comment_b200 import ORBextractor
    lib = ol.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/liborbref.so not shipped")
    for (w, h, nf), seeds in (((1241, 376, 2000), (0, 1, 2)), ((640, 480, 1000), (3, 4)), ((752, 480, 1200), (5,))):
        ex = ORBextractor(nf, 1.2, 8, 20, 7, device=0, max_width=w, max_height=h, max_batch=1)
        rex = ol.Extractor(lib, "orbref", nf, 1.2, 8, 20, 7)
        for s in seeds:
            img = synth.g_rects(w, h, s)
            kp, desc = ex(img)
            ekp, edesc = rex.extract(img)
            assert len(kp) == len(ekp)
            for f in ("x", "y", "size", "response", "octave"):
                assert np.array_equal(kp[f], ekp[f]), f
            assert np.abs(kp["angle"] - ekp["angle"]).max() <= 1e-3
            assert np.array_equal(desc, edesc)
        ex.close()

"""Parity of the CUDA Frame::ComputeStereoMatches (through the C ABI) against the CPU oracle.  The oracle is fed the GPU
extractor's own key points, descriptors and pyramids, so the comparison isolates the stereo function: mvuRight and mvDepth
must be bit-exact (the float arithmetic is the reference's, no FMA)."""
import numpy as np
import pytest

import oracle_lib as ol
from orb_slam2_with_comment_b200 import synth

pytestmark = pytest.mark.gpu

CASES = [(1241, 376, 2000, 0.537, 386.1448), (752, 480, 1200, 0.11, 47.9), (640, 480, 1000, 0.08, 40.0)]


def gpu_inputs(exL, exR, f, cntL, cntR, kpL, kpR, dL, dR):
    S = {"kpL": kpL[f, :cntL[f]], "kpR": kpR[f, :cntR[f]], "descL": dL[f, :cntL[f]], "descR": dR[f, :cntR[f]],
         "pyrL": [exL.level(l, frame=f) for l in range(8)], "pyrR": [exR.level(l, frame=f) for l in range(8)],
         "tables": np.stack([exL.GetScaleFactors(), exL.GetInverseScaleFactors(), exL.GetScaleSigmaSquares(), exL.GetInverseScaleSigmaSquares()])}
    return S


@pytest.mark.parametrize("w,h,nf,mb,mbf", CASES)
def test_stereo_vs_oracle(oracle, w, h, nf, mb, mbf):
    from orb_slam2_with_comment_b200 import ORBextractor
    B = 3
    pairs = [synth.stereo_pair(w, h, 60 + i) for i in range(B)]
    left, right = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    exL = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    kpL, dL, cL = exL.extract_batch(left)
    kpR, dR, cR = exR.extract_batch(right)
    ur, dp = exL.stereo_matches(exR, mb, mbf, batch=B)
    total = 0
    for f in range(B):
        S = gpu_inputs(exL, exR, f, cL, cR, kpL, kpR, dL, dR)
        eur, edp, _ = ol.stereo_matches(oracle, S, mb, mbf)
        assert np.array_equal(ur[f, :cL[f]], eur), f"mvuRight differs at {np.nonzero(ur[f, :cL[f]] != eur)[0][:8]}"
        assert np.array_equal(dp[f, :cL[f]], edp)
        total += int((eur >= 0).sum())
    assert total > 0.3 * cL.sum()
    # idempotent; and the single-pair call equals the batched one
    ur2, dp2 = exL.stereo_matches(exR, mb, mbf, batch=B)
    assert np.array_equal(ur, ur2) and np.array_equal(dp, dp2)
    exL.close(); exR.close()


def test_stereo_edge_cases(oracle):
    from orb_slam2_with_comment_b200 import ORBextractor
    from orb_slam2_with_comment_b200.capi import OrbGpuError
    w, h, nf = 640, 480, 1000
    exL = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
    exR = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
    with pytest.raises(OrbGpuError):
        exL.stereo_matches(exR, 0.1, 40.0)                      # nothing extracted yet
    left, _ = synth.stereo_pair(w, h, 5)
    kpL, _ = exL(left)
    exR(synth.g_flat(w, h))                                      # right image without key points: no matches at all
    ur, dp = exL.stereo_matches(exR, 0.1, 40.0)
    assert (ur[0, :len(kpL)] == -1).all() and (dp[0, :len(kpL)] == -1).all()
    # tiny maxD (mbf/mb = 3 px): every true disparity (6..14 px) is out of range
    _, right = synth.stereo_pair(w, h, 5)
    kpR, dR = exR(right)
    ur, dp = exL.stereo_matches(exR, 1.0, 3.0)
    S = gpu_inputs(exL, exR, 0, [len(kpL)], [len(kpR)], kpL[None], kpR[None], exL(left)[1][None], dR[None])
    exR(right)
    ur, dp = exL.stereo_matches(exR, 1.0, 3.0)
    eur, edp, _ = ol.stereo_matches(oracle, S, 1.0, 3.0)
    assert np.array_equal(ur[0, :len(kpL)], eur) and np.array_equal(dp[0, :len(kpL)], edp)
    exL.close(); exR.close()


def test_stereo_rejects_extractors_with_different_scale_tables():
    """The reference indexes both pyramids with ONE scale table (Frame.cc:598-616): two extractors that disagree on the scale
    factor must be refused instead of silently reading the wrong pixels."""
    from orb_slam2_with_comment_b200 import ORBextractor
    from orb_slam2_with_comment_b200.capi import OrbGpuError
    w, h, nf = 640, 480, 1000
    exL = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
    exR = ORBextractor(nf, 1.25, 8, 20, 7, max_width=w, max_height=h)
    left, right = synth.stereo_pair(w, h, 6)
    exL(left); exR(right)
    with pytest.raises(OrbGpuError):
        exL.stereo_matches(exR, 0.1, 40.0)
    exL.close(); exR.close()

"""Pins the extractor oracle three ways (the reference itself ships no tests or golden vectors):
  A. port (oracle/orb_oracle.cc)            == committed Oracle-B fixtures (cv2 primitives + Python glue)
  B. compiled reference (oracle/_ref)       == the same fixtures
  C. port == compiled reference on further seeded frames, stage by stage (pyramid, keypoints, descriptors)
"""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol
from orb_slam2_with_comment_b200 import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "extractor_golden.npz"))
CASES = [str(c) for c in G["cases"]]


def _case(name):
    w, h, seed, nf = [int(v) for v in G[name + "_meta"]]
    img = getattr(synth, str(G[name + "_gen"]))(w, h, seed)
    assert hashlib.sha1(img.tobytes()).hexdigest() == str(G[name + "_sha1"]), "synthetic generator drifted"
    return img, nf, G[name + "_kp"], G[name + "_desc"]


@pytest.mark.parametrize("name", CASES)
def test_port_matches_cv2_fixture(oracle, name):
    img, nf, kp, desc = _case(name)
    k, d = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7).extract(img)
    assert len(k) == len(kp)
    assert k.tobytes() == kp.tobytes()
    assert np.array_equal(d, desc)


@pytest.mark.parametrize("name", CASES)
def test_compiled_reference_matches_cv2_fixture(refso, name):
    img, nf, kp, desc = _case(name)
    k, d = ol.Extractor(refso, "orbref", nf, 1.2, 8, 20, 7).extract(img)
    assert k.tobytes() == kp.tobytes()
    assert np.array_equal(d, desc)


@pytest.mark.parametrize("shape", [(1241, 376, 2000), (640, 480, 1000), (752, 480, 1200)])
def test_port_equals_compiled_reference(oracle, refso, shape):
    w, h, nf = shape
    P = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7)
    R = ol.Extractor(refso, "orbref", nf, 1.2, 8, 20, 7)
    for a, b in zip(P.tables(), R.tables()):
        assert np.array_equal(a, b)
    for seed in (10, 11):
        img = synth.g_rects(w, h, seed)
        kp, dp = P.extract(img)
        kr, dr = R.extract(img)
        for l in range(8):
            assert np.array_equal(P.level(l, True), R.level(l, True)), f"level {l}"
        assert kp.tobytes() == kr.tobytes()
        assert np.array_equal(dp, dr)
        assert nf <= len(kp) <= nf + 24


def test_flat_image_gives_no_keypoints(oracle):
    k, d = ol.Extractor(oracle, "orbo", 1000, 1.2, 8, 20, 7).extract(synth.g_flat(640, 480))
    assert len(k) == 0 and d.shape == (0, 32)


def test_kitti_tables(oracle):
    s, f, u = ol.Extractor(oracle, "orbo", 2000, 1.2, 8, 20, 7).tables()
    assert list(f) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(u) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert abs(float(s[0][7]) - 3.5831816196) < 1e-6


def test_reference_nondeterminism_envelope_is_reported_and_small():
    """ORBextractor.cc:684 breaks size ties of the node sort by POINTER value; liborbref.so (the oracle) uses a stable sort instead.
    The verbatim build differs from the patched one on every frame, by about 1.2 % of the key points per side; key points both
    builds select carry identical descriptors.  The full report is profiles/r2_envelope.json (tools/envelope_report.py)."""
    import json
    import oracle_lib as ol
    from orb_slam2_with_comment_b200 import synth
    rep = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_envelope.json")))
    assert rep["total"]["frames"] >= 48 and rep["total"]["descriptor_rows_differing_on_common_keypoints"] == 0
    assert 0.0 < rep["total"]["fraction_of_keypoints_only_in_one"] < 0.04
    patched, verbatim = ol.load_ref(""), ol.load_ref("_verbatim")
    if patched is None or verbatim is None:
        pytest.skip("oracle/_ref not built")
    img = synth.g_rects(640, 480, 3)
    ka, da = ol.Extractor(patched, "orbref", 1000, 1.2, 8, 20, 7).extract(img)
    kb, db = ol.Extractor(verbatim, "orbref", 1000, 1.2, 8, 20, 7).extract(img)
    A = {(float(k["x"]), float(k["y"]), int(k["octave"])): i for i, k in enumerate(ka)}
    B = {(float(k["x"]), float(k["y"]), int(k["octave"])): i for i, k in enumerate(kb)}
    assert len(set(A) ^ set(B)) <= 0.05 * len(ka)
    assert all(np.array_equal(da[A[k]], db[B[k]]) for k in set(A) & set(B))

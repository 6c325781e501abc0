"""The stereo oracle (oracle/stereo_oracle.cc, the C++ restatement of Frame::ComputeStereoMatches, Frame.cc:501-675) against
the committed fixtures of the independent Python restatement (tests/golden/gen_stereo_golden.py) and a few properties."""
import os

import numpy as np

import oracle_lib as ol
from golden import gen_stereo_golden as gen

GOLD = os.path.join(os.path.dirname(__file__), "golden", "stereo_golden.npz")


def test_port_reproduces_stereo_fixtures(oracle):
    G = np.load(GOLD)
    assert set(str(c) for c in G["cases"]) == set(gen.CASES)
    for name in gen.CASES:
        S, mb, mbf = gen.case_inputs(name)
        ur, dp, sad = ol.stereo_matches(oracle, S, mb, mbf)
        assert np.array_equal(ur, G[name + "__u_right"]), name
        assert np.array_equal(dp, G[name + "__depth"]), name
        ok = ur >= 0
        assert ok.sum() > 100
        # the synthetic pairs have disparities of 6..14 px (integer at level 0, so up to half a level pixel off higher up)
        disp = S["kpL"]["x"][ok] - ur[ok]
        assert disp.min() > 3 and disp.max() < 17 and np.median(np.abs(disp - np.round(disp))) < 0.35
        assert np.allclose(dp[ok], mbf / disp, rtol=1e-6)


def test_stereo_without_candidates(oracle):
    S, mb, mbf = gen.case_inputs("small_a")
    S["kpR"], S["descR"] = S["kpR"][:0], S["descR"][:0]
    ur, dp, _ = ol.stereo_matches(oracle, S, mb, mbf)
    assert (ur == -1).all() and (dp == -1).all()

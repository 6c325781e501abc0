#!/usr/bin/env python3
"""Golden fixtures of the matcher path produced by THE REFERENCE ITSELF (TEST INFRASTRUCTURE): runs the suite of
tests/ref_parity.py with oracle/_ref/libslamref.so — the reference's ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc compiled
from /root/reference where they lie (oracle/Makefile, oracle/slam_ref.cc) — and records every result the reference's member
functions returned.  Needs /root/reference (build container only):

    python tests/golden/gen_reference_golden.py      # rewrites tests/golden/reference_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import match_cases as mc  # noqa: E402
import oracle_lib as ol  # noqa: E402
import ref_parity as rp  # noqa: E402


def main():
    lib = ol.load_slam_ref()
    assert lib is not None, "oracle/_ref/libslamref.so cannot be built without /root/reference"
    port = ol.load_port()
    store = {}
    for name, fn in rp.SUITE.items():
        tape = rp.TapeRef(store, name, lambda r, o: ol.MatcherRef(lib, r, o))
        fn(lambda r=0.6, o=True: ol.MatcherOracle(port, r, o), tape, sizes=rp.GOLDEN_SIZES[name])
        print(f"{name}: {tape.n} reference calls recorded")
    # Frame::isInFrustum, MapPoint::ComputeDistinctiveDescriptors and the stereo Frame constructor
    cam, lsf, nl, cosl, off, P, Nn, dmin, dmax, dref, rmin = mc.frustum_case(3, n_frames=2, n_mp=2500, raw=True)
    for k, v in ol.ref_is_in_frustum(lib, cam, lsf, nl, cosl, off, P, Nn, rmin, dref).items():
        store[f"frustum/{k}"] = v
    doff, ddesc = mc.distinctive_case(11, n_points=400)
    has, best = ol.ref_distinctive_descriptors(lib, doff, ddesc)
    store["distinctive/has"], store["distinctive/best"] = has, best
    from orb_slam2_with_comment_b200 import synth
    mbf, fx = np.float32(386.1448), np.float32(718.856)
    L, R = synth.stereo_pair(752, 480, 2)
    kp, desc, ur, dp = ol.ref_stereo_frame(lib, L, R, 1200, mbf, mbf / fx, fx=fx, fy=fx)
    store["stereo/kp"], store["stereo/desc"], store["stereo/u_right"], store["stereo/depth"] = kp, desc, ur, dp
    out = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()

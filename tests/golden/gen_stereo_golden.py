#!/usr/bin/env python3
"""Golden fixtures for Frame::ComputeStereoMatches from an INDEPENDENT Python restatement (TEST INFRASTRUCTURE).

py_stereo_matches follows /root/reference/src/Frame.cc:501-675 directly (row-band candidate lists as Python lists, float
arithmetic in numpy.float32 scalars, the 11x11 SAD with numpy on float32 patches), written separately from the C++ port in
oracle/stereo_oracle.cc.  Inputs (key points, descriptors, both pyramids) come from the extractor port, so both sides see
the same bytes.      python tests/golden/gen_stereo_golden.py   # rewrites tests/golden/stereo_golden.npz
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
from orb_slam2_with_comment_b200 import synth  # noqa: E402

f32 = np.float32
CASES = {"small_a": (320, 240, 41, 400, 0.54, 380.0), "small_b": (400, 300, 42, 500, 0.11, 40.0), "kitti_crop": (620, 188, 43, 700, 0.537, 386.1448)}


def c_round(v):
    return math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)


def py_stereo_matches(S, mb, mbf):
    kL, kR, dL, dR = S["kpL"], S["kpR"], S["descL"], S["descR"]
    sc, isc = S["tables"][0], S["tables"][1]
    N, n_rows = len(kL), S["pyrL"][0].shape[0]
    u_right, depth = np.full(N, -1, np.float32), np.full(N, -1, np.float32)
    rows = [[] for _ in range(n_rows)]
    for i in range(len(kR)):
        r = f32(2.0) * f32(sc[kR[i]["octave"]])
        lo, hi = math.floor(float(f32(kR[i]["y"]) - r)), math.ceil(float(f32(kR[i]["y"]) + r))
        for y in range(max(lo, 0), min(hi, n_rows - 1) + 1):
            rows[y].append(i)
    min_d, max_d = f32(0), f32(mbf) / f32(mb)
    accepted = []
    for i in range(N):
        lvl, uL, vL = int(kL[i]["octave"]), f32(kL[i]["x"]), f32(kL[i]["y"])
        cand = rows[int(vL)]
        if not cand:
            continue
        min_u, max_u = uL - max_d, uL - min_d
        if max_u < 0:
            continue
        best, best_r = 100, 0
        for r in cand:
            if abs(int(kR[r]["octave"]) - lvl) > 1:
                continue
            if min_u <= f32(kR[r]["x"]) <= max_u:
                d = int(np.unpackbits(dL[i] ^ dR[r]).sum())
                if d < best:
                    best, best_r = d, r
        if best >= 75:
            continue
        s = f32(isc[lvl])
        su_l, sv_l, su_r = int(c_round(float(uL * s))), int(c_round(float(vL * s))), int(c_round(float(f32(kR[best_r]["x"]) * s)))
        IL, IR = S["pyrL"][lvl].astype(np.float32), S["pyrR"][lvl].astype(np.float32)
        w = L = 5
        if su_r + L - w < 0 or su_r + L + w + 1 >= IL.shape[1]:
            continue
        pl = IL[sv_l - w:sv_l + w + 1, su_l - w:su_l + w + 1] - IL[sv_l, su_l]
        dists = []
        for inc in range(-L, L + 1):
            pr = IR[sv_l - w:sv_l + w + 1, su_r + inc - w:su_r + inc + w + 1] - IR[sv_l, su_r + inc]
            dists.append(f32(np.abs(pl - pr).astype(np.float64).sum()))
        k = int(np.argmin(dists))            # first minimum, like the strict '<' scan
        if k == 0 or k == 2 * L:
            continue
        d1, d2, d3 = dists[k - 1], dists[k], dists[k + 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = (d1 - d3) / (f32(2.0) * (d1 + d3 - f32(2.0) * d2))
        if delta < -1 or delta > 1:
            continue
        best_u = f32(sc[lvl]) * (f32(su_r) + f32(k - L) + delta)
        disp = uL - best_u
        if disp >= min_d and disp < max_d:
            if disp <= 0:
                disp, best_u = f32(0.01), f32(float(uL) - 0.01)
            depth[i], u_right[i] = f32(mbf) / disp, best_u
            accepted.append((int(dists[k]), i))
    if accepted:
        accepted.sort()
        th = f32(1.5) * f32(1.4) * f32(accepted[len(accepted) // 2][0])
        for d, i in reversed(accepted):
            if d < th:
                break
            u_right[i] = depth[i] = -1
    return u_right, depth


def case_inputs(name):
    w, h, seed, nf, mb, mbf = CASES[name]
    left, right = synth.stereo_pair(w, h, seed)
    return ol.stereo_inputs(ol.load_port(), left, right, nf), mb, mbf


def main():
    blob = {"cases": np.array(list(CASES))}
    for name in CASES:
        S, mb, mbf = case_inputs(name)
        ur, dp = py_stereo_matches(S, mb, mbf)
        blob[name + "__u_right"], blob[name + "__depth"] = ur, dp
        print(name, len(S["kpL"]), "left key points,", int((ur >= 0).sum()), "stereo matches")
    np.savez_compressed(os.path.join(HERE, "stereo_golden.npz"), **blob)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Golden fixtures for the matcher path, from an INDEPENDENT restatement of the reference (TEST INFRASTRUCTURE).

The reference ships no tests or golden vectors for ORBmatcher and ORBmatcher.cc cannot be compiled here (OpenCV,
DBoW2, Frame/KeyFrame/MapPoint), so the C++ port in oracle/match_oracle.cc is pinned by a second restatement written
separately, in plain Python, straight from the reference source:

    py_search_by_projection   ORBmatcher.cc:59-163 + Frame.cc:232-247, :353-422
    py_search_by_bow          ORBmatcher.cc:211-344 (kf_frame=True) / :635-768
    py_search_for_triangulation  ORBmatcher.cc:173-196, :783-975
    py_three_maxima           ORBmatcher.cc:1854-1895
    py_search_window_best     the candidate loops of Fuse (ORBmatcher.cc:1051-1112, :1211-1246) and SearchBySim3 (:1363-1401)
    py_search_for_initialization  ORBmatcher.cc:493-632
    py_distinctive_descriptors    MapPoint.cc:247-316
    py_is_in_frustum          Frame.cc:274-342 + MapPoint.cc:421-436 (cv::Mat algebra as cv2 4.13 evaluates it)

FeatureVectors are Python dicts (std::map<NodeId, vector<unsigned>>), the grid is a dict of lists, float arithmetic
uses numpy.float32 scalars so every product/sum rounds like the C++ float expressions.

    python tests/golden/gen_matcher_golden.py      # rewrites tests/golden/matcher_golden.npz
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import match_cases as mc  # noqa: E402

f32 = np.float32
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30


def popcount_distance(a, b) -> int:
    return sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(a, b))


def c_round(v: float) -> int:  # C round(): half away from zero
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def rot_bin(a1, a2) -> int:
    rot = f32(a1) - f32(a2)
    if rot < 0.0:
        rot = rot + f32(360.0)
    b = c_round(float(f32(rot) * f32(f32(1.0) / f32(HISTO_LENGTH))))
    return 0 if b == HISTO_LENGTH else b


def py_three_maxima(histo):
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, h in enumerate(histo):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if f32(max2) < f32(0.1) * f32(max1):
        ind2 = ind3 = -1
    elif f32(max3) < f32(0.1) * f32(max1):
        ind3 = -1
    return ind1, ind2, ind3


def fv_maps(fs, f):
    """{node id: [feature indices]} of frame f, like DBoW2::FeatureVector."""
    out = {}
    for k in range(fs.fv_node_off[f], fs.fv_node_off[f + 1]):
        out[int(fs.fv_node_id[k])] = [int(v) for v in fs.fv_feat[fs.fv_feat_off[k]:fs.fv_feat_off[k + 1]]]
    return out


def frame_view(fs, f):
    a, b = int(fs.kp_off[f]), int(fs.kp_off[f + 1])
    return (fs.keys_un[a:b], fs.desc[a:b], None if fs.kp_flags is None else fs.kp_flags[a:b],
            None if fs.u_right is None else fs.u_right[a:b])


def py_search_by_bow(fs1, fs2, idx1, idx2, nnratio, check_ori, kf_frame):
    nm_out, m_out, d_out = [], [], []
    for fa, fb in zip(idx1, idx2):
        K1, D1, FL1, _ = frame_view(fs1, fa)
        K2, D2, FL2, _ = frame_view(fs2, fb)
        v1, v2 = fv_maps(fs1, fa), fv_maps(fs2, fb)
        matches12 = [-1] * len(K1)
        dist12 = [-1] * len(K1)
        matched2 = [False] * len(K2)
        hist = [[] for _ in range(HISTO_LENGTH)]
        n = 0
        for node in sorted(set(v1) & set(v2)):      # the merge walk visits exactly the common ids, ascending
            for i1 in v1[node]:
                if FL1 is None or not (FL1[i1] & 1):
                    continue
                best1, best2, besti = 256, 256, -1
                for i2 in v2[node]:
                    if matched2[i2]:
                        continue
                    if not kf_frame and (FL2 is None or not (FL2[i2] & 1)):
                        continue
                    d = popcount_distance(D1[i1], D2[i2])
                    if d < best1:
                        best2, best1, besti = best1, d, i2
                    elif d < best2:
                        best2 = d
                ok = best1 <= TH_LOW if kf_frame else best1 < TH_LOW
                if ok and f32(best1) < f32(nnratio) * f32(best2):
                    matches12[i1], dist12[i1] = besti, best1
                    matched2[besti] = True
                    if check_ori:
                        hist[rot_bin(K1[i1]["angle"], K2[besti]["angle"])].append(i1)
                    n += 1
        if check_ori:
            keep = py_three_maxima(hist)
            for i in range(HISTO_LENGTH):
                if i in keep:
                    continue
                for i1 in hist[i]:
                    matches12[i1], dist12[i1] = -1, -1
                    n -= 1
        nm_out.append(n)
        m_out += matches12
        d_out += dist12
    return np.array(nm_out, np.int32), np.array(m_out, np.int32), np.array(d_out, np.int32)


def epipolar_ok(k1, k2, F, sigma2):
    x1, y1, x2, y2 = f32(k1["x"]), f32(k1["y"]), f32(k2["x"]), f32(k2["y"])
    a = x1 * F[0] + y1 * F[3] + F[6]
    b = x1 * F[1] + y1 * F[4] + F[7]
    c = x1 * F[2] + y1 * F[5] + F[8]
    num = a * x2 + b * y2 + c
    den = a * a + b * b
    if den == 0:
        return False
    dsqr = num * num / den
    return float(dsqr) < 3.84 * float(sigma2[int(k2["octave"])])


def py_search_for_triangulation(fs1, fs2, idx1, idx2, F12, epi, sf, s2, only_stereo, check_ori):
    nm_out, m_out, d_out = [], [], []
    for p, (fa, fb) in enumerate(zip(idx1, idx2)):
        K1, D1, FL1, U1 = frame_view(fs1, fa)
        K2, D2, FL2, U2 = frame_view(fs2, fb)
        F = [f32(v) for v in F12[p]]
        ex, ey = f32(epi[p][0]), f32(epi[p][1])
        v1, v2 = fv_maps(fs1, fa), fv_maps(fs2, fb)
        matches12 = [-1] * len(K1)
        dist12 = [-1] * len(K1)
        hist = [[] for _ in range(HISTO_LENGTH)]
        n = 0
        for node in sorted(set(v1) & set(v2)):
            for i1 in v1[node]:
                if FL1 is not None and (FL1[i1] & 1):
                    continue
                st1 = U1 is not None and U1[i1] >= 0
                if only_stereo and not st1:
                    continue
                best, besti = TH_LOW, -1
                for i2 in v2[node]:
                    if FL2 is not None and (FL2[i2] & 1):
                        continue
                    st2 = U2 is not None and U2[i2] >= 0
                    if only_stereo and not st2:
                        continue
                    d = popcount_distance(D1[i1], D2[i2])
                    if d > TH_LOW or d > best:
                        continue
                    if not st1 and not st2:
                        dx, dy = ex - f32(K2[i2]["x"]), ey - f32(K2[i2]["y"])
                        if dx * dx + dy * dy < f32(100) * f32(sf[int(K2[i2]["octave"])]):
                            continue
                    if epipolar_ok(K1[i1], K2[i2], F, s2):
                        besti, best = i2, d
                if besti >= 0:
                    matches12[i1], dist12[i1] = besti, best
                    n += 1
                    if check_ori:
                        hist[rot_bin(K1[i1]["angle"], K2[besti]["angle"])].append(i1)
        if check_ori:
            keep = py_three_maxima(hist)
            for i in range(HISTO_LENGTH):
                if i in keep:
                    continue
                for i1 in hist[i]:
                    matches12[i1], dist12[i1] = -1, -1
                    n -= 1
        nm_out.append(n)
        m_out += matches12
        d_out += dist12
    return np.array(nm_out, np.int32), np.array(m_out, np.int32), np.array(d_out, np.int32)


def py_search_by_projection(fs, mps, sf, th, nnratio):
    COLS, ROWS = 64, 48
    kp_match = np.full(int(fs.kp_off[-1]), -1, np.int32)
    bi = np.full(mps.n, -1, np.int32)
    bd = np.full(mps.n, 256, np.int32)
    sd = np.full(mps.n, 256, np.int32)
    nm_out = []
    for f in range(fs.n_frames):
        K, D, FL, UR = frame_view(fs, f)
        k0 = int(fs.kp_off[f])
        minx, miny, iw, ih = [f32(v) for v in fs.grid[f]]
        grid = {}
        for i in range(len(K)):
            px = c_round(float((f32(K[i]["x"]) - minx) * iw))
            py = c_round(float((f32(K[i]["y"]) - miny) * ih))
            if 0 <= px < COLS and 0 <= py < ROWS:
                grid.setdefault((px, py), []).append(i)
        state = [0] * len(K) if FL is None else [int(v) for v in FL]
        n = 0
        for q in range(int(mps.mp_off[f]), int(mps.mp_off[f + 1])):
            fl = int(mps.flags[q])
            if not (fl & 1) or (fl & 2):
                continue
            lvl = int(mps.level[q])
            r = f32(2.5) if float(mps.view_cos[q]) > 0.998 else f32(4.0)
            if th != 1.0:
                r = r * f32(th)
            rs = r * f32(sf[lvl])
            x, y = f32(mps.proj_x[q]), f32(mps.proj_y[q])
            cx0 = max(0, int(math.floor(float((x - minx - rs) * iw))))
            cx1 = min(COLS - 1, int(math.ceil(float((x - minx + rs) * iw))))
            cy0 = max(0, int(math.floor(float((y - miny - rs) * ih))))
            cy1 = min(ROWS - 1, int(math.ceil(float((y - miny + rs) * ih))))
            if cx0 >= COLS or cx1 < 0 or cy0 >= ROWS or cy1 < 0:
                continue
            cand = []
            for ix in range(cx0, cx1 + 1):
                for iy in range(cy0, cy1 + 1):
                    for i in grid.get((ix, iy), []):
                        o = int(K[i]["octave"])
                        if o < lvl - 1 or o > lvl:
                            continue
                        if abs(f32(K[i]["x"]) - x) < rs and abs(f32(K[i]["y"]) - y) < rs:
                            cand.append(i)
            if not cand:
                continue
            best, best2, lev, lev2, besti = 256, 256, -1, -1, -1
            for i in cand:
                if state[i] == 1:
                    continue
                if UR is not None and UR[i] > 0:
                    if abs(f32(mps.proj_xr[q]) - f32(UR[i])) > rs:
                        continue
                d = popcount_distance(mps.desc[q], D[i])
                if d < best:
                    best2, best, lev2, lev, besti = best, d, lev, int(K[i]["octave"]), i
                elif d < best2:
                    lev2, best2 = int(K[i]["octave"]), d
            bi[q], bd[q], sd[q] = besti, best, best2
            if best <= TH_HIGH:
                if lev == lev2 and f32(best) > f32(nnratio) * f32(best2):
                    continue
                state[besti] = 1 if (fl & 4) else 2
                kp_match[k0 + besti] = q - int(mps.mp_off[f])
                n += 1
        nm_out.append(n)
    return np.array(nm_out, np.int32), kp_match, bi, bd, sd


def py_search_windowed(fs, qs, th_dist, skip_any, check_ori):
    """ORBmatcher.cc:1581-1684 / :1760-1832 over projected queries."""
    COLS, ROWS = 64, 48
    kp_match = np.full(int(fs.kp_off[-1]), -1, np.int32)
    bi, bd = np.full(qs.n, -1, np.int32), np.full(qs.n, 256, np.int32)
    nm_out = []
    for f in range(fs.n_frames):
        K, D, FL, UR = frame_view(fs, f)
        k0 = int(fs.kp_off[f])
        minx, miny, iw, ih = [f32(v) for v in fs.grid[f]]
        grid = {}
        for i in range(len(K)):
            px = c_round(float((f32(K[i]["x"]) - minx) * iw))
            py = c_round(float((f32(K[i]["y"]) - miny) * ih))
            if 0 <= px < COLS and 0 <= py < ROWS:
                grid.setdefault((px, py), []).append(i)
        state = [0] * len(K) if FL is None else [int(v) for v in FL]
        hist = [[] for _ in range(HISTO_LENGTH)]
        n = 0
        for q in range(int(qs.q_off[f]), int(qs.q_off[f + 1])):
            if not (int(qs.flags[q]) & 1):
                continue
            x, y, r = f32(qs.u[q]), f32(qs.v[q]), f32(qs.radius[q])
            lo, hi = int(qs.min_level[q]), int(qs.max_level[q])
            cx0 = max(0, int(math.floor(float((x - minx - r) * iw))))
            cx1 = min(COLS - 1, int(math.ceil(float((x - minx + r) * iw))))
            cy0 = max(0, int(math.floor(float((y - miny - r) * ih))))
            cy1 = min(ROWS - 1, int(math.ceil(float((y - miny + r) * ih))))
            if cx0 >= COLS or cx1 < 0 or cy0 >= ROWS or cy1 < 0:
                continue
            check_levels = lo > 0 or hi >= 0
            best, besti = 256, -1
            for ix in range(cx0, cx1 + 1):
                for iy in range(cy0, cy1 + 1):
                    for i in grid.get((ix, iy), []):
                        o = int(K[i]["octave"])
                        if check_levels and (o < lo or (hi >= 0 and o > hi)):
                            continue
                        if not (abs(f32(K[i]["x"]) - x) < r and abs(f32(K[i]["y"]) - y) < r):
                            continue
                        if state[i] == 1 or (skip_any and state[i] != 0):
                            continue
                        if qs.ur is not None and UR is not None and UR[i] > 0 and abs(f32(qs.ur[q]) - f32(UR[i])) > r:
                            continue
                        d = popcount_distance(qs.desc[q], D[i])
                        if d < best:
                            best, besti = d, i
            bi[q], bd[q] = besti, best
            if best <= th_dist:
                state[besti] = 1 if (int(qs.flags[q]) & 4) else 2
                kp_match[k0 + besti] = q - int(qs.q_off[f])
                n += 1
                if check_ori:
                    hist[rot_bin(qs.angle[q], K[besti]["angle"])].append(besti)
        if check_ori:
            keep = py_three_maxima(hist)
            for b in range(HISTO_LENGTH):
                if b in keep:
                    continue
                for i in hist[b]:
                    kp_match[k0 + i] = -2
                    n -= 1
        nm_out.append(n)
    return np.array(nm_out, np.int32), kp_match, bi, bd


def _frame_grid(fs, f):
    """Frame::AssignFeaturesToGrid (Frame.cc:232-247): dict (cell x, cell y) -> key-point indices in ascending order."""
    COLS, ROWS = 64, 48
    K, D, FL, UR = frame_view(fs, f)
    minx, miny, iw, ih = [f32(v) for v in fs.grid[f]]
    grid = {}
    for i in range(len(K)):
        px = c_round(float((f32(K[i]["x"]) - minx) * iw))
        py = c_round(float((f32(K[i]["y"]) - miny) * ih))
        if 0 <= px < COLS and 0 <= py < ROWS:
            grid.setdefault((px, py), []).append(i)
    return K, D, FL, UR, (minx, miny, iw, ih), grid


def _features_in_area(K, g, grid, x, y, r, lo, hi):
    """GetFeaturesInArea (Frame.cc:353-410 / KeyFrame.cc:583-622) in its own order: cell columns, cell rows, index."""
    COLS, ROWS = 64, 48
    minx, miny, iw, ih = g
    cx0 = max(0, int(math.floor(float((x - minx - r) * iw))))
    cx1 = min(COLS - 1, int(math.ceil(float((x - minx + r) * iw))))
    cy0 = max(0, int(math.floor(float((y - miny - r) * ih))))
    cy1 = min(ROWS - 1, int(math.ceil(float((y - miny + r) * ih))))
    if cx0 >= COLS or cx1 < 0 or cy0 >= ROWS or cy1 < 0:
        return
    check_levels = lo > 0 or hi >= 0
    for ix in range(cx0, cx1 + 1):
        for iy in range(cy0, cy1 + 1):
            for i in grid.get((ix, iy), []):
                o = int(K[i]["octave"])
                if check_levels and (o < lo or (hi >= 0 and o > hi)):
                    continue
                if abs(f32(K[i]["x"]) - x) < r and abs(f32(K[i]["y"]) - y) < r:
                    yield i


def py_search_window_best(fs, qs, inv_s2, skip_flagged):
    bi, bd = np.full(qs.n, -1, np.int32), np.full(qs.n, 256, np.int32)
    for f in range(fs.n_frames):
        K, D, FL, UR, g, grid = _frame_grid(fs, f)
        for q in range(int(qs.q_off[f]), int(qs.q_off[f + 1])):
            if not (int(qs.flags[q]) & 1):
                continue
            u, v = f32(qs.u[q]), f32(qs.v[q])
            best, besti = 256, -1
            for i in _features_in_area(K, g, grid, u, v, f32(qs.radius[q]), int(qs.min_level[q]), int(qs.max_level[q])):
                if skip_flagged and FL is not None and int(FL[i]) != 0:
                    continue
                if inv_s2 is not None:
                    ex, ey = u - f32(K[i]["x"]), v - f32(K[i]["y"])
                    kur = f32(UR[i]) if UR is not None else f32(-1)
                    if kur >= 0:
                        er = f32(qs.ur[q]) - kur
                        e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                        if float(f32(e2 * f32(inv_s2[int(K[i]["octave"])]))) > 7.8:
                            continue
                    else:
                        e2 = f32(f32(ex * ex) + f32(ey * ey))
                        if float(f32(e2 * f32(inv_s2[int(K[i]["octave"])]))) > 5.99:
                            continue
                d = popcount_distance(qs.desc[q], D[i])
                if d < best:
                    best, besti = d, i
            bi[q], bd[q] = besti, best
    return bi, bd


def py_search_for_initialization(fs2, qs, nnratio, check_ori):
    INT_MAX = 2 ** 31 - 1
    m12_all = np.full(qs.n, -1, np.int32)
    nm_out = []
    for f in range(fs2.n_frames):
        K, D, FL, UR, g, grid = _frame_grid(fs2, f)
        q0, q1 = int(qs.q_off[f]), int(qs.q_off[f + 1])
        m12 = [-1] * (q1 - q0)
        mdist, m21 = [INT_MAX] * len(K), [-1] * len(K)
        hist = [[] for _ in range(HISTO_LENGTH)]
        n = 0
        for i1 in range(q1 - q0):
            q = q0 + i1
            if not (int(qs.flags[q]) & 1):
                continue
            best, best2, besti = INT_MAX, INT_MAX, -1
            for i2 in _features_in_area(K, g, grid, f32(qs.u[q]), f32(qs.v[q]), f32(qs.radius[q]), int(qs.min_level[q]), int(qs.max_level[q])):
                d = popcount_distance(qs.desc[q], D[i2])
                if mdist[i2] <= d:
                    continue
                if d < best:
                    best2, best, besti = best, d, i2
                elif d < best2:
                    best2 = d
            if best <= TH_LOW and f32(best) < f32(f32(best2) * f32(nnratio)):
                if m21[besti] >= 0:
                    m12[m21[besti]] = -1
                    n -= 1
                m12[i1], m21[besti], mdist[besti] = besti, i1, best
                n += 1
                if check_ori:
                    hist[rot_bin(qs.angle[q], K[besti]["angle"])].append(i1)
        if check_ori:
            keep = py_three_maxima(hist)
            for b in range(HISTO_LENGTH):
                if b in keep:
                    continue
                for i1 in hist[b]:
                    if m12[i1] >= 0:
                        m12[i1] = -1
                        n -= 1
        m12_all[q0:q1] = m12
        nm_out.append(n)
    return np.array(nm_out, np.int32), m12_all


def py_distinctive_descriptors(off, desc):
    idx, med = np.full(len(off) - 1, -1, np.int32), np.full(len(off) - 1, 2 ** 31 - 1, np.int32)
    for p in range(len(off) - 1):
        D = desc[off[p]:off[p + 1]]
        N = len(D)
        if N == 0:
            continue
        best_med, best_i = 2 ** 31 - 1, 0
        dist = [[popcount_distance(D[i], D[j]) if i != j else 0 for j in range(N)] for i in range(N)]
        for i in range(N):
            m = sorted(dist[i])[int(0.5 * (N - 1))]
            if m < best_med:
                best_med, best_i = m, i
        idx[p], med[p] = best_i, best_med
    return idx, med


def py_is_in_frustum(cam, log_sf, n_levels, cos_limit, mp_off, P, Nn, dmin, dmax, dref):
    n = int(mp_off[-1])
    out = {"in_view": np.zeros(n, np.uint8), "proj_x": np.zeros(n, f32), "proj_y": np.zeros(n, f32), "proj_xr": np.zeros(n, f32),
           "level": np.zeros(n, np.int32), "view_cos": np.zeros(n, f32)}
    for f in range(len(mp_off) - 1):
        c = cam[f].astype(f32)
        fx, fy, cx, cy, mbf, minx, maxx, miny, maxy = c[15:24]
        for q in range(int(mp_off[f]), int(mp_off[f + 1])):
            p = P[q].astype(f32)
            pc = []
            for r in range(3):   # cv::gemm small-matrix path: float dot product left to right, addend joined in double
                t = f32(f32(f32(c[3 * r] * p[0]) + f32(c[3 * r + 1] * p[1])) + f32(c[3 * r + 2] * p[2]))
                pc.append(f32(float(t) + float(c[9 + r])))
            if pc[2] < 0:
                continue
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                invz = f32(1.0) / pc[2]
                u = f32(f32(f32(fx * pc[0]) * invz) + cx)
                v = f32(f32(f32(fy * pc[1]) * invz) + cy)
            if u < minx or u > maxx or v < miny or v > maxy:
                continue
            po = [f32(p[k] - c[12 + k]) for k in range(3)]
            dist = f32(math.sqrt(sum(float(x) * float(x) for x in po)))
            if dist < dmin[q] or dist > dmax[q]:
                continue
            dot = float(po[0]) * float(Nn[q][0]) + float(po[1]) * float(Nn[q][1]) + float(po[2]) * float(Nn[q][2])
            vc = f32(dot / float(dist))
            if vc < f32(cos_limit):
                continue
            ratio = f32(f32(dref[q]) / dist)
            lvl = int(math.ceil(float(f32(f32(math.log(float(ratio))) / f32(log_sf)))))
            lvl = 0 if lvl < 0 else (n_levels - 1 if lvl >= n_levels else lvl)
            out["in_view"][q], out["proj_x"][q], out["proj_y"][q] = 1, u, v
            out["proj_xr"][q] = f32(u - f32(mbf * invz))
            out["level"][q], out["view_cos"][q] = lvl, vc
    return out


def checksum(*arrays) -> np.ndarray:
    import zlib
    c = 0
    for a in arrays:
        if a is not None:
            c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.array([c], np.uint32)


def fs_checksum(fs):
    return checksum(fs.kp_off, fs.keys_un, fs.desc, fs.kp_flags, fs.u_right, fs.grid, fs.fv_node_off, fs.fv_node_id, fs.fv_feat_off,
                    fs.fv_feat)


# case name -> (kind, generator kwargs, matcher kwargs)
CASES = {
    "bow_kfkf_nodes": ("bow", dict(seed=11, n_frames=4, n_lo=100, n_hi=160), dict(nnratio=0.75, check_ori=True, kf_frame=False)),
    "bow_kff_nodes": ("bow", dict(seed=12, n_frames=4, n_lo=100, n_hi=160), dict(nnratio=0.7, check_ori=True, kf_frame=True)),
    "bow_brute": ("bow", dict(seed=13, n_frames=3, n_lo=120, n_hi=150, single_node=True, flag_density=1.0),
                  dict(nnratio=0.75, check_ori=True, kf_frame=False)),
    "bow_brute_noori": ("bow", dict(seed=14, n_frames=3, n_lo=100, n_hi=130, single_node=True), dict(nnratio=0.9, check_ori=False, kf_frame=False)),
    "tri_mono": ("tri", dict(seed=21, n_frames=4, n_lo=120, n_hi=200), dict(only_stereo=False, check_ori=False)),
    "tri_ori_stereo": ("tri", dict(seed=22, n_frames=4, n_lo=120, n_hi=200, stereo_frac=0.5), dict(only_stereo=False, check_ori=True)),
    "tri_only_stereo": ("tri", dict(seed=23, n_frames=3, n_lo=120, n_hi=200, stereo_frac=0.6), dict(only_stereo=True, check_ori=False)),
    "sbp_mono_th1": ("sbp", dict(seed=31, n_frames=2, n_lo=250, n_hi=400, n_mp=500, th=1.0), dict(nnratio=0.8)),
    "sbp_stereo_th3": ("sbp", dict(seed=32, n_frames=2, n_lo=250, n_hi=400, n_mp=500, stereo_frac=0.5, th=3.0), dict(nnratio=0.8)),
    "sbp_wide_th15": ("sbp", dict(seed=33, n_frames=1, n_lo=300, n_hi=300, n_mp=300, th=15.0), dict(nnratio=0.9)),
    "win_frame_mono": ("win", dict(seed=41, n_frames=3, n_lo=250, n_hi=400, n_q=300, mode="frame", th=15.0), dict(th_dist=100, skip_any=False, check_ori=True)),
    "win_frame_stereo": ("win", dict(seed=42, n_frames=3, n_lo=250, n_hi=400, n_q=300, mode="frame", th=7.0, stereo_frac=0.5), dict(th_dist=100, skip_any=False, check_ori=True)),
    "win_keyframe": ("win", dict(seed=43, n_frames=2, n_lo=250, n_hi=400, n_q=300, mode="keyframe", th=10.0), dict(th_dist=64, skip_any=True, check_ori=False)),
    "best_fuse_gate_stereo": ("best", dict(seed=51, n_frames=2, n_lo=250, n_hi=400, n_q=300, mode="keyframe", th=12.0, stereo_frac=0.5), dict(gate=True, skip=False)),
    "best_fuse_gate_mono": ("best", dict(seed=52, n_frames=2, n_lo=250, n_hi=400, n_q=300, mode="keyframe", th=12.0), dict(gate=True, skip=False)),
    "best_sim3_skip": ("best", dict(seed=53, n_frames=2, n_lo=250, n_hi=400, n_q=300, mode="keyframe", th=10.0), dict(gate=False, skip=True)),
    "init_ori": ("init", dict(seed=61, n_frames=2, n_lo=300, n_hi=500), dict(nnratio=0.9, check_ori=True)),
    "init_noori": ("init", dict(seed=62, n_frames=2, n_lo=300, n_hi=500), dict(nnratio=0.9, check_ori=False)),
    "distinctive": ("distinctive", dict(seed=71, n_points=250, n_max=40), dict()),
    "frustum": ("frustum", dict(seed=81, n_frames=2, n_mp=700), dict()),
}


def best_inputs(gen):
    """win_case queries restricted to the level range [l-1, l] of Fuse / SearchBySim3, ur always present."""
    from orb_slam2_with_comment_b200 import synth
    from orb_slam2_with_comment_b200.matcher import WindowQuerySet
    fs, qs = mc.win_case(**gen)
    q2 = WindowQuerySet(qs.q_off, qs.u, qs.v, qs.radius, qs.min_level, qs.min_level + 1, qs.flags, qs.desc,
                        ur=qs.ur if qs.ur is not None else (qs.u - 5.0).astype(np.float32), angle=qs.angle)
    _, s2 = synth.scale_tables()
    return fs, q2, (np.float32(1.0) / s2).astype(np.float32)


def run_case(name):
    kind, gen, mk = CASES[name]
    out = {}
    if kind == "bow":
        s1, s2, i1, i2 = mc.bow_case(**gen)
        nm, m12, md = py_search_by_bow(s1, s2, i1, i2, mk["nnratio"], mk["check_ori"], mk["kf_frame"])
        out = {"nmatches": nm, "match12": m12, "match_dist": md, "input_crc": fs_checksum(s1)}
    elif kind == "tri":
        s1, s2, i1, i2, F12, epi, sf, s2t = mc.tri_case(**gen)
        nm, m12, md = py_search_for_triangulation(s1, s2, i1, i2, F12, epi, sf, s2t, mk["only_stereo"], mk["check_ori"])
        out = {"nmatches": nm, "match12": m12, "match_dist": md, "input_crc": fs_checksum(s1)}
    elif kind == "best":
        fs, qs, inv = best_inputs(gen)
        bi, bd = py_search_window_best(fs, qs, inv if mk["gate"] else None, mk["skip"])
        out = {"q_best_idx": bi, "q_best_dist": bd, "nmatches": np.array([(bi >= 0).sum()], np.int32), "input_crc": checksum(fs_checksum(fs), qs.u, qs.v, qs.desc)}
    elif kind == "init":
        fs2, qs = mc.init_case(**gen)
        nm, m12 = py_search_for_initialization(fs2, qs, mk["nnratio"], mk["check_ori"])
        out = {"nmatches": nm, "match12": m12, "input_crc": checksum(fs_checksum(fs2), qs.u, qs.v, qs.desc)}
    elif kind == "distinctive":
        off, desc = mc.distinctive_case(**gen)
        idx, med = py_distinctive_descriptors(off, desc)
        out = {"best_idx": idx, "best_median": med, "nmatches": np.array([(idx > 0).sum()], np.int32), "input_crc": checksum(off, desc)}
    elif kind == "frustum":
        args = mc.frustum_case(**gen)
        out = py_is_in_frustum(*args)
        out["nmatches"] = np.array([out["in_view"].sum()], np.int32)
        out["input_crc"] = checksum(*args[4:])
    elif kind == "win":
        fs, qs = mc.win_case(**gen)
        nm, kpm, bi, bd = py_search_windowed(fs, qs, mk["th_dist"], mk["skip_any"], mk["check_ori"])
        out = {"nmatches": nm, "kp_match": kpm, "q_best_idx": bi, "q_best_dist": bd, "input_crc": checksum(fs_checksum(fs), qs.u, qs.v, qs.desc)}
    else:
        fs, mps, sf, th = mc.sbp_case(**gen)
        nm, kpm, bi, bd, sd = py_search_by_projection(fs, mps, sf, th, mk["nnratio"])
        out = {"nmatches": nm, "kp_match": kpm, "mp_best_idx": bi, "mp_best_dist": bd, "mp_second_dist": sd,
               "input_crc": checksum(fs_checksum(fs), mps.proj_x, mps.proj_y, mps.desc, mps.flags)}
    return out


def main():
    blob = {"cases": np.array(list(CASES))}
    for name in CASES:
        res = run_case(name)
        for k, v in res.items():
            blob[f"{name}__{k}"] = v
        print(name, "nmatches", res["nmatches"].tolist())
    np.savez_compressed(os.path.join(HERE, "matcher_golden.npz"), **blob)


if __name__ == "__main__":
    main()

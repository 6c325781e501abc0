"""Seeded matcher workloads shared by the CPU-oracle tests, the golden-fixture generator and the GPU parity tests."""
from __future__ import annotations

import numpy as np

from orb_slam2_with_comment_b200 import synth
from orb_slam2_with_comment_b200.matcher import FrameSet, MapPointSet

W, H = 640, 480
K_TUM = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float64)


def _frames(seed, n_frames, n_lo, n_hi, p_flip=0.06, frac_rel=0.6, partner=None):
    """n_frames frames; frame f>0 re-observes a fraction of frame f-1's keypoints (flipped descriptors)."""
    rs = np.random.RandomState(seed)
    keys, descs = [], []
    for f in range(n_frames):
        n = int(rs.randint(n_lo, n_hi + 1))
        k = synth.synth_keypoints(n, W, H, seed * 1000 + f)
        d = rs.randint(0, 256, (n, 32)).astype(np.uint8)
        if f > 0 and n and len(keys[-1]):
            m = min(n, len(keys[-1]))
            take = rs.permutation(len(keys[-1]))[:int(frac_rel * m)]
            slot = rs.permutation(n)[:len(take)]
            src = keys[-1][take]
            if partner is not None:
                src = partner(src, rs)
            k[slot] = src
            k["angle"][slot] = (keys[-1]["angle"][take] + rs.normal(0, 8, len(take))).astype(np.float32) % np.float32(360)
            d[slot] = synth.flip_bits(descs[-1][take], rs, p_flip)
        keys.append(k)
        descs.append(d)
    kp_off = np.concatenate([[0], np.cumsum([len(k) for k in keys])]).astype(np.int32)
    return rs, kp_off, np.concatenate(keys), np.concatenate(descs)


def bow_case(seed, n_frames=6, n_lo=150, n_hi=400, single_node=False, flag_density=0.85, all_pairs=False):
    rs, kp_off, keys, desc = _frames(seed, n_frames, n_lo, n_hi)
    flags = (rs.uniform(size=len(keys)) < flag_density).astype(np.uint8)
    if single_node:
        fs = FrameSet.single_node(kp_off, keys, desc, kp_flags=flags)
    else:
        fv = synth.pack_feature_vectors(kp_off, desc, synth.synth_vocabulary())
        fs = FrameSet(kp_off, keys, desc, kp_flags=flags, fv_node_off=fv[0], fv_node_id=fv[1], fv_feat_off=fv[2], fv_feat=fv[3])
    if all_pairs:
        idx1, idx2 = np.meshgrid(np.arange(n_frames), np.arange(n_frames), indexing="ij")
        idx1, idx2 = idx1.ravel().astype(np.int32), idx2.ravel().astype(np.int32)
    else:
        idx1 = np.arange(1, n_frames, dtype=np.int32)
        idx2 = np.arange(0, n_frames - 1, dtype=np.int32)
    return fs, fs, idx1, idx2


def tri_case(seed, n_frames=6, n_lo=200, n_hi=500, stereo_frac=0.0, flag_density=0.3):
    R = np.eye(3)
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.5, 0.02, 0.05])
    F, ep = synth.fundamental_and_epipole(K_TUM, R, t)

    # frame f = image "1" of pair (f, f-1): its re-observed keypoints sit near the epipolar line of the partner in f-1
    def partner(src, rs):
        # src are keypoints of frame f-1 (image 2); produce image-1 points whose line passes near them: use F^T
        return synth.epipolar_partner(src, F.reshape(3, 3).T.copy().reshape(9), rs, W, H)
    rs, kp_off, keys, desc = _frames(seed, n_frames, n_lo, n_hi, partner=partner)
    flags = (rs.uniform(size=len(keys)) < flag_density).astype(np.uint8)
    u_right = np.where(rs.uniform(size=len(keys)) < stereo_frac, keys["x"] - rs.uniform(1, 30, len(keys)), -1).astype(np.float32)
    fv = synth.pack_feature_vectors(kp_off, desc, synth.synth_vocabulary())
    fs = FrameSet(kp_off, keys, desc, kp_flags=flags, u_right=u_right, fv_node_off=fv[0], fv_node_id=fv[1], fv_feat_off=fv[2],
                  fv_feat=fv[3])
    idx1 = np.arange(1, n_frames, dtype=np.int32)
    idx2 = np.arange(0, n_frames - 1, dtype=np.int32)
    F12 = np.tile(F, (len(idx1), 1))
    epi = np.tile(ep, (len(idx1), 1))
    # put the epipole of one pair inside the image so the epipole-distance rule (:889-898) fires
    if len(idx1) > 1:
        epi[1] = (W / 2, H / 2)
    sf, s2 = synth.scale_tables()
    return fs, fs, idx1, idx2, F12, epi, sf, s2


def sbp_case(seed, n_frames=4, n_lo=300, n_hi=700, n_mp=1500, stereo_frac=0.0, occupied=0.1, th=3.0):
    rs, kp_off, keys, desc = _frames(seed, n_frames + 1, n_lo, n_hi)
    # the searched frames are 1..n_frames; their local maps are built from the frame before
    ko = kp_off[1:] - kp_off[1]
    sel = slice(kp_off[1], kp_off[-1])
    fkeys, fdesc = keys[sel], desc[sel]
    flags = np.zeros(len(fkeys), np.uint8)
    r = rs.uniform(size=len(fkeys))
    flags[r < occupied] = 1           # holds a MapPoint with observations: skipped
    flags[(r >= occupied) & (r < occupied + 0.05)] = 2
    u_right = np.where(rs.uniform(size=len(fkeys)) < stereo_frac, fkeys["x"] - rs.uniform(1, 30, len(fkeys)), -1).astype(np.float32)
    grid = np.tile(synth.frame_grid(W, H), (n_frames, 1))
    fs = FrameSet(ko.astype(np.int32), fkeys, fdesc, kp_flags=flags, u_right=u_right if stereo_frac > 0 else None, grid=grid)
    parts, mp_off = [], [0]
    for f in range(n_frames):
        nm = int(n_mp * rs.uniform(0.7, 1.0))
        m = synth.local_map(keys[kp_off[f]:kp_off[f + 1]], desc[kp_off[f]:kp_off[f + 1]], nm, W, H, seed * 77 + f)
        # the re-observed keypoints of frame f+1 were copied from frame f, so projecting frame f's keypoints lands on them
        fl = m["flags"]
        u = rs.uniform(size=nm)
        fl[u < 0.05] &= ~np.uint8(1)          # not in view
        fl[(u >= 0.05) & (u < 0.08)] |= 2     # bad
        fl[(u >= 0.08) & (u < 0.2)] &= ~np.uint8(4)   # no observations: does not block its keypoint
        m["flags"] = fl
        m["view_cos"][rs.uniform(size=nm) < 0.5] = np.float32(0.9995)
        m["proj_xr"] = (m["proj_x"] - rs.uniform(1, 30, nm)).astype(np.float32)
        parts.append(m)
        mp_off.append(mp_off[-1] + nm)
    cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    mps = MapPointSet(np.array(mp_off, np.int32), cat["proj_x"], cat["proj_y"], cat["view_cos"], cat["level"], cat["flags"],
                      cat["desc"], proj_xr=cat["proj_xr"])
    sf, _ = synth.scale_tables()
    return fs, mps, sf, th


def win_case(seed, n_frames=3, n_lo=400, n_hi=800, n_q=700, stereo_frac=0.0, mode="frame", th=15.0, mbf=None):
    """Queries for the generic windowed search: the map points of a 'last frame' projected into the current frame
    (mode 'frame': level range l-1..l+1 or, for a third of the frames each, the forward [l, inf) / backward [0, l] ranges
    of ORBmatcher.cc:1604-1611; mode 'keyframe': always l-1..l+1).  With `mbf` the view is one the reference function itself
    can produce (oracle/slam_ref.cc): ur = u - mbf as at ORBmatcher.cc:1626, and no live projection left of / above the image."""
    from orb_slam2_with_comment_b200.matcher import WindowQuerySet
    rs, kp_off, keys, desc = _frames(seed, n_frames + 1, n_lo, n_hi)
    ko = kp_off[1:] - kp_off[1]
    sel = slice(kp_off[1], kp_off[-1])
    fkeys, fdesc = keys[sel], desc[sel]
    flags = np.zeros(len(fkeys), np.uint8)
    r = rs.uniform(size=len(fkeys))
    flags[r < 0.08] = 1
    flags[(r >= 0.08) & (r < 0.14)] = 2
    u_right = np.where(rs.uniform(size=len(fkeys)) < stereo_frac, fkeys["x"] - rs.uniform(1, 30, len(fkeys)), -1).astype(np.float32)
    grid = np.tile(synth.frame_grid(W, H), (n_frames, 1))
    fs = FrameSet(ko.astype(np.int32), fkeys, fdesc, kp_flags=flags, u_right=u_right if stereo_frac > 0 else None, grid=grid)
    sf, _ = synth.scale_tables()
    q_off, parts = [0], []
    for f in range(n_frames):
        prev_k, prev_d = keys[kp_off[f]:kp_off[f + 1]], desc[kp_off[f]:kp_off[f + 1]]
        nq = min(int(n_q * rs.uniform(0.7, 1.0)), len(prev_k))
        src = rs.permutation(len(prev_k))[:nq]
        k = prev_k[src]
        lvl = k["octave"].astype(np.int32)
        if mode == "frame" and f % 3 == 1:
            lo, hi = lvl, np.full(nq, -1, np.int32)          # forward
        elif mode == "frame" and f % 3 == 2:
            lo, hi = np.zeros(nq, np.int32), lvl              # backward
        else:
            lo, hi = lvl - 1, lvl + 1
        u = (k["x"] + rs.normal(0, 3, nq)).astype(np.float32)
        fl = np.full(nq, 1 | 4, np.uint8)
        z = rs.uniform(size=nq)
        fl[z < 0.05] = 0                                      # projection failed
        fl[(z >= 0.05) & (z < 0.25)] &= ~np.uint8(4)          # temporal MapPoint without observations
        v = (k["y"] + rs.normal(0, 3, nq)).astype(np.float32)
        ur = (u - rs.uniform(1, 30, nq)).astype(np.float32)
        if mbf is not None:
            ur = (u - np.float32(mbf)).astype(np.float32)
            fl[(u < 0) | (v < 0)] = 0
        parts.append({"u": u, "v": v, "radius": (np.float32(th) * sf[lvl]).astype(np.float32),
                      "lo": lo.astype(np.int32), "hi": hi.astype(np.int32), "ur": ur, "flags": fl,
                      "desc": synth.flip_bits(prev_d[src], rs, 0.05), "angle": k["angle"]})
        q_off.append(q_off[-1] + nq)
    cat = {k2: np.concatenate([p[k2] for p in parts]) for k2 in parts[0]}
    qs = WindowQuerySet(np.array(q_off, np.int32), cat["u"], cat["v"], cat["radius"], cat["lo"], cat["hi"], cat["flags"], cat["desc"],
                        ur=cat["ur"] if stereo_frac > 0 else None, angle=cat["angle"])
    return fs, qs


def frustum_case(seed, n_frames=4, n_mp=6000, raw=False):
    """Poses and local-map points for Frame::isInFrustum: a point cloud in front of (and partly behind / beside) slightly
    rotated cameras, normals roughly facing them, distance-invariance ranges that cut some of the points."""
    rs = np.random.RandomState(seed)
    cams, offs, P, Nn, dmin, dmax, dref, rmin = [], [0], [], [], [], [], [], []
    for f in range(n_frames):
        a, b, c = rs.normal(0, 0.05, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        R = (Rx @ Ry @ Rz).astype(np.float32)
        t = rs.normal(0, 0.5, 3).astype(np.float32)
        Ow = (-(R.astype(np.float64).T @ t.astype(np.float64))).astype(np.float32)
        cams.append(np.concatenate([R.ravel(), t, Ow, [517.3, 516.5, 318.6, 255.3, 40.0, 0.0, 640.0, 0.0, 480.0]]).astype(np.float32))
        n = int(n_mp * rs.uniform(0.6, 1.0))
        pc = np.stack([rs.uniform(-6, 6, n), rs.uniform(-5, 5, n), rs.uniform(-2, 25, n)], 1)          # camera coordinates
        pw = ((pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)                     # R^T (pc - t)
        d = np.linalg.norm(pw.astype(np.float64) - Ow, axis=1)
        nrm = (pw.astype(np.float64) - Ow) / np.maximum(d, 1e-6)[:, None] + rs.normal(0, 0.6, (n, 3))
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        ref = d * rs.uniform(0.3, 6.0, n)                                                               # mfMaxDistance
        P.append(pw); Nn.append(nrm.astype(np.float32))
        dref.append(ref.astype(np.float32))
        dmax.append((np.float32(1.2) * ref.astype(np.float32)).astype(np.float32))
        dmin.append((np.float32(0.8) * (ref / 1.2 ** 7).astype(np.float32)).astype(np.float32))
        rmin.append((ref / 1.2 ** 7).astype(np.float32))
        offs.append(offs[-1] + n)
    out = (np.stack(cams), np.float32(np.log(np.float32(1.2))), 8, 0.5, np.array(offs, np.int32), np.concatenate(P), np.concatenate(Nn),
           np.concatenate(dmin), np.concatenate(dmax), np.concatenate(dref))
    # raw: also mfMinDistance itself (the reference applies the 0.8 / 1.2 factors of Get{Min,Max}DistanceInvariance)
    return out + (np.concatenate(rmin),) if raw else out


def distinctive_case(seed, n_points=3000, n_max=60):
    """Observation sets of map points: a true descriptor seen from 0..n_max key frames with a few bits flipped each time,
    a share of outliers, exact duplicates (ties) and a few large sets."""
    rs = np.random.RandomState(seed)
    sizes = rs.randint(0, n_max + 1, n_points)
    sizes[:6] = [0, 1, 2, 3, 255, 256]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = np.zeros((int(off[-1]), 32), np.uint8)
    for p in range(n_points):
        n = sizes[p]
        if not n:
            continue
        base = rs.randint(0, 256, (1, 32)).astype(np.uint8)
        d = synth.flip_bits(np.repeat(base, n, 0), rs, rs.uniform(0.0, 0.12))
        out = rs.uniform(size=n) < 0.15
        d[out] = rs.randint(0, 256, (int(out.sum()), 32)).astype(np.uint8)
        if n > 3 and p % 3 == 0:
            d[n // 2] = d[0]
        desc[off[p]:off[p + 1]] = d
    return off, desc


def init_case(seed, n_frames=3, n_lo=800, n_hi=1600, window=100.0, p_flip=0.08):
    """SearchForInitialization: F2 = frame f+1, F1 = frame f; vbPrevMatched = F1's own key-point positions (as at the first call,
    Tracking.cc:612-614), windowSize 100, only level-0 key points of F1 are live.  Heavy contention: repeated descriptors make
    several F1 points want the same F2 point, so the stealing rule fires."""
    from orb_slam2_with_comment_b200.matcher import WindowQuerySet
    rs, kp_off, keys, desc = _frames(seed, n_frames + 1, n_lo, n_hi, p_flip=p_flip, frac_rel=0.7)
    ko = kp_off[1:] - kp_off[1]
    sel = slice(kp_off[1], kp_off[-1])
    fkeys, fdesc = keys[sel].copy(), desc[sel].copy()
    grid = np.tile(synth.frame_grid(W, H), (n_frames, 1))
    fs2 = FrameSet(ko.astype(np.int32), fkeys, fdesc, grid=grid)
    q_off, parts = [0], []
    for f in range(n_frames):
        k1, d1 = keys[kp_off[f]:kp_off[f + 1]], desc[kp_off[f]:kp_off[f + 1]].copy()
        n1 = len(k1)
        dup = rs.permutation(n1)[:n1 // 6]                      # near-duplicates of other F1 descriptors: rivals for one F2 point
        d1[dup] = synth.flip_bits(d1[rs.randint(0, n1, len(dup))], rs, 0.01)
        parts.append({"u": k1["x"].astype(np.float32), "v": k1["y"].astype(np.float32), "radius": np.full(n1, window, np.float32),
                      "lo": np.zeros(n1, np.int32), "hi": np.zeros(n1, np.int32), "flags": (k1["octave"] == 0).astype(np.uint8),
                      "desc": d1, "angle": k1["angle"].astype(np.float32)})
        q_off.append(q_off[-1] + n1)
    cat = {k2: np.concatenate([p[k2] for p in parts]) for k2 in parts[0]}
    qs = WindowQuerySet(np.array(q_off, np.int32), cat["u"], cat["v"], cat["radius"], cat["lo"], cat["hi"], cat["flags"], cat["desc"], angle=cat["angle"])
    return fs2, qs


def local_map_case(seed, n_frames=3, n_lo=500, n_hi=900, n_mp=1500):
    """Tracking::SearchLocalPoints in full: frames with key points, and a local map in WORLD coordinates whose points project
    (through each frame's pose) onto some of that frame's key points — so isInFrustum's outputs feed SearchByProjection."""
    rs, kp_off, keys, desc = _frames(seed, n_frames, n_lo, n_hi)
    grid = np.tile(synth.frame_grid(W, H), (n_frames, 1))
    flags_kp = np.zeros(len(keys), np.uint8)
    flags_kp[rs.uniform(size=len(keys)) < 0.1] = 1
    fs = FrameSet(kp_off, keys, desc, kp_flags=flags_kp, grid=grid)
    fx, fy, cx, cy = 517.3, 516.5, 318.6, 255.3
    cams, offs = [], [0]
    P, Nn, dmin, dmax, dref, fl, dd = [], [], [], [], [], [], []
    for f in range(n_frames):
        a, b, c = rs.normal(0, 0.04, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        R = (Rx @ Ry @ Rz).astype(np.float32)
        t = rs.normal(0, 0.4, 3).astype(np.float32)
        Ow = (-(R.astype(np.float64).T @ t.astype(np.float64))).astype(np.float32)
        cams.append(np.concatenate([R.ravel(), t, Ow, [fx, fy, cx, cy, 40.0, 0.0, float(W), 0.0, float(H)]]).astype(np.float32))
        k, d = keys[kp_off[f]:kp_off[f + 1]], desc[kp_off[f]:kp_off[f + 1]]
        n = int(n_mp * rs.uniform(0.7, 1.0))
        src = rs.randint(0, len(k), n)
        z = rs.uniform(1.0, 20.0, n)
        u = k["x"][src] + rs.normal(0, 1.5, n)
        v = k["y"][src] + rs.normal(0, 1.5, n)
        far = rs.uniform(size=n) < 0.15                       # some points nowhere near a key point / outside the image
        u[far] = rs.uniform(-100, W + 100, int(far.sum()))
        v[far] = rs.uniform(-100, H + 100, int(far.sum()))
        z[rs.uniform(size=n) < 0.05] *= -1                     # behind the camera
        pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
        pw = ((pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)
        dist = np.linalg.norm(pw.astype(np.float64) - Ow, axis=1)
        nrm = (pw.astype(np.float64) - Ow) / np.maximum(dist, 1e-6)[:, None] + rs.normal(0, 0.3, (n, 3))
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        ref = dist * 1.2 ** (k["octave"][src] - rs.uniform(0.2, 0.8, n))      # PredictScale gives the key point's octave
        ref[rs.uniform(size=n) < 0.05] *= 50.0                                  # outside the scale-invariance range
        P.append(pw); Nn.append(nrm.astype(np.float32)); dref.append(ref.astype(np.float32))
        dmax.append((np.float32(1.2) * ref.astype(np.float32)).astype(np.float32))
        dmin.append((np.float32(0.8) * (ref / 1.2 ** 7).astype(np.float32)).astype(np.float32))
        fb = np.full(n, 4, np.uint8)
        r2 = rs.uniform(size=n)
        fb[r2 < 0.04] |= 2                                     # bad
        fb[(r2 >= 0.04) & (r2 < 0.15)] &= ~np.uint8(4)         # no observations
        fl.append(fb)
        dd.append(synth.flip_bits(d[src], rs, 0.05))
        offs.append(offs[-1] + n)
    sf, _ = synth.scale_tables()
    return (fs, sf, np.stack(cams), np.float32(np.log(np.float32(1.2))), 8, 0.5, np.array(offs, np.int32), np.concatenate(P), np.concatenate(Nn),
            np.concatenate(dmin), np.concatenate(dmax), np.concatenate(dref), np.concatenate(fl), np.concatenate(dd))

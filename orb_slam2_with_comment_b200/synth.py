"""Deterministic synthetic inputs for the ORB front-end hot path (SURVEY.md §8d generators).

Pure numpy (no cv2) so the same bytes are produced in the build container and on the GPU box.
"""
from __future__ import annotations

import numpy as np


def g_rects(w: int, h: int, seed: int) -> np.ndarray:
    """G_rects: grey background, w*h/600 filled rotated rectangles drawn in order, N(0,3) noise.

    Gives a KITTI-like corner load (~13 k FAST candidates at 1241x376) and exercises the minThFAST
    fallback in a few percent of the cells.
    """
    rs = np.random.RandomState(seed)
    img = np.full((h, w), 128.0, dtype=np.float32)
    n = (w * h) // 600
    cx = rs.uniform(0, w, n)
    cy = rs.uniform(0, h, n)
    sa = rs.uniform(4, 60, n)
    sb = rs.uniform(4, 60, n)
    ang = np.deg2rad(rs.uniform(0, 180, n))
    grey = rs.randint(0, 256, n)
    for i in range(n):
        c, s = np.cos(ang[i]), np.sin(ang[i])
        ra, rb = sa[i] / 2, sb[i] / 2
        ext_x = abs(ra * c) + abs(rb * s)
        ext_y = abs(ra * s) + abs(rb * c)
        x0, x1 = max(int(np.floor(cx[i] - ext_x)), 0), min(int(np.ceil(cx[i] + ext_x)) + 1, w)
        y0, y1 = max(int(np.floor(cy[i] - ext_y)), 0), min(int(np.ceil(cy[i] + ext_y)) + 1, h)
        if x0 >= x1 or y0 >= y1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dx, dy = xx - cx[i], yy - cy[i]
        u = dx * c + dy * s
        v = -dx * s + dy * c
        m = (np.abs(u) <= ra) & (np.abs(v) <= rb)
        img[y0:y1, x0:x1][m] = grey[i]
    img += rs.normal(0, 3, (h, w)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def g_uniform(w: int, h: int, seed: int) -> np.ndarray:
    """Stress generator: i.i.d. U[0,256) pixels (~100 k FAST candidates per KITTI frame)."""
    return np.random.RandomState(seed).randint(0, 256, (h, w)).astype(np.uint8)


def g_blurnoise(w: int, h: int, seed: int) -> np.ndarray:
    """Band-limited noise: box-blurred (3 passes of 5x5) uniform noise, min-max normalised."""
    a = np.random.RandomState(seed).uniform(0, 1, (h, w))
    for _ in range(3):
        p = np.pad(a, 2, mode="reflect")
        c = np.cumsum(np.cumsum(p, 0), 1)
        c = np.pad(c, ((1, 0), (1, 0)))
        a = (c[5:, 5:] - c[:-5, 5:] - c[5:, :-5] + c[:-5, :-5]) / 25.0
    a = (a - a.min()) / max(a.max() - a.min(), 1e-12)
    return np.clip(np.rint(a * 255), 0, 255).astype(np.uint8)


def g_flat(w: int, h: int, value: int = 128) -> np.ndarray:
    return np.full((h, w), value, dtype=np.uint8)


def g_half_flat(w: int, h: int, seed: int) -> np.ndarray:
    img = g_rects(w, h, seed)
    img[:, w // 2:] = 128
    return img


def random_descriptors(n: int, seed: int) -> np.ndarray:
    return np.random.RandomState(seed).randint(0, 256, (n, 32)).astype(np.uint8)


def perturbed_descriptors(a: np.ndarray, seed: int, p_flip: float = 0.08, frac_match: float = 0.5) -> np.ndarray:
    """B = permutation of A; `frac_match` of the rows get Binomial(256,p_flip) bit flips, the rest are fresh
    random rows."""
    rs = np.random.RandomState(seed)
    n = a.shape[0]
    b = a[rs.permutation(n)].copy()
    keep = rs.uniform(size=n) < frac_match
    flips = (rs.uniform(size=(n, 256)) < p_flip)
    fl = np.packbits(flips, axis=1, bitorder="little")
    b[keep] ^= fl[keep]
    fresh = rs.randint(0, 256, (n, 32)).astype(np.uint8)
    b[~keep] = fresh[~keep]
    return b


# ---- matching workloads ---------------------------------------------------
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def scale_tables(n_levels: int = 8, scale: float = 1.2):
    """mvScaleFactor / mvLevelSigma2 as the reference constructor computes them (ORBextractor.cc:413-431)."""
    sf = np.zeros(n_levels, np.float32)
    sf[0] = 1.0
    for i in range(1, n_levels):
        sf[i] = np.float32(float(sf[i - 1]) * float(np.float32(scale)))
    return sf, (sf * sf).astype(np.float32)


def synth_keypoints(n: int, w: int, h: int, seed: int, n_levels: int = 8) -> np.ndarray:
    """Keypoints shaped like an extractor's output: level-space integer coordinates times the level scale,
    geometric level quotas, angles U[0,360)."""
    rs = np.random.RandomState(seed)
    sf, _ = scale_tables(n_levels)
    p = (1 / 1.2) ** np.arange(n_levels)
    octv = np.sort(rs.choice(n_levels, n, p=p / p.sum())).astype(np.int32)
    kp = np.zeros(n, KP_DTYPE)
    lw, lh = np.rint(w / sf[octv]), np.rint(h / sf[octv])
    kp["x"] = (np.floor(rs.uniform(19, np.maximum(lw - 19, 20))).astype(np.float32) * sf[octv]).astype(np.float32)
    kp["y"] = (np.floor(rs.uniform(19, np.maximum(lh - 19, 20))).astype(np.float32) * sf[octv]).astype(np.float32)
    kp["size"] = np.floor(31 * sf[octv])
    kp["angle"] = rs.uniform(0, 360, n).astype(np.float32)
    kp["response"] = rs.randint(7, 120, n)
    kp["octave"] = octv
    kp["class_id"] = -1
    return kp


def flip_bits(desc: np.ndarray, rs, p_flip: float) -> np.ndarray:
    fl = np.packbits(rs.uniform(size=(len(desc), 256)) < p_flip, axis=1, bitorder="little")
    return desc ^ fl


def synth_vocabulary(seed: int = 12345, k: int = 10):
    """Stand-in for the missing ORBvoc blob: a 2-level tree of k x k random 256-bit centroids."""
    rs = np.random.RandomState(seed)
    return rs.randint(0, 256, (k, 32)).astype(np.uint8), rs.randint(0, 256, (k, k, 32)).astype(np.uint8)


_POP8 = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return _POP8[a[:, None, :] ^ b[None, :, :]].sum(-1)


def feature_vector(desc: np.ndarray, vocab):
    """DBoW2-style FeatureVector of one frame: node id = argmin Hamming at each tree level (first minimum wins);
    returns (node ids ascending, per-node feature index lists in feature order)."""
    l1, l2 = vocab
    k = len(l1)
    if len(desc) == 0:
        return np.zeros(0, np.int32), []
    a = hamming_matrix(desc, l1).argmin(1)
    node = np.zeros(len(desc), np.int32)
    for c in range(k):
        m = np.nonzero(a == c)[0]
        if len(m):
            node[m] = c * k + hamming_matrix(desc[m], l2[c]).argmin(1)
    ids = np.unique(node)
    return ids.astype(np.int32), [np.nonzero(node == i)[0].astype(np.int32) for i in ids]


def pack_feature_vectors(kp_off, desc, vocab):
    """FeatureVector arrays of a frame set in the C-ABI layout (fv_node_off, fv_node_id, fv_feat_off, fv_feat)."""
    node_off, node_id, feat_off, feat = [0], [], [0], []
    for f in range(len(kp_off) - 1):
        ids, lists = feature_vector(desc[kp_off[f]:kp_off[f + 1]], vocab)
        for i, l in zip(ids, lists):
            node_id.append(int(i))
            feat.append(l)
            feat_off.append(feat_off[-1] + len(l))
        node_off.append(len(node_id))
    feat = np.concatenate(feat).astype(np.int32) if feat else np.zeros(0, np.int32)
    return (np.array(node_off, np.int32), np.array(node_id, np.int32), np.array(feat_off, np.int32), feat)


def bruteforce_sets(n_sets: int, n: int, seed: int):
    """Config #5: set A_i = n uniform random descriptors; B_i = permutation of A_i, half the rows with
    Binomial(256, 0.08) bit flips, the rest fresh random.  Returns (descA [n_sets,n,32], descB, anglesA, anglesB)."""
    A = np.stack([random_descriptors(n, seed + i) for i in range(n_sets)])
    B = np.stack([perturbed_descriptors(A[i], seed + 7919 + i) for i in range(n_sets)])
    rs = np.random.RandomState(seed + 13)
    return A, B, rs.uniform(0, 360, (n_sets, n)).astype(np.float32), rs.uniform(0, 360, (n_sets, n)).astype(np.float32)


def local_map(keys_prev: np.ndarray, desc_prev: np.ndarray, n_mp: int, w: int, h: int, seed: int, n_levels: int = 8):
    """Config #2: n_mp synthetic map points for one frame.  60 % re-observe a keypoint of the previous frame
    (descriptor with Binomial(256,0.05) flips, projection = its position + N(0,2) px, level = its octave), 40 % are
    random.  Returns dict of arrays (proj_x, proj_y, view_cos, level, flags, desc)."""
    rs = np.random.RandomState(seed)
    real = (rs.uniform(size=n_mp) < 0.6) & (len(keys_prev) > 0)
    src = rs.randint(0, max(len(keys_prev), 1), n_mp)
    desc = rs.randint(0, 256, (n_mp, 32)).astype(np.uint8)
    px = rs.uniform(0, w, n_mp).astype(np.float32)
    py = rs.uniform(0, h, n_mp).astype(np.float32)
    lvl = rs.randint(0, n_levels, n_mp).astype(np.int32)
    if real.any():
        k = keys_prev[src[real]]
        desc[real] = flip_bits(desc_prev[src[real]], rs, 0.05)
        px[real] = k["x"] + rs.normal(0, 2, real.sum()).astype(np.float32)
        py[real] = k["y"] + rs.normal(0, 2, real.sum()).astype(np.float32)
        lvl[real] = k["octave"]
    return {"proj_x": px, "proj_y": py, "view_cos": rs.uniform(0.99, 1.0, n_mp).astype(np.float32), "level": lvl,
            "flags": np.full(n_mp, 1 | 4, np.uint8), "desc": desc}


def frame_grid(w: int, h: int):
    """mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv of an undistorted w x h image (Frame.cc:96-102)."""
    return np.array([0.0, 0.0, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)], np.float32)


def skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float64)


def fundamental_and_epipole(K: np.ndarray, R12: np.ndarray, t12: np.ndarray):
    """F12 = K^-T [t12]x R12 K^-1 (LocalMapping::ComputeF12, LocalMapping.cc:676-693) in float32 row-major, and the
    epipole of camera 1 in image 2 (ORBmatcher.cc:790-799) for the same relative pose."""
    Ki = np.linalg.inv(K)
    F = Ki.T @ skew(t12) @ R12 @ Ki
    F = (F / np.abs(F).max()).astype(np.float32)
    # camera-1 centre in camera-2 coordinates: x1 = R12 x2 + t12  ->  C2 = -R12^T t12
    C2 = -R12.T @ t12
    ex = np.float32(K[0, 0] * C2[0] / C2[2] + K[0, 2])
    ey = np.float32(K[1, 1] * C2[1] / C2[2] + K[1, 2])
    return F.reshape(9), np.array([ex, ey], np.float32)


def epipolar_partner(kp1: np.ndarray, F: np.ndarray, rs, w: int, h: int, noise: float = 0.7) -> np.ndarray:
    """For every keypoint of image 1 a point of image 2 near its epipolar line x1^T F12 (so that the chi-square
    test of CheckDistEpipolarLine passes for most and fails for some)."""
    F = F.reshape(3, 3).astype(np.float64)
    x1 = np.stack([kp1["x"], kp1["y"], np.ones(len(kp1))], 1).astype(np.float64)
    l = x1 @ F   # (a, b, c)
    a, b, c = l[:, 0], l[:, 1], l[:, 2]
    nrm = np.sqrt(a * a + b * b) + 1e-30
    # foot of kp1 on the line, then slide along the line and jitter across it
    d = (a * kp1["x"] + b * kp1["y"] + c) / nrm
    fx, fy = kp1["x"] - d * a / nrm, kp1["y"] - d * b / nrm
    s = rs.uniform(-40, 40, len(kp1))
    e = rs.normal(0, noise, len(kp1)) * (1 + 2 * (rs.uniform(size=len(kp1)) < 0.15))
    out = kp1.copy()
    out["x"] = np.clip(fx - s * b / nrm + e * a / nrm, 0, w - 1).astype(np.float32)
    out["y"] = np.clip(fy + s * a / nrm + e * b / nrm, 0, h - 1).astype(np.float32)
    return out


def stereo_pair(w: int, h: int, seed: int, d0: int = 6, bands: int = 9):
    """A rectified stereo pair: left = G_rects(seed); right = left shifted left by a disparity that is constant inside each of
    `bands` horizontal bands (d0 .. d0 + bands - 1 px), so a point at uL appears at uR = uL - d in the same row."""
    left = g_rects(w, h, seed)
    right = np.empty_like(left)
    bh = (h + bands - 1) // bands
    for b in range(bands):
        y0, y1 = b * bh, min((b + 1) * bh, h)
        right[y0:y1] = np.roll(left[y0:y1], -(d0 + b), axis=1)
    return left, right


def vocabulary_tree(k: int = 10, L: int = 3, seed: int = 7, p_flip: float = 0.12, ragged: bool = False, stop_frac: float = 0.0):
    """Synthetic DBoW2 vocabulary (stand-in for the missing ORBvoc.txt) as the records of its text file, in file order
    (TemplatedVocabulary.h:1338-1423: node id = 1 + record index; ids are handed out the way HKmeansStep does — all
    children of a node first, then each child's subtree, :560-636).  A child's descriptor is its parent's with a fraction
    of bits flipped, so descents are decisive near the top and full of near-ties at the bottom.  Weights mimic idf values
    (ln(N/Ni)); `stop_frac` of the words get weight 0 (stopped words, :1185).  `ragged`: some nodes have fewer than k
    children (k-means clusters can come out empty); every leaf stays at depth L.
    Returns dict(k, L, parent[int32], is_leaf[u8], desc[n,32 u8], weight[f64])."""
    rs = np.random.RandomState(seed)
    parent, is_leaf, desc, weight = [], [], [], []

    def expand(pid, pdesc, level):
        nch = k if not ragged else int(rs.randint(max(1, k // 2), k + 1))
        ids = []
        for _ in range(nch):
            d = flip_bits(pdesc[None, :], rs, p_flip)[0] if level > 1 else rs.randint(0, 256, 32).astype(np.uint8)
            parent.append(pid)
            leaf = level == L
            is_leaf.append(1 if leaf else 0)
            desc.append(d)
            w = float(np.log(rs.uniform(1.5, 400.0))) if leaf else 0.0
            if leaf and rs.uniform() < stop_frac:
                w = 0.0
            weight.append(w)
            ids.append(len(parent))       # node id of the record just appended
        if level < L:
            for i in ids:
                expand(i, desc[i - 1], level + 1)

    expand(0, np.zeros(32, np.uint8), 1)
    return {"k": k, "L": L, "parent": np.array(parent, np.int32), "is_leaf": np.array(is_leaf, np.uint8),
            "desc": np.stack(desc).astype(np.uint8), "weight": np.array(weight, np.float64)}


def vocabulary_tree_full(k: int = 10, L: int = 6, seed: int = 7):
    """The same at ORBvoc size (k=10, L=6: 1,111,110 nodes), vectorised, level by level (breadth-first ids: a valid file
    order too, since every parent precedes its children).  A child flips each bit of its parent with probability 1/8
    (the AND of three random bytes)."""
    rs = np.random.RandomState(seed)
    parent, is_leaf, desc, weight = [], [], [], []
    prev_ids = np.zeros(1, np.int64)
    prev_desc = np.zeros((1, 32), np.uint8)
    next_id = 1
    for level in range(1, L + 1):
        n = len(prev_ids) * k
        par = np.repeat(prev_ids, k)
        if level == 1:
            d = rs.randint(0, 256, (n, 32)).astype(np.uint8)
        else:
            r = np.frombuffer(rs.bytes(3 * n * 32), np.uint8).reshape(3, n, 32)
            d = np.repeat(prev_desc, k, axis=0) ^ (r[0] & r[1] & r[2])
        parent.append(par)
        is_leaf.append(np.full(n, 1 if level == L else 0, np.uint8))
        desc.append(d)
        weight.append(np.log(rs.uniform(1.5, 400.0, n)) if level == L else np.zeros(n))
        prev_ids = np.arange(next_id, next_id + n, dtype=np.int64)
        prev_desc = d
        next_id += n
    return {"k": k, "L": L, "parent": np.concatenate(parent).astype(np.int32), "is_leaf": np.concatenate(is_leaf),
            "desc": np.concatenate(desc), "weight": np.concatenate(weight).astype(np.float64)}


def vocabulary_descriptors(voc, n: int, seed: int, p_flip: float = 0.06):
    """n descriptors near random words of the vocabulary (plus 20 % uniform random rows)."""
    rs = np.random.RandomState(seed)
    leaves = np.nonzero(voc["is_leaf"])[0]
    pick = leaves[rs.randint(0, len(leaves), n)]
    d = flip_bits(voc["desc"][pick], rs, p_flip)
    r = rs.uniform(size=n) < 0.2
    d[r] = rs.randint(0, 256, (int(r.sum()), 32)).astype(np.uint8)
    return d


def vocabulary_descriptors_fast(voc, n: int, seed: int):
    """The same for millions of rows: bit flips with probability 1/16 (the AND of four random bytes), 20 % random rows."""
    rs = np.random.RandomState(seed)
    leaves = np.nonzero(voc["is_leaf"])[0]
    pick = leaves[rs.randint(0, len(leaves), n)]
    r = np.frombuffer(rs.bytes(4 * n * 32), np.uint8).reshape(4, n, 32)
    d = voc["desc"][pick] ^ (r[0] & r[1] & r[2] & r[3])
    rnd = rs.uniform(size=n) < 0.2
    d[rnd] = np.frombuffer(rs.bytes(int(rnd.sum()) * 32), np.uint8).reshape(-1, 32)
    return d

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

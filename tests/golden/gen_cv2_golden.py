#!/usr/bin/env python3
"""Generate tests/golden/cv2_primitives.npz from the genuine OpenCV library (Python cv2 4.13.0).

The reference has no tests and its OpenCV dependency is neither vendored nor version-pinned, so these
vectors are what pins the oracle's restated primitives (oracle/cvlite.cc).  Run in the build container:
    python tests/golden/gen_cv2_golden.py
The output is committed; tests only read it.
"""
import os, sys
import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from orb_slam2_with_comment_b200.synth import g_rects, g_uniform, g_blurnoise  # noqa: E402

cv2.setNumThreads(1)
out = {"cv2_version": np.array(cv2.__version__)}

# ---- resize INTER_LINEAR (ORBextractor.cc:1120): chained x1/1.2 levels of small frames + odd ratios
cases = []
for seed, (w, h) in enumerate([(161, 97), (200, 150), (310, 94), (127, 127)]):
    img = g_rects(w, h, 100 + seed) if seed % 2 == 0 else g_uniform(w, h, 100 + seed)
    sf = np.float32(1.0)
    cur = img
    for l in range(1, 4):
        sf = np.float32(sf * 1.2000000476837158)
        inv = np.float32(1.0) / sf
        dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        nxt = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        cases.append((cur, nxt))
        cur = nxt
for (w, h, dw, dh) in [(64, 48, 50, 31), (99, 77, 98, 76), (120, 80, 61, 41), (40, 30, 57, 44)]:
    img = g_uniform(w, h, 7 * w + h)
    cases.append((img, cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)))
out["resize_n"] = np.array(len(cases))
for i, (a, b) in enumerate(cases):
    out[f"resize_src_{i}"] = a
    out[f"resize_dst_{i}"] = b

# ---- copyMakeBorder REFLECT_101, 19 px (ORBextractor.cc:1122-1128)
img = g_uniform(45, 33, 5)
out["border_src"] = img
out["border_dst"] = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)

# ---- GaussianBlur 7x7 sigma 2 REFLECT_101 (ORBextractor.cc:1086)
cases = []
for seed, (w, h) in enumerate([(161, 97), (120, 90), (64, 200), (33, 21)]):
    img = [g_rects, g_uniform, g_blurnoise, g_uniform][seed](w, h, 200 + seed)
    cases.append((img, cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)))
out["blur_n"] = np.array(len(cases))
for i, (a, b) in enumerate(cases):
    out[f"blur_src_{i}"] = a
    out[f"blur_dst_{i}"] = b

# ---- FAST-9/16 with NMS, thresholds 20 and 7 (ORBextractor.cc:809,814), cell-sized and larger images
cases = []
for seed in range(14):
    w, h = [(37, 38), (43, 46), (37, 7), (9, 40), (96, 64), (128, 100), (7, 7)][seed % 7]
    gen = [g_rects, g_blurnoise, g_uniform][seed % 3]
    img = gen(max(w, 64), max(h, 64), 300 + seed)[:h, :w].copy()
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kps = det.detect(img, None)
        arr = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], dtype=np.float32).reshape(-1, 3)
        cases.append((img, th, arr))
out["fast_n"] = np.array(len(cases))
for i, (a, th, k) in enumerate(cases):
    out[f"fast_img_{i}"] = a
    out[f"fast_th_{i}"] = np.array(th)
    out[f"fast_kps_{i}"] = k

# ---- fastAtan2 (ORBextractor.cc:103): integer-valued moments as produced by IC_Angle + edge cases
rs = np.random.RandomState(11)
yy = rs.randint(-2900000, 2900001, 20000).astype(np.float32)
xx = rs.randint(-2900000, 2900001, 20000).astype(np.float32)
yy[:8] = [0, 0, 1, -1, 0, 5, -5, 3]
xx[:8] = [0, 1, 0, 0, -1, 5, -5, -3]
out["atan_y"], out["atan_x"] = yy, xx
out["atan_out"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(yy, xx)], dtype=np.float32)

path = os.path.join(os.path.dirname(__file__), "cv2_primitives.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")

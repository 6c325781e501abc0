"""cv2 4.13 golden vectors for cvtColor(..., RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY) on 8-bit images.
Run in the build container:  python tests/golden/gen_cvtcolor_golden.py"""
import os

import cv2
import numpy as np

rs = np.random.RandomState(9)
out = {"cv_version": np.array(cv2.__version__)}
for name, ch in (("c3", 3), ("c4", 4)):
    img = rs.randint(0, 256, (2, 120, 203, ch)).astype(np.uint8)    # odd width: the byte tail of a row
    img[0, :8, :32, :3] = np.array([[r, g, b] for r in (0, 255) for g in (0, 255) for b in (0, 255)], np.uint8).repeat(32, 0).reshape(8, 32, 3)
    out[name] = img
    rgb, bgr = (cv2.COLOR_RGB2GRAY, cv2.COLOR_BGR2GRAY) if ch == 3 else (cv2.COLOR_RGBA2GRAY, cv2.COLOR_BGRA2GRAY)
    out[name + "_rgb"] = np.stack([cv2.cvtColor(f, rgb) for f in img])
    out[name + "_bgr"] = np.stack([cv2.cvtColor(f, bgr) for f in img])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cvtcolor_golden.npz"), **out)
print("ok")

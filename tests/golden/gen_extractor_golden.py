#!/usr/bin/env python3
"""Oracle B: genuine OpenCV primitives (Python cv2 4.13.0) + an independent Python restatement of the
reference's glue (cell loop, DistributeOctTree with the stable tie-break, IC_Angle, rotated BRIEF).

Writes tests/golden/extractor_golden.npz: for each case the final keypoint array (cv::KeyPoint layout)
and the N x 32 descriptor matrix.  The C++ port (oracle/orb_oracle.cc), the verbatim-compiled reference
(oracle/_ref) and the CUDA path must all reproduce these bytes.  Inputs are regenerated from seeds
(orb_slam2_with_comment_b200.synth), only a SHA-1 of each input image is stored.

Run in the build container:  python tests/golden/gen_extractor_golden.py
"""
import ctypes, hashlib, math, os, re, sys
import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from orb_slam2_with_comment_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
libm = ctypes.CDLL("libm.so.6")
libm.cosf.restype = libm.sinf.restype = ctypes.c_float
libm.cosf.argtypes = libm.sinf.argtypes = [ctypes.c_float]
f32 = np.float32

KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def load_pattern():
    inc = open(os.path.join(os.path.dirname(__file__), "..", "..", "include", "orbgpu_pattern.inc")).read()
    def arr(name):
        body = inc[inc.index(name):].split("\n#define")[0]
        body = body[len(name):]
        return np.array([int(v) for v in re.findall(r"-?\d+", body)], np.int32)
    return arr("ORB_PATTERN_X_INIT"), arr("ORB_PATTERN_Y_INIT")


PX, PY = load_pattern()
assert len(PX) == 512 and len(PY) == 512


def rnd(v):  # cvRound of a float32
    return int(np.rint(f32(v)))


class Ex:
    def __init__(self, nfeatures, scale, nlevels, ini, mn):
        self.nf, self.nl, self.ini, self.mn = nfeatures, nlevels, ini, mn
        sfd = float(f32(scale))  # double member initialised from float
        self.scale = [f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sfd))
        self.inv = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sfd)
        nd = f32(f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
        self.quota, s = [], 0
        for l in range(nlevels - 1):
            self.quota.append(rnd(nd)); s += self.quota[-1]; nd = f32(nd * factor)
        self.quota.append(max(nfeatures - s, 0))
        hp = 15
        umax = [0] * 16
        vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1)); vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
        for v in range(vmax + 1):
            umax[v] = int(np.rint(math.sqrt(hp * hp - v * v)))
        v0 = 0
        for v in range(hp, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0; v0 += 1
        self.umax = umax

    def pyramid(self, img):
        h, w = img.shape
        lv = [img]
        for l in range(1, self.nl):
            dw, dh = rnd(f32(w) * self.inv[l]), rnd(f32(h) * self.inv[l])
            lv.append(cv2.resize(lv[-1], (dw, dh), interpolation=cv2.INTER_LINEAR))
        return lv

    def cells(self, im):
        h, w = im.shape
        minBX = minBY = 16; maxBX, maxBY = w - 16, h - 16
        width, height = f32(maxBX - minBX), f32(maxBY - minBY)
        nCols, nRows = int(width / f32(30)), int(height / f32(30))
        wCell, hCell = int(math.ceil(width / f32(nCols))), int(math.ceil(height / f32(nRows)))
        out = []
        dets = {t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16) for t in (self.ini, self.mn)}
        for i in range(nRows):
            iniY = minBY + i * hCell; maxY = iniY + hCell + 6
            if iniY >= maxBY - 3: continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell; maxX = iniX + wCell + 6
                if iniX >= maxBX - 6: continue
                maxX = min(maxX, maxBX)
                sub = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
                k = dets[self.ini].detect(sub, None)
                if not k: k = dets[self.mn].detect(sub, None)
                for p in k:
                    out.append((f32(p.pt[0]) + f32(j * wCell), f32(p.pt[1]) + f32(i * hCell), f32(p.response)))
        return out, (minBX, maxBX, minBY, maxBY)

    @staticmethod
    def divide(node, K):
        (ulx, uly, urx, bry, keys) = node
        halfX = int(math.ceil(f32(urx - ulx) / f32(2))); halfY = int(math.ceil(f32(bry - uly) / f32(2)))
        sx, sy = ulx + halfX, uly + halfY
        ch = [[ulx, uly, sx, sy, []], [sx, uly, urx, sy, []], [ulx, sy, sx, bry, []], [sx, sy, urx, bry, []]]
        for k in keys:
            x, y = K[k][0], K[k][1]
            if x < sx: (ch[0] if y < sy else ch[2])[4].append(k)
            else: (ch[1] if y < sy else ch[3])[4].append(k)
        return ch

    def octree(self, K, box, N):
        minX, maxX, minY, maxY = box
        nIni = int(np.round(f32(maxX - minX) / f32(maxY - minY))) if True else 0
        # C round(): half away from zero
        q = float(f32(maxX - minX) / f32(maxY - minY)); nIni = int(math.floor(q + 0.5))
        hX = f32(maxX - minX) / f32(nIni)
        nodes = []  # list order; each node: dict
        ctr = 0
        roots = []
        for i in range(nIni):
            roots.append({"b": [int(hX * f32(i)), 0, int(hX * f32(i + 1)), maxY - minY], "k": [], "nm": False, "id": ctr}); ctr += 1
        for idx, k in enumerate(K):
            roots[int(k[0] / hX)]["k"].append(idx)
        L = []
        for r in roots:
            if len(r["k"]) == 1: r["nm"] = True; L.append(r)
            elif len(r["k"]) > 1: L.append(r)
        finish = False
        while not finish:
            prev = len(L); nexp = 0; rec = []
            newfront = []
            keep = []
            for n in L:
                if n["nm"]: keep.append(n); continue
                for c in self.divide((n["b"][0], n["b"][1], n["b"][2], n["b"][3], n["k"]), K):
                    if c[4]:
                        m = {"b": c[:4], "k": c[4], "nm": len(c[4]) == 1, "id": ctr}; ctr += 1
                        newfront.append(m)
                        if len(c[4]) > 1: nexp += 1; rec.append(m)
            L = newfront[::-1] + keep
            if len(L) >= N or len(L) == prev: finish = True
            elif len(L) + nexp * 3 > N:
                while not finish:
                    prev = len(L)
                    P = sorted(rec, key=lambda m: len(m["k"]))  # stable: ties keep creation order
                    rec = []
                    for m in P[::-1]:
                        created = []
                        for c in self.divide((m["b"][0], m["b"][1], m["b"][2], m["b"][3], m["k"]), K):
                            if c[4]:
                                mm = {"b": c[:4], "k": c[4], "nm": len(c[4]) == 1, "id": ctr}; ctr += 1
                                created.append(mm)
                                if len(c[4]) > 1: rec.append(mm)
                        L = created[::-1] + [n for n in L if n is not m]
                        if len(L) >= N: break
                    if len(L) >= N or len(L) == prev: finish = True
        res = []
        for n in L:
            best = n["k"][0]
            for k in n["k"][1:]:
                if K[k][2] > K[best][2]: best = k
            res.append(K[best])
        return res

    def angle(self, im, x, y):
        cx, cy = rnd(x), rnd(y)
        m01 = m10 = 0
        for v in range(-15, 16):
            d = self.umax[abs(v)]
            row = im[cy + v, cx - d:cx + d + 1].astype(np.int64)
            u = np.arange(-d, d + 1)
            m10 += int((u * row).sum()); m01 += v * int(row.sum())
        return f32(cv2.fastAtan2(float(m01), float(m10)))

    def descriptor(self, bl, x, y, ang):
        a_ = f32(ang) * f32(math.pi / float(f32(180.0)))
        a, b = f32(libm.cosf(float(a_))), f32(libm.sinf(float(a_)))
        cx, cy = rnd(x), rnd(y)
        px, py = PX.astype(np.float32), PY.astype(np.float32)
        yy = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
        xx = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
        vals = bl[cy + yy, cx + xx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits, bitorder="little")

    def __call__(self, img):
        lv = self.pyramid(img)
        kps, descs = [], []
        for l, im in enumerate(lv):
            K, box = self.cells(im)
            if not K: continue
            sel = self.octree(K, box, self.quota[l])
            padded = cv2.copyMakeBorder(im, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
            bl = cv2.GaussianBlur(im, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            size = f32(int(f32(31) * self.scale[l]))
            for (x, y, r) in sel:
                x, y = f32(x + f32(box[0])), f32(y + f32(box[2]))
                ang = self.angle(im, x, y)
                descs.append(self.descriptor(bl, x, y, ang))
                ox, oy = (x, y) if l == 0 else (f32(x * self.scale[l]), f32(y * self.scale[l]))
                kps.append((ox, oy, size, ang, r, l, -1))
        return np.array(kps, KP), (np.stack(descs) if descs else np.zeros((0, 32), np.uint8))


CASES = [  # name, generator, w, h, seed, nfeatures
    ("kitti_rects", "g_rects", 1241, 376, 0, 2000),
    ("tum_rects", "g_rects", 640, 480, 1, 1000),
    ("euroc_rects", "g_rects", 752, 480, 2, 1200),
    ("small_blurnoise", "g_blurnoise", 320, 240, 3, 500),
    ("small_uniform", "g_uniform", 320, 240, 4, 300),
    ("half_flat", "g_half_flat", 400, 300, 5, 800),
]

if __name__ == "__main__":
    out = {"cases": np.array([c[0] for c in CASES])}
    for name, gen, w, h, seed, nf in CASES:
        img = getattr(synth, gen)(w, h, seed)
        kp, d = Ex(nf, 1.2, 8, 20, 7)(img)
        out[name + "_meta"] = np.array([w, h, seed, nf])
        out[name + "_gen"] = np.array(gen)
        out[name + "_sha1"] = np.array(hashlib.sha1(img.tobytes()).hexdigest())
        out[name + "_kp"] = kp
        out[name + "_desc"] = d
        print(name, len(kp))
    path = os.path.join(os.path.dirname(__file__), "extractor_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))

"""ctypes bindings to the CPU oracle libraries (TEST INFRASTRUCTURE — never imported by the product).

  oracle/_build/liborboracle.so   restatement ("port"), built by `make -C oracle port`
  oracle/_ref/liborbref*.so       the reference's own ORBextractor.cc compiled against the cv:: shim
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28
CVL_KP = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])

u8p = C.POINTER(C.c_uint8)


def _p(a, t=u8p):
    return a.ctypes.data_as(t)


def build_port():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])


def load_port():
    path = os.path.join(ORACLE_DIR, "_build", "liborboracle.so")
    if not os.path.exists(path):
        build_port()
    lib = C.CDLL(path)
    lib.cvl_fast_atan2.restype = C.c_float
    lib.cvl_fast_atan2.argtypes = [C.c_float, C.c_float]
    lib.orbo_create.restype = C.c_void_p
    lib.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    lib.orbo_destroy.argtypes = [C.c_void_p]
    lib.orbo_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, u8p]
    lib.orbo_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orbo_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.orbo_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p]
    lib.orbo_get_blurred.argtypes = [C.c_void_p, C.c_int, u8p]
    lib.orbo_get_level_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.orbo_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return lib


def load_ref(kind: str = ""):
    """kind: '' (patched tie-break, parity), '_verbatim', '_fast'."""
    path = os.path.join(ORACLE_DIR, "_ref", f"liborbref{kind}.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
        else:
            return None
    lib = C.CDLL(path)
    lib.orbref_create.restype = C.c_void_p
    lib.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    lib.orbref_destroy.argtypes = [C.c_void_p]
    lib.orbref_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, u8p]
    lib.orbref_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orbref_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.orbref_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p]
    lib.orbref_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.orbref_bench.restype = C.c_double
    lib.orbref_bench.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.POINTER(C.c_long)]
    return lib


# ---- primitives -------------------------------------------------------------------------------------
def resize(lib, src, dw, dh):
    src = np.ascontiguousarray(src)
    dst = np.empty((dh, dw), np.uint8)
    lib.cvl_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def border(lib, src, b=19):
    src = np.ascontiguousarray(src)
    h, w = src.shape
    dst = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    lib.cvl_border_reflect101_u8(_p(src), w, h, w, _p(dst), w + 2 * b, b, b, b, b)
    return dst


def blur(lib, src):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib.cvl_gaussian7x7_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1])
    return dst


def fast(lib, img, th, nms=True):
    img = np.ascontiguousarray(img)
    cap = img.size // 2 + 16
    out = np.zeros(cap, CVL_KP)
    n = lib.cvl_fast9_16(_p(img), img.shape[1], img.shape[0], img.shape[1], th, int(nms), out.ctypes.data_as(C.c_void_p), cap)
    return out[:n]


def score_map(lib, img):
    img = np.ascontiguousarray(img)
    out = np.zeros_like(img)
    lib.cvl_fast_score_map(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1])
    return out


# ---- extractor objects ------------------------------------------------------------------------------
class Extractor:
    """Wraps either the port (prefix 'orbo') or the compiled reference (prefix 'orbref')."""

    def __init__(self, lib, prefix, nfeatures, scale, nlevels, ini, mn):
        self.lib, self.prefix, self.nlevels = lib, prefix, nlevels
        self.h = getattr(lib, prefix + "_create")(nfeatures, scale, nlevels, ini, mn)
        self.cap = nfeatures + 3 * nlevels + 64

    def __del__(self):
        try:
            getattr(self.lib, self.prefix + "_destroy")(self.h)
        except Exception:
            pass

    def tables(self):
        s = np.zeros(4 * self.nlevels, np.float32)
        f = np.zeros(self.nlevels, np.int32)
        u = np.zeros(16, np.int32)
        getattr(self.lib, self.prefix + "_tables")(self.h, s.ctypes.data, f.ctypes.data, u.ctypes.data)
        return s.reshape(4, self.nlevels), f, u

    def extract(self, img):
        img = np.ascontiguousarray(img)
        kp = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = getattr(self.lib, self.prefix + "_extract")(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                        kp.ctypes.data_as(C.c_void_p), self.cap, _p(desc))
        assert n <= self.cap
        return kp[:n].copy(), desc[:n].copy()

    def level(self, l, bordered=False):
        w, h = C.c_int(), C.c_int()
        getattr(self.lib, self.prefix + "_level_dims")(self.h, l, C.byref(w), C.byref(h))
        shape = (h.value + 38, w.value + 38) if bordered else (h.value, w.value)
        out = np.zeros(shape, np.uint8)
        getattr(self.lib, self.prefix + "_get_level")(self.h, l, int(bordered), _p(out))
        return out

    # port only
    def blurred(self, l):
        lv = self.level(l)
        out = np.zeros_like(lv)
        ok = self.lib.orbo_get_blurred(self.h, l, _p(out))
        return out if ok else None

    def level_points(self, l, which):
        n = self.lib.orbo_get_level_points(self.h, l, which, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.lib.orbo_get_level_points(self.h, l, which, out.ctypes.data_as(C.c_void_p), n)
        return out[:n]


def octree(lib, prefix, cand, minX, maxX, minY, maxY, N):
    cand = np.ascontiguousarray(cand)
    out = np.zeros(N + 1024, KP_DTYPE)
    n = getattr(lib, prefix + "_octree")(cand.ctypes.data_as(C.c_void_p), len(cand), minX, maxX, minY, maxY, N,
                                         out.ctypes.data_as(C.c_void_p), len(out))
    return out[:n]

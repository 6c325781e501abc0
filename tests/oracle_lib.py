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


# ---- matcher oracle (oracle/match_oracle.cc) --------------------------------------------------------------
def bind_match(lib):
    from orb_slam2_with_comment_b200.matcher import CFrameSet, CMapPointSet
    if getattr(lib, "_match_bound", False):
        return lib
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.orbm_hamming.argtypes = [u8p, u8p]
    lib.orbm_search_by_projection.restype = None
    lib.orbm_search_by_projection.argtypes = [C.POINTER(CFrameSet), C.POINTER(CMapPointSet), vp, i, f, f, vp, vp, vp, vp, vp]
    from orb_slam2_with_comment_b200.matcher import CWindowQuerySet
    lib.orbm_search_windowed.restype = None
    lib.orbm_search_windowed.argtypes = [C.POINTER(CFrameSet), C.POINTER(CWindowQuerySet), i, i, i, vp, vp, vp, vp]
    lib.orbm_search_for_triangulation.restype = None
    lib.orbm_search_for_triangulation.argtypes = [C.POINTER(CFrameSet), C.POINTER(CFrameSet), i, vp, vp, vp, vp, vp, vp, i, i, i, vp,
                                                  vp, vp, vp]
    lib.orbm_search_by_bow.restype = None
    lib.orbm_search_by_bow.argtypes = [C.POINTER(CFrameSet), C.POINTER(CFrameSet), i, vp, vp, f, i, i, i, i, vp, vp, vp, vp]
    lib.orbm_bench_bow.restype = C.c_double
    lib.orbm_bench_bow.argtypes = [C.POINTER(CFrameSet), C.POINTER(CFrameSet), i, vp, vp, f, i, i, i, i, vp, vp, vp, i]
    lib._match_bound = True
    return lib


class MatcherOracle:
    """Same method names / return dicts as orb_slam2_with_comment_b200.matcher.ORBmatcher, computed by the CPU port."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, lib, nnratio=0.6, checkOri=True):
        self.lib, self.mfNNratio, self.mbCheckOrientation = bind_match(lib), float(nnratio), bool(checkOri)

    def hamming_pairs(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        return np.array([self.lib.orbm_hamming(_p(a[i]), _p(b[i])) for i in range(len(a))], np.int32)

    def SearchByProjection(self, frames, mps, scale_factors, th=3.0):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        kp_match = np.full(int(frames.kp_off[-1]), -1, np.int32)
        bi, bd, sd = np.full(mps.n, -1, np.int32), np.full(mps.n, 256, np.int32), np.full(mps.n, 256, np.int32)
        nm = np.zeros(frames.n_frames, np.int32)
        self.lib.orbm_search_by_projection(C.byref(frames.c), C.byref(mps.c), sf.ctypes.data, len(sf), th, self.mfNNratio,
                                           kp_match.ctypes.data, bi.ctypes.data, bd.ctypes.data, sd.ctypes.data, nm.ctypes.data)
        return {"nmatches": nm, "kp_match": kp_match, "mp_best_idx": bi, "mp_best_dist": bd, "mp_second_dist": sd}

    def SearchWindowed(self, frames, queries, th_dist=100, skip_any_mappoint=False):
        kp_match = np.full(int(frames.kp_off[-1]), -1, np.int32)
        bi, bd = np.full(queries.n, -1, np.int32), np.full(queries.n, 256, np.int32)
        nm = np.zeros(frames.n_frames, np.int32)
        self.lib.orbm_search_windowed(C.byref(frames.c), C.byref(queries.c), th_dist, int(skip_any_mappoint), int(self.mbCheckOrientation),
                                      kp_match.ctypes.data, bi.ctypes.data, bd.ctypes.data, nm.ctypes.data)
        return {"nmatches": nm, "kp_match": kp_match, "q_best_idx": bi, "q_best_dist": bd}

    def SearchForInitialization(self, frames2, queries1):
        m12, nm = np.full(queries1.n, -1, np.int32), np.zeros(frames2.n_frames, np.int32)
        self.lib.orbm_search_for_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.orbm_search_for_initialization(C.byref(frames2.c), C.byref(queries1.c), self.mfNNratio, int(self.mbCheckOrientation),
                                                m12.ctypes.data, nm.ctypes.data)
        return {"match12": m12, "nmatches": nm}

    def SearchWindowBest(self, frames, queries, inv_level_sigma2=None, skip_flagged=False):
        bi, bd = np.full(queries.n, -1, np.int32), np.full(queries.n, 256, np.int32)
        s2 = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        self.lib.orbm_search_window_best.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.orbm_search_window_best(C.byref(frames.c), C.byref(queries.c), None if s2 is None else s2.ctypes.data, int(skip_flagged),
                                         bi.ctypes.data, bd.ctypes.data)
        return {"q_best_idx": bi, "q_best_dist": bd}

    def SearchForTriangulation(self, set1, set2, idx1, idx2, F12, epipole, scale_factors, level_sigma2, bOnlyStereo=False):
        from orb_slam2_with_comment_b200.matcher import match_offsets
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        F12 = np.ascontiguousarray(F12, np.float32).reshape(len(idx1), 9)
        ep = np.ascontiguousarray(epipole, np.float32).reshape(len(idx1), 2)
        sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
        off, total = match_offsets(set1, idx1)
        m12, md, nm = np.full(total, -1, np.int32), np.full(total, -1, np.int32), np.zeros(len(idx1), np.int32)
        self.lib.orbm_search_for_triangulation(C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data, idx2.ctypes.data,
                                               F12.ctypes.data, ep.ctypes.data, sf.ctypes.data, s2.ctypes.data, len(sf),
                                               int(bOnlyStereo), int(self.mbCheckOrientation), off.ctypes.data, m12.ctypes.data,
                                               md.ctypes.data, nm.ctypes.data)
        return {"nmatches": nm, "match12": m12, "match_dist": md, "match_off": off}

    def SearchByBoW(self, set1, set2, idx1, idx2, kf_frame=False):
        from orb_slam2_with_comment_b200.matcher import match_offsets
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        off, total = match_offsets(set1, idx1)
        m12, md, nm = np.full(total, -1, np.int32), np.full(total, -1, np.int32), np.zeros(len(idx1), np.int32)
        self.lib.orbm_search_by_bow(C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data, idx2.ctypes.data, self.mfNNratio,
                                    int(self.mbCheckOrientation), self.TH_LOW, int(kf_frame), int(not kf_frame), off.ctypes.data,
                                    m12.ctypes.data, md.ctypes.data, nm.ctypes.data)
        return {"nmatches": nm, "match12": m12, "match_dist": md, "match_off": off}

    def bench_bow(self, set1, set2, idx1, idx2, threads, kf_frame=False):
        from orb_slam2_with_comment_b200.matcher import match_offsets
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        off, total = match_offsets(set1, idx1)
        m12, nm = np.full(total, -1, np.int32), np.zeros(len(idx1), np.int32)
        sec = self.lib.orbm_bench_bow(C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data, idx2.ctypes.data, self.mfNNratio,
                                      int(self.mbCheckOrientation), self.TH_LOW, int(kf_frame), int(not kf_frame), off.ctypes.data,
                                      m12.ctypes.data, nm.ctypes.data, threads)
        return sec, int(nm.sum())


# ---- stereo oracle (oracle/stereo_oracle.cc) ----------------------------------------------------------------
def stereo_inputs(lib, img_l, img_r, nfeatures, nlevels=8):
    """Extracts both images with the extractor port and returns everything Frame::ComputeStereoMatches reads."""
    out = {}
    for side, img in (("L", img_l), ("R", img_r)):
        ex = Extractor(lib, "orbo", nfeatures, 1.2, nlevels, 20, 7)
        kp, desc = ex.extract(img)
        out["kp" + side], out["desc" + side] = kp, desc
        out["pyr" + side] = [np.ascontiguousarray(ex.level(l)) for l in range(nlevels)]
        out["tables"] = ex.tables()[0]
    return out


def stereo_matches(lib, S, mb, mbf):
    kpL, kpR = np.ascontiguousarray(S["kpL"]), np.ascontiguousarray(S["kpR"])
    dL, dR = np.ascontiguousarray(S["descL"]), np.ascontiguousarray(S["descR"])
    nl = len(S["pyrL"])
    PL = (C.c_void_p * nl)(*[p.ctypes.data for p in S["pyrL"]])
    PR = (C.c_void_p * nl)(*[p.ctypes.data for p in S["pyrR"]])
    lw = np.array([p.shape[1] for p in S["pyrL"]], np.int32)
    lh = np.array([p.shape[0] for p in S["pyrL"]], np.int32)
    sc, isc = np.ascontiguousarray(S["tables"][0]), np.ascontiguousarray(S["tables"][1])
    ur, dp, sad = np.zeros(len(kpL), np.float32), np.zeros(len(kpL), np.float32), np.zeros(len(kpL), np.int32)
    lib.orbs_stereo_matches.restype = None
    lib.orbs_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orbs_stereo_matches(kpL.ctypes.data, dL.ctypes.data, len(kpL), kpR.ctypes.data, dR.ctypes.data, len(kpR), PL, PR, lw.ctypes.data,
                            lh.ctypes.data, sc.ctypes.data, isc.ctypes.data, nl, mb, mbf, ur.ctypes.data, dp.ctypes.data, sad.ctypes.data)
    return ur, dp, sad


# ---- vocabulary (DBoW2 transform) -------------------------------------------------------------------
def _bind_voc(lib, prefix):
    vp, i = C.c_void_p, C.c_int
    if prefix == "orbo_voc":
        lib.orbo_voc_create.restype = vp
        lib.orbo_voc_create.argtypes = [i, i, i, i, i, vp, vp, vp, vp]
        lib.orbo_voc_free.argtypes = [vp]
        lib.orbo_voc_words.argtypes = [vp]
        lib.orbo_voc_transform.argtypes = [vp, vp, i, i, C.POINTER(i), vp, vp, C.POINTER(i), vp, vp, vp, vp, vp]
        lib.orbo_voc_bench.argtypes = [vp, vp, i, i, i, i]
    else:
        lib.dbowref_load.restype = vp
        lib.dbowref_load.argtypes = [C.c_char_p]
        lib.dbowref_free.argtypes = [vp]
        lib.dbowref_words.argtypes = [vp]
        lib.dbowref_transform.argtypes = [vp, vp, i, i, C.POINTER(i), vp, vp, C.POINTER(i), vp, vp, vp]
        lib.dbowref_score.restype = C.c_double
        lib.dbowref_score.argtypes = [vp, i, vp, vp, i, vp, vp]
    return lib


def load_dbow_ref():
    """The reference's own DBoW2 (oracle/_ref/libdbowref.so); None where it cannot be built and was not shipped."""
    path = os.path.join(ORACLE_DIR, "_ref", "libdbowref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
        else:
            return None
    return _bind_voc(C.CDLL(path), "dbowref")


class _VocBase:
    def _run(self, fn, desc, levelsup, extra):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        bw, bv = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
        fn_, fo, ff = np.zeros(max(n, 1), np.uint32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.uint32)
        nb, nf = C.c_int(0), C.c_int(0)
        vp = C.c_void_p
        fn(self.h, _p(desc, vp), n, levelsup, C.byref(nb), _p(bw, vp), _p(bv, vp), C.byref(nf), _p(fn_, vp), _p(fo, vp), _p(ff, vp), *extra)
        return {"bv_word": bw[:nb.value].copy(), "bv_value": bv[:nb.value].copy(), "fv_node_id": fn_[:nf.value].astype(np.int32),
                "fv_feat_off": fo[:nf.value + 1].copy(), "fv_feat": ff[:fo[nf.value]].astype(np.int32)}


class VocabularyOracle(_VocBase):
    """oracle/bow_oracle.cc (the port)."""

    def __init__(self, lib, rec, scoring=0, weighting=0):
        self.lib = _bind_voc(lib, "orbo_voc")
        vp = C.c_void_p
        self._keep = [np.ascontiguousarray(rec["parent"], np.int32), np.ascontiguousarray(rec["is_leaf"], np.uint8),
                      np.ascontiguousarray(rec["desc"], np.uint8), np.ascontiguousarray(rec["weight"], np.float64)]
        self.h = self.lib.orbo_voc_create(int(rec["k"]), int(rec["L"]), scoring, weighting, len(self._keep[0]), *[_p(a, vp) for a in self._keep])
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orbo_voc_free(self.h)

    def words(self):
        return self.lib.orbo_voc_words(self.h)

    def transform(self, desc, levelsup=4):
        n = len(np.asarray(desc).reshape(-1, 32))
        fw, fnode = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
        out = self._run(self.lib.orbo_voc_transform, desc, levelsup, (_p(fw, C.c_void_p), _p(fnode, C.c_void_p)))
        out["word_of_feature"], out["node_of_feature"] = fw[:n], fnode[:n]
        return out

    def bench(self, desc, n_frames, per, levelsup, threads):
        desc = np.ascontiguousarray(desc, np.uint8)
        self.lib.orbo_voc_bench(self.h, _p(desc, C.c_void_p), n_frames, per, levelsup, threads)


class VocabularyRef(_VocBase):
    """The reference's TemplatedVocabulary, loaded from a text file."""

    def __init__(self, lib, text_path):
        self.lib = lib
        self.h = lib.dbowref_load(text_path.encode())
        assert self.h, "reference DBoW2 failed to load the vocabulary text file"

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.dbowref_free(self.h)

    def words(self):
        return self.lib.dbowref_words(self.h)

    def transform(self, desc, levelsup=4):
        return self._run(self.lib.dbowref_transform, desc, levelsup, ())


# ---- Frame::isInFrustum ---------------------------------------------------------------------------------
def is_in_frustum(lib, cam, log_scale_factor, n_levels, viewing_cos_limit, mp_off, world_pos, normal, min_dist_inv, max_dist_inv, max_distance):
    vp, f, i = C.c_void_p, C.c_float, C.c_int
    lib.orbm_is_in_frustum.argtypes = [i, vp, f, i, f] + [vp] * 12
    cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 24)
    mp_off = np.ascontiguousarray(mp_off, np.int32)
    n = int(mp_off[-1])
    ins = [np.ascontiguousarray(a, np.float32) for a in (world_pos, normal, min_dist_inv, max_dist_inv, max_distance)]
    out = {"in_view": np.zeros(n, np.uint8), "proj_x": np.zeros(n, np.float32), "proj_y": np.zeros(n, np.float32),
           "proj_xr": np.zeros(n, np.float32), "level": np.zeros(n, np.int32), "view_cos": np.zeros(n, np.float32)}
    lib.orbm_is_in_frustum(len(cam), _p(cam, vp), float(log_scale_factor), n_levels, float(viewing_cos_limit), _p(mp_off, vp), *[_p(a, vp) for a in ins],
                           *[_p(out[k], vp) for k in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos")])
    return out


def distinctive_descriptors(lib, obs_off, desc):
    vp = C.c_void_p
    lib.orbm_distinctive_descriptors.argtypes = [C.c_int, vp, vp, vp, vp]
    obs_off = np.ascontiguousarray(obs_off, np.int32)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(obs_off) - 1
    idx, med = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib.orbm_distinctive_descriptors(n, _p(obs_off, vp), _p(desc, vp), _p(idx, vp), _p(med, vp))
    return idx, med


# ---- the reference's own ORBmatcher / Frame / KeyFrame / MapPoint (oracle/slam_ref.cc -> oracle/_ref/libslamref.so) -------------
def load_slam_ref():
    """The reference's ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc, ... compiled where they lie; None where it cannot be
    built (no /root/reference) and was not shipped."""
    path = os.path.join(ORACLE_DIR, "_ref", "libslamref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
        else:
            return None
    lib = C.CDLL(path)
    from orb_slam2_with_comment_b200.matcher import CFrameSet, CMapPointSet, CWindowQuerySet
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    FS, MS, QS = C.POINTER(CFrameSet), C.POINTER(CMapPointSet), C.POINTER(CWindowQuerySet)
    lib.slamref_hamming.argtypes = [u8p, u8p]
    lib.slamref_search_by_projection.restype = None
    lib.slamref_search_by_projection.argtypes = [FS, MS, vp, i, f, f, vp, vp, vp, vp, vp]
    lib.slamref_search_windowed.argtypes = [FS, QS, i, i, i, vp, vp, vp, vp, vp, i, f, f]
    lib.slamref_search_for_triangulation.restype = None
    lib.slamref_search_for_triangulation.argtypes = [FS, FS, i, vp, vp, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp]
    lib.slamref_search_by_bow.argtypes = [FS, FS, i, vp, vp, f, i, i, i, i, vp, vp, vp, vp]
    lib.slamref_search_for_initialization.argtypes = [FS, QS, f, i, vp, vp]
    lib.slamref_fuse_best.argtypes = [FS, QS, vp, vp, i, f, f, vp]
    lib.slamref_is_in_frustum.restype = None
    lib.slamref_is_in_frustum.argtypes = [i, vp, f, i, f] + [vp] * 11
    lib.slamref_distinctive_descriptors.restype = None
    lib.slamref_distinctive_descriptors.argtypes = [i, vp, vp, vp, vp]
    lib.slamref_stereo_frame.argtypes = [vp, vp, i, i, i, f, i, i, i, f, f, f, f, f, f, vp, vp, vp, vp, i]
    lib.slamref_features_in_area.argtypes = [FS, i, f, f, f, i, i, vp, i]
    lib.slamref_bench_bow.restype = C.c_double
    lib.slamref_bench_bow.argtypes = [FS, FS, i, vp, vp, f, i, i, C.POINTER(C.c_longlong)]
    return lib


class MatcherRef:
    """Same method names / return dicts as MatcherOracle, computed by the reference's own member functions on real Frame / KeyFrame /
    MapPoint objects.  Returns only what the reference function makes observable (no internal best / second distances)."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, lib, nnratio=0.6, checkOri=True):
        self.lib, self.mfNNratio, self.mbCheckOrientation = lib, float(nnratio), bool(checkOri)

    def hamming_pairs(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        return np.array([self.lib.slamref_hamming(_p(a[i]), _p(b[i])) for i in range(len(a))], np.int32)

    def SearchByProjection(self, frames, mps, scale_factors, th=3.0):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        kp_match = np.full(int(frames.kp_off[-1]), -1, np.int32)
        nm = np.zeros(frames.n_frames, np.int32)
        self.lib.slamref_search_by_projection(C.byref(frames.c), C.byref(mps.c), sf.ctypes.data, len(sf), th, self.mfNNratio,
                                              kp_match.ctypes.data, None, None, None, nm.ctypes.data)
        return {"nmatches": nm, "kp_match": kp_match}

    def SearchWindowed(self, frames, queries, scale_factors, th, mbf, th_dist=100, skip_any_mappoint=False):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        kp_match = np.full(int(frames.kp_off[-1]), -1, np.int32)
        nm = np.zeros(frames.n_frames, np.int32)
        rc = self.lib.slamref_search_windowed(C.byref(frames.c), C.byref(queries.c), th_dist, int(skip_any_mappoint), int(self.mbCheckOrientation),
                                              kp_match.ctypes.data, None, None, nm.ctypes.data, sf.ctypes.data, len(sf), th, mbf)
        assert rc == 0, f"the view is not representable by the reference function (code {rc})"
        return {"nmatches": nm, "kp_match": kp_match}

    def SearchForInitialization(self, frames2, queries1):
        m12, nm = np.full(queries1.n, -1, np.int32), np.zeros(frames2.n_frames, np.int32)
        rc = self.lib.slamref_search_for_initialization(C.byref(frames2.c), C.byref(queries1.c), self.mfNNratio, int(self.mbCheckOrientation),
                                                        m12.ctypes.data, nm.ctypes.data)
        assert rc == 0, rc
        return {"match12": m12, "nmatches": nm}

    def FuseBest(self, frames, queries, scale_factors, level_sigma2, th, mbf):
        """Winner of Fuse(KeyFrame*, vpMapPoints, th)'s candidate loop per query, -1 where the reference did not fuse (best > TH_LOW)."""
        sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
        bi = np.full(queries.n, -1, np.int32)
        rc = self.lib.slamref_fuse_best(C.byref(frames.c), C.byref(queries.c), sf.ctypes.data, s2.ctypes.data, len(sf), th, mbf, bi.ctypes.data)
        assert rc == 0, f"the view is not representable by the reference function (code {rc})"
        return {"q_best_idx": bi}

    def SearchForTriangulation(self, set1, set2, idx1, idx2, F12, epipole, scale_factors, level_sigma2, bOnlyStereo=False):
        from orb_slam2_with_comment_b200.matcher import match_offsets
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        F12 = np.ascontiguousarray(F12, np.float32).reshape(len(idx1), 9)
        ep = np.ascontiguousarray(epipole, np.float32).reshape(len(idx1), 2)
        sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
        off, total = match_offsets(set1, idx1)
        m12, nm = np.full(total, -1, np.int32), np.zeros(len(idx1), np.int32)
        self.lib.slamref_search_for_triangulation(C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data, idx2.ctypes.data,
                                                  F12.ctypes.data, ep.ctypes.data, sf.ctypes.data, s2.ctypes.data, len(sf),
                                                  int(bOnlyStereo), int(self.mbCheckOrientation), off.ctypes.data, m12.ctypes.data, None, nm.ctypes.data)
        return {"nmatches": nm, "match12": m12, "match_off": off}

    def SearchByBoW(self, set1, set2, idx1, idx2, kf_frame=False):
        from orb_slam2_with_comment_b200.matcher import match_offsets
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        off, total = match_offsets(set1, idx1)
        m12, nm = np.full(total, -1, np.int32), np.zeros(len(idx1), np.int32)
        rc = self.lib.slamref_search_by_bow(C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data, idx2.ctypes.data, self.mfNNratio,
                                            int(self.mbCheckOrientation), self.TH_LOW, int(kf_frame), int(not kf_frame), off.ctypes.data,
                                            m12.ctypes.data, None, nm.ctypes.data)
        assert rc == 0, rc
        return {"nmatches": nm, "match12": m12, "match_off": off}


def ref_bench_bow(lib, set1, set2, idx1, idx2, nnratio, check_ori, threads):
    """Seconds the reference's own ORBmatcher::SearchByBoW(KF, KF) needs for the pairs (search calls only), and the matches found."""
    idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
    n = C.c_longlong(0)
    sec = lib.slamref_bench_bow(C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data, idx2.ctypes.data, nnratio, int(check_ori), threads, C.byref(n))
    return sec, int(n.value)


def ref_is_in_frustum(lib, cam, log_scale_factor, n_levels, viewing_cos_limit, mp_off, world_pos, normal, min_distance, max_distance):
    """Frame::isInFrustum of the reference; min_distance / max_distance are the raw mfMinDistance / mfMaxDistance."""
    vp = C.c_void_p
    cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 24)
    mp_off = np.ascontiguousarray(mp_off, np.int32)
    n = int(mp_off[-1])
    ins = [np.ascontiguousarray(a, np.float32) for a in (world_pos, normal, min_distance, max_distance)]
    out = {"in_view": np.zeros(n, np.uint8), "proj_x": np.zeros(n, np.float32), "proj_y": np.zeros(n, np.float32),
           "proj_xr": np.zeros(n, np.float32), "level": np.zeros(n, np.int32), "view_cos": np.zeros(n, np.float32)}
    lib.slamref_is_in_frustum(len(cam), _p(cam, vp), float(log_scale_factor), n_levels, float(viewing_cos_limit), _p(mp_off, vp),
                              *[_p(a, vp) for a in ins], *[_p(out[k], vp) for k in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos")])
    return out


def ref_distinctive_descriptors(lib, obs_off, desc):
    """MapPoint::ComputeDistinctiveDescriptors of the reference: (has_descriptor[n], descriptor[n][32])."""
    vp = C.c_void_p
    obs_off = np.ascontiguousarray(obs_off, np.int32)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(obs_off) - 1
    best, has = np.zeros((n, 32), np.uint8), np.zeros(n, np.uint8)
    lib.slamref_distinctive_descriptors(n, _p(obs_off, vp), _p(desc, vp), _p(best, vp), _p(has, vp))
    return has, best


def ref_stereo_frame(lib, img_l, img_r, nfeatures, mbf, mb, fx=718.856, fy=718.856, cx=607.19, cy=185.2, nlevels=8):
    """The reference's stereo Frame constructor (both extractions + ComputeStereoMatches): kp, desc, mvuRight, mvDepth.  `mb` is the
    value Frame::mb holds when ComputeStereoMatches runs (the reference sets it only afterwards, Frame.cc:88 vs :113)."""
    vp = C.c_void_p
    img_l, img_r = np.ascontiguousarray(img_l), np.ascontiguousarray(img_r)
    cap = nfeatures + 3 * nlevels + 64
    kp, desc = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    n = lib.slamref_stereo_frame(_p(img_l, vp), _p(img_r, vp), img_l.shape[1], img_l.shape[0], nfeatures, 1.2, nlevels, 20, 7, fx, fy, cx, cy, mbf, mb,
                                 _p(kp, vp), _p(desc, vp), _p(ur, vp), _p(dp, vp), cap)
    assert 0 <= n <= cap
    return kp[:n].copy(), desc[:n].copy(), ur[:n].copy(), dp[:n].copy()


def ref_features_in_area(lib, frames, f, x, y, r, min_level=-1, max_level=-1):
    out = np.zeros(max(int(frames.kp_off[f + 1] - frames.kp_off[f]), 1), np.int32)
    n = lib.slamref_features_in_area(C.byref(frames.c), f, x, y, r, min_level, max_level, out.ctypes.data, len(out))
    return out[:n].copy()

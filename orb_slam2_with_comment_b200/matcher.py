"""Host-side mirror of ORB_SLAM2::ORBmatcher (/root/reference/include/ORBmatcher.h:37-102) over the C ABI.

The reference's search functions read Frame / KeyFrame / MapPoint members; here those members arrive as flat
numpy arrays (`FrameSet`, `MapPointSet` = orbgpu_frame_set / orbgpu_mappoint_set of include/orbgpu.h), batched
over many frames so that one call matches many independent frame pairs.  Same constructor arguments
(nnratio, checkOri), same constants, same function names and return values (number of matches); the match
vectors come back as integer index arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE

_f32p, _i32p, _u8p, _i64p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)


class CFrameSet(C.Structure):  # orbgpu_frame_set
    _fields_ = [("n_frames", C.c_int32), ("kp_off", _i32p), ("keys_un", C.c_void_p), ("desc", _u8p), ("u_right", _f32p),
                ("kp_flags", _u8p), ("grid", _f32p), ("fv_node_off", _i32p), ("fv_node_id", _i32p), ("fv_feat_off", _i32p),
                ("fv_feat", _i32p)]


class CMapPointSet(C.Structure):  # orbgpu_mappoint_set
    _fields_ = [("mp_off", _i32p), ("proj_x", _f32p), ("proj_y", _f32p), ("proj_xr", _f32p), ("view_cos", _f32p),
                ("level", _i32p), ("flags", _u8p), ("desc", _u8p)]


class CWindowQuerySet(C.Structure):  # orbgpu_window_query_set
    _fields_ = [("q_off", _i32p), ("u", _f32p), ("v", _f32p), ("radius", _f32p), ("min_level", _i32p), ("max_level", _i32p),
                ("ur", _f32p), ("flags", _u8p), ("desc", _u8p), ("angle", _f32p)]


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class FrameSet:
    """A batch of Frame / KeyFrame views.  kp_off[f]..kp_off[f+1] are frame f's keypoints."""

    def __init__(self, kp_off, keys_un, desc, u_right=None, kp_flags=None, grid=None, fv_node_off=None, fv_node_id=None,
                 fv_feat_off=None, fv_feat=None):
        self.kp_off = _arr(kp_off, np.int32)
        self.n_frames = len(self.kp_off) - 1
        self.keys_un = np.ascontiguousarray(keys_un, dtype=KP_DTYPE)
        self.desc = _arr(desc, np.uint8).reshape(-1, 32)
        assert len(self.keys_un) == len(self.desc) == self.kp_off[-1]
        self.u_right = _arr(u_right, np.float32)
        self.kp_flags = _arr(kp_flags, np.uint8)
        self.grid = _arr(grid, np.float32)
        self.fv_node_off, self.fv_node_id = _arr(fv_node_off, np.int32), _arr(fv_node_id, np.int32)
        self.fv_feat_off, self.fv_feat = _arr(fv_feat_off, np.int32), _arr(fv_feat, np.int32)
        self.c = CFrameSet(self.n_frames, _ptr(self.kp_off, _i32p), self.keys_un.ctypes.data, _ptr(self.desc, _u8p),
                           _ptr(self.u_right, _f32p), _ptr(self.kp_flags, _u8p), _ptr(self.grid, _f32p),
                           _ptr(self.fv_node_off, _i32p), _ptr(self.fv_node_id, _i32p), _ptr(self.fv_feat_off, _i32p),
                           _ptr(self.fv_feat, _i32p))

    def n_kp(self, f):
        return int(self.kp_off[f + 1] - self.kp_off[f])

    @staticmethod
    def single_node(kp_off, keys_un, desc, **kw):
        """Brute-force view: every frame has one node (id 0) holding all of its keypoint indices in order."""
        kp_off = np.asarray(kp_off, np.int32)
        n = len(kp_off) - 1
        feat = np.concatenate([np.arange(kp_off[f + 1] - kp_off[f], dtype=np.int32) for f in range(n)]) if n else np.zeros(0, np.int32)
        return FrameSet(kp_off, keys_un, desc, fv_node_off=np.arange(n + 1, dtype=np.int32), fv_node_id=np.zeros(n, np.int32),
                        fv_feat_off=kp_off.copy(), fv_feat=feat, **kw)


class MapPointSet:
    """The local map points handed to SearchByProjection, per frame (mp_off[f]..mp_off[f+1])."""

    def __init__(self, mp_off, proj_x, proj_y, view_cos, level, flags, desc, proj_xr=None):
        self.mp_off = _arr(mp_off, np.int32)
        self.proj_x, self.proj_y = _arr(proj_x, np.float32), _arr(proj_y, np.float32)
        self.proj_xr = _arr(proj_xr, np.float32)
        self.view_cos, self.level = _arr(view_cos, np.float32), _arr(level, np.int32)
        self.flags, self.desc = _arr(flags, np.uint8), _arr(desc, np.uint8).reshape(-1, 32)
        self.n = int(self.mp_off[-1])
        self.c = CMapPointSet(_ptr(self.mp_off, _i32p), _ptr(self.proj_x, _f32p), _ptr(self.proj_y, _f32p),
                              _ptr(self.proj_xr, _f32p), _ptr(self.view_cos, _f32p), _ptr(self.level, _i32p),
                              _ptr(self.flags, _u8p), _ptr(self.desc, _u8p))


class WindowQuerySet:
    """Projected map points for the generic windowed search (orbgpu_search_windowed), per frame."""

    def __init__(self, q_off, u, v, radius, min_level, max_level, flags, desc, ur=None, angle=None):
        self.q_off = _arr(q_off, np.int32)
        self.u, self.v, self.radius = _arr(u, np.float32), _arr(v, np.float32), _arr(radius, np.float32)
        self.min_level, self.max_level = _arr(min_level, np.int32), _arr(max_level, np.int32)
        self.ur, self.angle = _arr(ur, np.float32), _arr(angle, np.float32)
        self.flags, self.desc = _arr(flags, np.uint8), _arr(desc, np.uint8).reshape(-1, 32)
        self.n = int(self.q_off[-1])
        self.c = CWindowQuerySet(_ptr(self.q_off, _i32p), _ptr(self.u, _f32p), _ptr(self.v, _f32p), _ptr(self.radius, _f32p),
                                 _ptr(self.min_level, _i32p), _ptr(self.max_level, _i32p), _ptr(self.ur, _f32p),
                                 _ptr(self.flags, _u8p), _ptr(self.desc, _u8p), _ptr(self.angle, _f32p))


def match_offsets(fs1: FrameSet, idx1) -> np.ndarray:
    """Packed output layout: pair p writes keypoints-of-frame-idx1[p] entries at match_off[p]."""
    n = np.array([fs1.n_kp(int(f)) for f in idx1], np.int64)
    off = np.zeros(len(idx1), np.int64)
    if len(idx1) > 1:
        off[1:] = np.cumsum(n)[:-1]
    return off, int(n.sum())


def _bind(L):
    if getattr(L, "_matcher_bound", False):
        return
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.orbgpu_matcher_create.argtypes = [C.POINTER(vp), i]
    L.orbgpu_matcher_destroy.argtypes = [vp]
    L.orbgpu_matcher_sync.argtypes = [vp]
    L.orbgpu_matcher_stream.argtypes = [vp, C.POINTER(vp)]
    L.orbgpu_matcher_last_launches.argtypes = [vp]
    L.orbgpu_matcher_last_stats.argtypes = [vp, C.POINTER(f), C.POINTER(C.c_int64)]
    L.orbgpu_matcher_configure.argtypes = [vp, i, i, i]
    L.orbgpu_frame_set_upload.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(vp)]
    L.orbgpu_frame_set_release.argtypes = [vp]
    L.orbgpu_mappoint_set_upload.argtypes = [vp, C.POINTER(CMapPointSet), i, C.POINTER(vp)]
    L.orbgpu_mappoint_set_release.argtypes = [vp]
    L.orbgpu_hamming_pairs.argtypes = [vp, vp, vp, i, vp]
    L.orbgpu_search_by_projection.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(CMapPointSet), vp, i, f, f, vp, vp, vp, vp, vp]
    L.orbgpu_search_by_projection_dev.argtypes = [vp, vp, vp, vp, i, f, f, vp, vp, vp, vp, vp]
    L.orbgpu_search_windowed.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(CWindowQuerySet), i, i, i, vp, vp, vp, vp]
    L.orbgpu_search_for_triangulation.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(CFrameSet), i, vp, vp, vp, vp, vp, vp, i, i, i,
                                                  vp, vp, vp, vp]
    L.orbgpu_search_for_triangulation_dev.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp]
    L.orbgpu_search_by_bow.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(CFrameSet), i, vp, vp, f, i, i, i, i, vp, vp, vp, vp]
    L.orbgpu_search_by_bow_dev.argtypes = [vp, vp, vp, i, vp, vp, f, i, i, i, i, vp, vp, vp, vp]
    L.orbgpu_search_for_initialization.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(CWindowQuerySet), f, i, vp, vp]
    L.orbgpu_search_window_best.argtypes = [vp, C.POINTER(CFrameSet), C.POINTER(CWindowQuerySet), vp, i, i, vp, vp]
    L.orbgpu_distinctive_descriptors.argtypes = [vp, i, vp, vp, vp, vp]
    L.orbgpu_mappoint_set_project.argtypes = [vp, i, vp, f, i, f] + [vp] * 8 + [C.POINTER(vp)]
    L.orbgpu_is_in_frustum.argtypes = [vp, i, vp, f, i, f] + [vp] * 12
    L.orbgpu_is_in_frustum_dev.argtypes = [vp, i, vp, f, i, f] + [vp] * 12
    L.orbgpu_frame_set_from_extraction.argtypes = [vp, vp, vp, i, i, vp, i, vp, C.POINTER(vp)]
    L.orbgpu_frame_set_dev_info.argtypes = [vp, C.POINTER(i), vp, C.POINTER(i), C.POINTER(i)]
    L.orbgpu_frame_set_download.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L._matcher_bound = True


class ORBmatcher:
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # ORBmatcher.cc:37-39

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, *, device: int = 0):
        self._lib = capi.lib()
        _bind(self._lib)
        self.mfNNratio, self.mbCheckOrientation, self.device = float(nnratio), bool(checkOri), device
        self._h = C.c_void_p()
        capi.check(self._lib.orbgpu_matcher_create(C.byref(self._h), device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.orbgpu_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, min_queries=512, min_candidates=256, queries_per_thread=0):
        capi.check(self._lib.orbgpu_matcher_configure(self._h, min_queries, min_candidates, queries_per_thread))

    def sync(self):
        capi.check(self._lib.orbgpu_matcher_sync(self._h))

    def stream(self) -> int:
        s = C.c_void_p()
        capi.check(self._lib.orbgpu_matcher_stream(self._h, C.byref(s)))
        return s.value or 0

    def last_launches(self) -> int:
        return self._lib.orbgpu_matcher_last_launches(self._h)

    def last_stats(self):
        ms, ev = C.c_float(0), C.c_int64(0)
        capi.check(self._lib.orbgpu_matcher_last_stats(self._h, C.byref(ms), C.byref(ev)))
        return ms.value, ev.value

    # ---- DescriptorDistance (ORBmatcher.cc:1901-1917)
    def hamming_pairs(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.zeros(len(a), np.int32)
        capi.check(self._lib.orbgpu_hamming_pairs(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    def DescriptorDistance(self, a, b) -> int:
        return int(self.hamming_pairs(np.asarray(a).reshape(1, 32), np.asarray(b).reshape(1, 32))[0])

    # ---- SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:59-155), batched over frames
    def SearchByProjection(self, frames: FrameSet, mps: MapPointSet, scale_factors, th: float = 3.0):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        kp_match = np.full(int(frames.kp_off[-1]), -1, np.int32)
        best_idx = np.full(mps.n, -1, np.int32)
        best_dist = np.full(mps.n, 256, np.int32)
        second = np.full(mps.n, 256, np.int32)
        nm = np.zeros(frames.n_frames, np.int32)
        capi.check(self._lib.orbgpu_search_by_projection(self._h, C.byref(frames.c), C.byref(mps.c), sf.ctypes.data, len(sf), th,
                                                         self.mfNNratio, kp_match.ctypes.data, best_idx.ctypes.data,
                                                         best_dist.ctypes.data, second.ctypes.data, nm.ctypes.data))
        return {"nmatches": nm, "kp_match": kp_match, "mp_best_idx": best_idx, "mp_best_dist": best_dist, "mp_second_dist": second}

    # ---- the search loop of SearchByProjection(Frame&, const Frame&, th, bMono) (:1540) and (Frame&, KeyFrame*, ...) (:1711)
    def SearchWindowed(self, frames: FrameSet, queries: "WindowQuerySet", th_dist: int = 100, skip_any_mappoint: bool = False):
        kp_match = np.full(int(frames.kp_off[-1]), -1, np.int32)
        bi, bd = np.full(queries.n, -1, np.int32), np.full(queries.n, 256, np.int32)
        nm = np.zeros(frames.n_frames, np.int32)
        capi.check(self._lib.orbgpu_search_windowed(self._h, C.byref(frames.c), C.byref(queries.c), th_dist, int(skip_any_mappoint),
                                                    int(self.mbCheckOrientation), kp_match.ctypes.data, bi.ctypes.data, bd.ctypes.data,
                                                    nm.ctypes.data))
        return {"nmatches": nm, "kp_match": kp_match, "q_best_idx": bi, "q_best_dist": bd}

    # ---- SearchForTriangulation (ORBmatcher.cc:783-975), batched over keyframe pairs
    def SearchForTriangulation(self, set1: FrameSet, set2: FrameSet, idx1, idx2, F12, epipole, scale_factors, level_sigma2,
                               bOnlyStereo: bool = False):
        idx1, idx2 = _arr(idx1, np.int32), _arr(idx2, np.int32)
        F12 = np.ascontiguousarray(F12, np.float32).reshape(len(idx1), 9)
        ep = np.ascontiguousarray(epipole, np.float32).reshape(len(idx1), 2)
        sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
        off, total = match_offsets(set1, idx1)
        m12 = np.full(total, -1, np.int32)
        md = np.full(total, -1, np.int32)
        nm = np.zeros(len(idx1), np.int32)
        capi.check(self._lib.orbgpu_search_for_triangulation(self._h, C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data,
                                                             idx2.ctypes.data, F12.ctypes.data, ep.ctypes.data, sf.ctypes.data,
                                                             s2.ctypes.data, len(sf), int(bOnlyStereo), int(self.mbCheckOrientation),
                                                             off.ctypes.data, m12.ctypes.data, md.ctypes.data, nm.ctypes.data))
        return {"nmatches": nm, "match12": m12, "match_dist": md, "match_off": off}

    # ---- SearchByBoW (KeyFrame-KeyFrame :635-768 by default; kf_frame=True gives the KeyFrame-Frame rule :211-344)
    def SearchByBoW(self, set1: FrameSet, set2: FrameSet, idx1, idx2, kf_frame: bool = False):
        idx1, idx2 = _arr(idx1, np.int32), _arr(idx2, np.int32)
        off, total = match_offsets(set1, idx1)
        m12 = np.full(total, -1, np.int32)
        md = np.full(total, -1, np.int32)
        nm = np.zeros(len(idx1), np.int32)
        capi.check(self._lib.orbgpu_search_by_bow(self._h, C.byref(set1.c), C.byref(set2.c), len(idx1), idx1.ctypes.data,
                                                  idx2.ctypes.data, self.mfNNratio, int(self.mbCheckOrientation), self.TH_LOW,
                                                  int(kf_frame), int(not kf_frame), off.ctypes.data, m12.ctypes.data,
                                                  md.ctypes.data, nm.ctypes.data))
        return {"nmatches": nm, "match12": m12, "match_dist": md, "match_off": off}

    # ---- device-resident handles (bench path)
    def isInFrustum(self, cam, log_scale_factor, n_levels, viewing_cos_limit, mp_off, world_pos, normal, min_dist_inv, max_dist_inv,
                    max_distance):
        """Frame::isInFrustum + MapPoint::PredictScale for every map point of a batch of frames (orbgpu_is_in_frustum).
        cam: (n_frames, 24) float32 = Rcw(9) tcw(3) Ow(3) fx fy cx cy mbf minX maxX minY maxY."""
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 24)
        mp_off = np.ascontiguousarray(mp_off, np.int32)
        n = int(mp_off[-1])
        ins = [np.ascontiguousarray(a, np.float32) for a in (world_pos, normal, min_dist_inv, max_dist_inv, max_distance)]
        out = {"in_view": np.zeros(n, np.uint8), "proj_x": np.zeros(n, np.float32), "proj_y": np.zeros(n, np.float32),
               "proj_xr": np.zeros(n, np.float32), "level": np.zeros(n, np.int32), "view_cos": np.zeros(n, np.float32)}
        capi.check(self._lib.orbgpu_is_in_frustum(self._h, len(cam), cam.ctypes.data, float(log_scale_factor), n_levels, float(viewing_cos_limit),
                                                  mp_off.ctypes.data, *[a.ctypes.data for a in ins], *[out[k].ctypes.data for k in
                                                  ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos")]))
        return out

    def SearchForInitialization(self, frames2: FrameSet, queries1: "WindowQuerySet"):
        """ORBmatcher::SearchForInitialization for a batch of (F1, F2) pairs: vnMatches12 and the match count per pair."""
        nq = int(queries1.q_off[-1])
        m12, nm = np.zeros(nq, np.int32), np.zeros(frames2.n_frames, np.int32)
        capi.check(self._lib.orbgpu_search_for_initialization(self._h, C.byref(frames2.c), C.byref(queries1.c), self.mfNNratio,
                                                              int(self.mbCheckOrientation), m12.ctypes.data, nm.ctypes.data))
        return {"match12": m12, "nmatches": nm}

    def SearchWindowBest(self, frames: FrameSet, queries: "WindowQuerySet", inv_level_sigma2=None, skip_flagged: bool = False):
        """The candidate loops of Fuse / SearchBySim3: independent queries, best candidate only (orbgpu_search_window_best)."""
        nq = int(queries.q_off[-1])
        bi, bd = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        s2 = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        capi.check(self._lib.orbgpu_search_window_best(self._h, C.byref(frames.c), C.byref(queries.c), None if s2 is None else s2.ctypes.data,
                                                       0 if s2 is None else len(s2), int(skip_flagged), bi.ctypes.data, bd.ctypes.data))
        return {"q_best_idx": bi, "q_best_dist": bd}

    def ComputeDistinctiveDescriptors(self, obs_off, desc):
        """MapPoint::ComputeDistinctiveDescriptors for a batch of map points: (best row per point, its median distance)."""
        obs_off = np.ascontiguousarray(obs_off, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(obs_off) - 1
        idx, med = np.zeros(n, np.int32), np.zeros(n, np.int32)
        capi.check(self._lib.orbgpu_distinctive_descriptors(self._h, n, obs_off.ctypes.data, desc.ctypes.data, idx.ctypes.data, med.ctypes.data))
        return idx, med

    def project_mappoints(self, cam, log_scale_factor, n_levels, viewing_cos_limit, mp_off, world_pos, normal, min_dist_inv, max_dist_inv,
                          max_distance, flags, desc) -> C.c_void_p:
        """isInFrustum on the device straight into a device-resident map-point set (orbgpu_mappoint_set_project)."""
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 24)
        mp_off = np.ascontiguousarray(mp_off, np.int32)
        ins = [np.ascontiguousarray(a, np.float32) for a in (world_pos, normal, min_dist_inv, max_dist_inv, max_distance)]
        fl, ds = np.ascontiguousarray(flags, np.uint8), np.ascontiguousarray(desc, np.uint8)
        h = C.c_void_p()
        capi.check(self._lib.orbgpu_mappoint_set_project(self._h, len(cam), cam.ctypes.data, float(log_scale_factor), n_levels,
                                                         float(viewing_cos_limit), mp_off.ctypes.data, *[a.ctypes.data for a in ins],
                                                         fl.ctypes.data, ds.ctypes.data, C.byref(h)))
        return h

    def upload(self, fs: FrameSet) -> C.c_void_p:
        h = C.c_void_p()
        capi.check(self._lib.orbgpu_frame_set_upload(self._h, C.byref(fs.c), C.byref(h)))
        return h

    def release(self, h):
        self._lib.orbgpu_frame_set_release(h)

    def frame_set_from_extraction(self, extractor, vocabulary=None, levelsup: int = 4, kp_flag: int = 0, u_right_ptr: int = 0,
                                  u_right_stride: int = 0, grid=None) -> C.c_void_p:
        """Device-resident frame set from the batch `extractor` has just processed (+ FeatureVectors from `vocabulary`):
        extraction -> vocabulary -> matching without a host round trip (orbgpu_frame_set_from_extraction)."""
        h = C.c_void_p()
        g = None if grid is None else np.ascontiguousarray(grid, np.float32)
        capi.check(self._lib.orbgpu_frame_set_from_extraction(self._h, extractor._h, vocabulary._h if vocabulary is not None else None, levelsup,
                                                              kp_flag, u_right_ptr or None, u_right_stride, None if g is None else g.ctypes.data,
                                                              C.byref(h)))
        return h

    def frame_set_info(self, h):
        n, nn, nf = C.c_int(0), C.c_int(0), C.c_int(0)
        capi.check(self._lib.orbgpu_frame_set_dev_info(h, C.byref(n), None, C.byref(nn), C.byref(nf)))
        kp_off = np.zeros(n.value + 1, np.int32)
        capi.check(self._lib.orbgpu_frame_set_dev_info(h, None, kp_off.ctypes.data, None, None))
        return kp_off, nn.value, nf.value

    def frame_set_download(self, h):
        kp_off, nn, nf = self.frame_set_info(h)
        out = {"kp_off": kp_off, "keys_un": np.zeros(int(kp_off[-1]), capi.KP_DTYPE), "desc": np.zeros((int(kp_off[-1]), 32), np.uint8),
               "fv_node_off": np.zeros(len(kp_off), np.int32), "fv_node_id": np.zeros(nn, np.int32), "fv_feat_off": np.zeros(nn + 1, np.int32),
               "fv_feat": np.zeros(nf, np.int32)}
        capi.check(self._lib.orbgpu_frame_set_download(self._h, h, *[out[k].ctypes.data for k in
                                                                    ("keys_un", "desc", "fv_node_off", "fv_node_id", "fv_feat_off", "fv_feat")]))
        return out

    def upload_mappoints(self, mps: MapPointSet, n_frames: int) -> C.c_void_p:
        h = C.c_void_p()
        capi.check(self._lib.orbgpu_mappoint_set_upload(self._h, C.byref(mps.c), n_frames, C.byref(h)))
        return h

    def release_mappoints(self, h):
        self._lib.orbgpu_mappoint_set_release(h)

    def search_by_bow_dev(self, h1, h2, idx1, idx2, match_off, d_match12: int, d_match_dist: int, d_nmatches: int,
                          kf_frame: bool = False):
        capi.check(self._lib.orbgpu_search_by_bow_dev(self._h, h1, h2, len(idx1), idx1.ctypes.data, idx2.ctypes.data, self.mfNNratio,
                                                      int(self.mbCheckOrientation), self.TH_LOW, int(kf_frame), int(not kf_frame),
                                                      match_off.ctypes.data, d_match12, d_match_dist, d_nmatches))

    def search_for_triangulation_dev(self, h1, h2, idx1, idx2, F12, epipole, sf, s2, match_off, d_match12: int, d_match_dist: int,
                                     d_nmatches: int, bOnlyStereo: bool = False):
        capi.check(self._lib.orbgpu_search_for_triangulation_dev(self._h, h1, h2, len(idx1), idx1.ctypes.data, idx2.ctypes.data,
                                                                 F12.ctypes.data, epipole.ctypes.data, sf.ctypes.data, s2.ctypes.data,
                                                                 len(sf), int(bOnlyStereo), int(self.mbCheckOrientation),
                                                                 match_off.ctypes.data, d_match12, d_match_dist, d_nmatches))

    def search_by_projection_dev(self, hf, hm, sf, th, d_kp_match: int, d_best_idx: int, d_best_dist: int, d_second: int,
                                 d_nmatches: int):
        capi.check(self._lib.orbgpu_search_by_projection_dev(self._h, hf, hm, sf.ctypes.data, len(sf), th, self.mfNNratio, d_kp_match,
                                                             d_best_idx, d_best_dist, d_second, d_nmatches))

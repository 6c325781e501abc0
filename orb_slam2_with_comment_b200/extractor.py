"""Host-side mirror of ORB_SLAM2::ORBextractor (/root/reference/include/ORBextractor.h:45-111) over the C ABI.

Same constructor arguments, same getters, `__call__(image)` plays operator() and returns the keypoints (a
structured array with cv::KeyPoint's seven fields) and the N x 32 descriptor matrix; `mvImagePyramid` is read
back from the device on demand.  `extract_batch` is the frame-batched entry the benchmark drives.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int, *,
                 device: int = 0, max_width: int = 1241, max_height: int = 480, max_batch: int = 1):
        self._lib = capi.lib()
        self._h = C.c_void_p()
        self.nlevels, self.scaleFactor = nlevels, float(np.float32(scaleFactor))
        self.max_batch, self.device = max_batch, device
        capi.check(self._lib.orbgpu_extractor_create(C.byref(self._h), device, nfeatures, scaleFactor, nlevels, iniThFAST,
                                                     minThFAST, max_width, max_height, max_batch))
        self.kp_cap = self._lib.orbgpu_extractor_max_keypoints(self._h)
        s = np.zeros(4 * nlevels, np.float32)
        q = np.zeros(nlevels, np.int32)
        u = np.zeros(16, np.int32)
        capi.check(self._lib.orbgpu_extractor_tables(self._h, s.ctypes.data, q.ctypes.data, u.ctypes.data))
        self._tables = s.reshape(4, nlevels)
        self.mnFeaturesPerLevel, self.umax = q, u

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.orbgpu_extractor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ORBextractor.h:63-83
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self._tables[0].copy()
    def GetInverseScaleFactors(self): return self._tables[1].copy()
    def GetScaleSigmaSquares(self): return self._tables[2].copy()
    def GetInverseScaleSigmaSquares(self): return self._tables[3].copy()

    # ---- operator() (ORBextractor.cc:1043)
    def __call__(self, image: np.ndarray, mask=None):
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 image expected (ORBextractor.cc:1050)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        kp = np.zeros(self.kp_cap, KP_DTYPE)
        desc = np.zeros((self.kp_cap, 32), np.uint8)
        n = C.c_int(0)
        capi.check(self._lib.orbgpu_extract(self._h, image.ctypes.data, image.shape[1], image.shape[0], image.strides[0],
                                            kp.ctypes.data, desc.ctypes.data, self.kp_cap, C.byref(n)))
        return kp[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images: np.ndarray, kp_out=None, desc_out=None, counts=None):
        """images: (B, H, W) uint8, C-contiguous.  Returns (kp (B,cap), desc (B,cap,32), counts (B,))."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.flags.c_contiguous
        B, H, W = images.shape
        kp = np.zeros((B, self.kp_cap), KP_DTYPE) if kp_out is None else kp_out
        desc = np.zeros((B, self.kp_cap, 32), np.uint8) if desc_out is None else desc_out
        cnt = np.zeros(B, np.int32) if counts is None else counts
        capi.check(self._lib.orbgpu_extract_batch(self._h, images.ctypes.data, B, W, H, W, W * H, kp.ctypes.data,
                                                  desc.ctypes.data, self.kp_cap, cnt.ctypes.data))
        return kp, desc, cnt

    def extract_batch_color(self, images: np.ndarray, rgb: bool = True):
        """images: (B, H, W, 3 or 4) uint8 — Tracking::GrabImage*'s cvtColor fused in front of the extraction.
        Returns (gray (B,H,W), kp, desc, counts)."""
        assert images.dtype == np.uint8 and images.ndim == 4 and images.shape[3] in (3, 4) and images.flags.c_contiguous
        B, H, W, Cn = images.shape
        gray = np.zeros((B, H, W), np.uint8)
        kp = np.zeros((B, self.kp_cap), KP_DTYPE)
        desc = np.zeros((B, self.kp_cap, 32), np.uint8)
        cnt = np.zeros(B, np.int32)
        capi.check(self._lib.orbgpu_extract_batch_color(self._h, images.ctypes.data, B, W, H, Cn, int(rgb), W * Cn, W * H * Cn, gray.ctypes.data,
                                                        kp.ctypes.data, desc.ctypes.data, self.kp_cap, cnt.ctypes.data))
        return gray, kp, desc, cnt

    def extract_batch_dev(self, images_ptr: int, B: int, W: int, H: int, kp_ptr: int, desc_ptr: int, counts_ptr: int):
        """Device-resident variant: raw device pointers (e.g. torch tensors' data_ptr()); asynchronous."""
        capi.check(self._lib.orbgpu_extract_batch_dev(self._h, images_ptr, B, W, H, W, W * H, kp_ptr, desc_ptr, self.kp_cap,
                                                      counts_ptr))

    # ---- Frame::ComputeStereoMatches (Frame.cc:501-675): self = left extractor, both have just processed the same batch
    def stereo_matches(self, right: "ORBextractor", mb: float, mbf: float, batch: int = 1):
        """Returns (mvuRight, mvDepth) as (batch, kp_cap) float32 arrays (-1 = no stereo match)."""
        ur = np.full((batch, self.kp_cap), -1, np.float32)
        dp = np.full((batch, self.kp_cap), -1, np.float32)
        capi.check(self._lib.orbgpu_stereo_matches(self._h, right._h, mb, mbf, ur.ctypes.data, dp.ctypes.data, self.kp_cap))
        return ur, dp

    def stereo_matches_dev(self, right: "ORBextractor", mb: float, mbf: float, u_right_ptr: int, depth_ptr: int):
        capi.check(self._lib.orbgpu_stereo_matches_dev(self._h, right._h, mb, mbf, u_right_ptr, depth_ptr, self.kp_cap))

    def sync(self):
        capi.check(self._lib.orbgpu_extractor_sync(self._h))

    def stream(self) -> int:
        s = C.c_void_p()
        capi.check(self._lib.orbgpu_extractor_stream(self._h, C.byref(s)))
        return s.value or 0

    def last_launches(self) -> int:
        return self._lib.orbgpu_extractor_last_launches(self._h)

    STAGES = ("pyramid", "fast_cells", "octree", "blur", "orient_desc")

    def set_profiling(self, on: bool):
        capi.check(self._lib.orbgpu_extractor_set_profiling(self._h, int(on)))

    def set_eager_frame(self, on: bool):
        """Write the 19-px reflect-101 frame of every level with each call (as the reference does) instead of on the first
        bordered read-back; the bytes are the same either way."""
        capi.check(self._lib.orbgpu_extractor_set_eager_frame(self._h, int(on)))

    def stage_ms(self):
        ms = np.zeros(5, np.float32)
        capi.check(self._lib.orbgpu_extractor_stage_ms(self._h, ms.ctypes.data))
        return dict(zip(self.STAGES, (float(v) for v in ms)))

    # ---- mvImagePyramid (ORBextractor.h:86) and stage taps
    def level_dims(self, level):
        w, h = C.c_int(), C.c_int()
        capi.check(self._lib.orbgpu_extractor_level_dims(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level(self, level, frame=0, bordered=False):
        w, h = self.level_dims(level)
        shape = (h + 38, w + 38) if bordered else (h, w)
        out = np.zeros(shape, np.uint8)
        capi.check(self._lib.orbgpu_extractor_read_level(self._h, frame, level, int(bordered), out.ctypes.data, shape[1]))
        return out

    @property
    def mvImagePyramid(self):
        return [self.level(l) for l in range(self.nlevels)]

    def blurred(self, level, frame=0):
        w, h = self.level_dims(level)
        out = np.zeros((h, w), np.uint8)
        capi.check(self._lib.orbgpu_extractor_read_blurred(self._h, frame, level, out.ctypes.data, w))
        return out

    def level_points(self, level, stage, frame=0):
        cap = 1 << 18
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int(0)
        capi.check(self._lib.orbgpu_extractor_read_points(self._h, frame, level, stage, out.ctypes.data, cap, C.byref(n)))
        assert n.value <= cap
        return out[:n.value].copy()

    def octree(self, cand: np.ndarray, min_x, max_x, min_y, max_y, n_features):
        cand = np.ascontiguousarray(cand)
        out = np.zeros(n_features + 1024, KP_DTYPE)
        n = C.c_int(0)
        capi.check(self._lib.orbgpu_octree(self._h, cand.ctypes.data, len(cand), min_x, max_x, min_y, max_y, n_features,
                                           out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value].copy()

    def octree_last_path(self) -> int:
        """1: the pass-free construction answered the last octree() call, 0: the division-pass state machine."""
        return int(self._lib.orbgpu_octree_last_path(self._h))


class MultiGpuExtractor:
    """ORBextractor::operator() for a batch of frames over several GPUs of one box from ONE process
    (orbgpu_multi_extract_batch): contiguous frame ranges per device, one host thread and one copy/compute pipeline per
    device, results gathered in input order into one host array.  No collective: frames are independent."""

    def __init__(self, devices, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int, *,
                 max_width: int = 1241, max_height: int = 480, max_batch_per_device: int = 256):
        self._lib = capi.lib()
        L = self._lib
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.orbgpu_multi_extractor_create.argtypes = [C.POINTER(vp), vp, i, i, f, i, i, i, i, i, i]
        L.orbgpu_multi_extractor_destroy.argtypes = [vp]
        L.orbgpu_multi_extractor_device_count.argtypes = [vp]
        L.orbgpu_multi_extractor_max_keypoints.argtypes = [vp]
        L.orbgpu_multi_extractor_last_launches.argtypes = [vp]
        L.orbgpu_multi_extractor_frame_range.argtypes = [vp, i, i, C.POINTER(i), C.POINTER(i)]
        L.orbgpu_multi_extract_batch.argtypes = [vp, vp, i, i, i, C.c_size_t, C.c_size_t, vp, vp, i, vp]
        self._h = vp()
        dev = np.ascontiguousarray(list(devices), np.int32)
        capi.check(L.orbgpu_multi_extractor_create(C.byref(self._h), dev.ctypes.data, len(dev), nfeatures, scaleFactor, nlevels, iniThFAST,
                                                   minThFAST, max_width, max_height, max_batch_per_device))
        self.devices = dev.tolist()
        self.kp_cap = L.orbgpu_multi_extractor_max_keypoints(self._h)

    def frame_range(self, batch: int, g: int):
        a, b = C.c_int(), C.c_int()
        capi.check(self._lib.orbgpu_multi_extractor_frame_range(self._h, batch, g, C.byref(a), C.byref(b)))
        return a.value, b.value

    def extract_batch(self, images: np.ndarray, kp_out=None, desc_out=None, counts=None):
        """images: (B, H, W) uint8, C-contiguous (page-locked for full copy bandwidth).  Returns (kp (B,cap), desc (B,cap,32), counts (B,))."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.flags.c_contiguous
        B, H, W = images.shape
        kp = np.zeros((B, self.kp_cap), KP_DTYPE) if kp_out is None else kp_out
        desc = np.zeros((B, self.kp_cap, 32), np.uint8) if desc_out is None else desc_out
        cnt = np.zeros(B, np.int32) if counts is None else counts
        capi.check(self._lib.orbgpu_multi_extract_batch(self._h, images.ctypes.data, B, W, H, W, W * H, kp.ctypes.data, desc.ctypes.data,
                                                        self.kp_cap, cnt.ctypes.data))
        return kp, desc, cnt

    def last_launches(self) -> int:
        return self._lib.orbgpu_multi_extractor_last_launches(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.orbgpu_multi_extractor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

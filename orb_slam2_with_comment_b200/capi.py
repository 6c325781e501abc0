"""ctypes binding of liborbgpu.so (include/orbgpu.h).  Loading fails loudly when the library has not been
built; every compute call fails loudly without a CUDA device — there is no CPU fallback anywhere."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_lib = None


class OrbGpuError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        raise OrbGpuError(f"{path} is missing: run `python -m orb_slam2_with_comment_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    L = C.CDLL(path)
    L.orbgpu_last_error.restype = C.c_char_p
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.orbgpu_device_count.argtypes = [C.POINTER(i)]
    L.orbgpu_extractor_create.argtypes = [C.POINTER(vp), i, i, f, i, i, i, i, i, i]
    L.orbgpu_extractor_destroy.argtypes = [vp]
    L.orbgpu_extractor_tables.argtypes = [vp, vp, vp, vp]
    L.orbgpu_extractor_max_keypoints.argtypes = [vp]
    L.orbgpu_extract.argtypes = [vp, vp, i, i, sz, vp, vp, i, C.POINTER(i)]
    L.orbgpu_extract_batch.argtypes = [vp, vp, i, i, i, sz, sz, vp, vp, i, vp]
    L.orbgpu_extract_batch_color.argtypes = [vp, vp, i, i, i, i, i, sz, sz, vp, vp, vp, i, vp]
    L.orbgpu_extract_batch_dev.argtypes = [vp, vp, i, i, i, sz, sz, vp, vp, i, vp]
    L.orbgpu_extractor_sync.argtypes = [vp]
    L.orbgpu_extractor_stream.argtypes = [vp, C.POINTER(vp)]
    L.orbgpu_extractor_last_launches.argtypes = [vp]
    L.orbgpu_extractor_set_profiling.argtypes = [vp, i]
    L.orbgpu_extractor_set_eager_frame.argtypes = [vp, i]
    L.orbgpu_extractor_stage_ms.argtypes = [vp, vp]
    L.orbgpu_extractor_level_dims.argtypes = [vp, i, C.POINTER(i), C.POINTER(i)]
    L.orbgpu_extractor_read_level.argtypes = [vp, i, i, i, vp, sz]
    L.orbgpu_extractor_read_blurred.argtypes = [vp, i, i, vp, sz]
    L.orbgpu_extractor_read_points.argtypes = [vp, i, i, i, vp, i, C.POINTER(i)]
    L.orbgpu_octree.argtypes = [vp, vp, i, i, i, i, i, i, vp, i, C.POINTER(i)]
    L.orbgpu_octree_last_path.argtypes = [vp]
    L.orbgpu_extractor_static_tables.argtypes = [i, f, i, vp, vp, vp]
    L.orbgpu_stereo_matches.argtypes = [vp, vp, f, f, vp, vp, i]
    L.orbgpu_stereo_matches_dev.argtypes = [vp, vp, f, f, vp, vp, i]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise OrbGpuError(f"orbgpu error {rc}: {lib().orbgpu_last_error().decode()}")


def device_count() -> int:
    n = C.c_int(0)
    check(lib().orbgpu_device_count(C.byref(n)))
    return n.value

"""B200-native ORB front-end: CUDA (sm_100a) kernels behind ORB-SLAM2's ORBextractor / ORBmatcher interfaces.

The product is liborbgpu.so (csrc/, C ABI in include/orbgpu.h) plus C++ shells with the reference's class
signatures (csrc/host/).  This Python package is the thin host-side mirror used by tests and bench.py.
"""
from .extractor import ORBextractor, KP_DTYPE  # noqa: F401

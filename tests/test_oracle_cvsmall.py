"""CPU: the restated small-matrix cv::Mat algebra (oracle/cvlite.cc) against cv2 4.13 golden vectors."""
import ctypes as C
import os

import numpy as np

import oracle_lib

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cvsmall_golden.npz"))


def test_gemm_norm_match_cv2():
    lib = oracle_lib.load_port()
    fp = C.POINTER(C.c_float)
    lib.cvl_gemm3_f32.argtypes = [fp, fp, fp, fp]
    lib.cvl_gemm3t_neg_f32.argtypes = [fp, fp, fp]
    lib.cvl_norm3_f32.argtypes = [fp]
    lib.cvl_norm3_f32.restype = C.c_double
    A, x, c = G["A"], G["x"], G["c"]
    o = np.zeros(3, np.float32)
    for i in range(len(A)):
        lib.cvl_gemm3_f32(A[i].ctypes.data_as(fp), x[i].ctypes.data_as(fp), c[i].ctypes.data_as(fp), o.ctypes.data_as(fp))
        assert o.tobytes() == G["gemm"][i].tobytes(), i
        lib.cvl_gemm3t_neg_f32(A[i].ctypes.data_as(fp), x[i].ctypes.data_as(fp), o.ctypes.data_as(fp))
        assert o.tobytes() == G["gemm_t_neg"][i].tobytes(), i
        assert lib.cvl_norm3_f32(x[i].ctypes.data_as(fp)) == G["norm"][i], i

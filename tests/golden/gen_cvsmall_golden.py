"""cv2 4.13 golden vectors for the small float cv::Mat algebra of the reference's pose arithmetic: A*x+c (cv::gemm),
-A^T*x (GEMM_1_T), cv::norm.  Run in the build container:  python tests/golden/gen_cvsmall_golden.py"""
import os

import cv2
import numpy as np

rs = np.random.RandomState(4)
N = 4000
A = rs.uniform(-1, 1, (N, 3, 3)).astype(np.float32)
x = (rs.uniform(-60, 60, (N, 3, 1)) * rs.choice([1e-3, 1.0, 30.0], (N, 1, 1))).astype(np.float32)
c = rs.uniform(-8, 8, (N, 3, 1)).astype(np.float32)
out = {"A": A, "x": x, "c": c,
       "gemm": np.stack([cv2.gemm(A[i], x[i], 1.0, c[i], 1.0) for i in range(N)]),
       "gemm_t_neg": np.stack([cv2.gemm(A[i], x[i], -1.0, None, 0.0, flags=cv2.GEMM_1_T) for i in range(N)]),
       "norm": np.array([cv2.norm(x[i]) for i in range(N)], np.float64),
       "cv_version": np.array(cv2.__version__)}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cvsmall_golden.npz"), **out)
print("wrote", N)

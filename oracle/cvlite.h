/* cvlite — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain restatements, over raw 8-bit buffers, of the five OpenCV routines whose arithmetic decides
 * every output bit of the ORB front-end hot path.  OpenCV itself is an un-vendored, un-pinned
 * dependency of the reference (CMakeLists.txt:32-38) and no OpenCV C++ exists in this image, so each
 * routine is pinned instead by golden vectors generated from Python cv2 4.13.0
 * (tests/golden/gen_cv2_golden.py -> tests/golden/cv2_primitives.npz, checked by
 * tests/test_oracle_primitives.py).
 *
 * Reference call sites (all in /root/reference/src/ORBextractor.cc):
 *   resize(INTER_LINEAR)            :1120      -> cvl_resize_linear_u8
 *   copyMakeBorder(REFLECT_101)     :1122-1128 -> cvl_border_reflect101_u8
 *   FAST(img, kps, th, true)        :809,:814  -> cvl_fast9_16
 *   GaussianBlur(7x7, sigma 2)      :1086      -> cvl_gaussian7x7_u8
 *   fastAtan2                       :103       -> cvl_fast_atan2
 */
#ifndef ORBGPU_ORACLE_CVLITE_H
#define ORBGPU_ORACLE_CVLITE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvl_kp {
    int32_t x, y;   /* pixel position inside the (sub-)image handed to cvl_fast9_16 */
    int32_t score;  /* FAST corner score V = max(A,B)-1 (== cv::KeyPoint::response)   */
} cvl_kp;

/* INTER_LINEAR for 8UC1: 11-bit fixed-point weights, (>>4, *w >>16, +2 >>2) vertical pass. */
void cvl_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride);

/* dst is (w+left+right) x (h+top+bottom); interior copied (memmove-safe when src already sits inside
 * dst), frame filled with BORDER_REFLECT_101 of the interior. */
void cvl_border_reflect101_u8(const uint8_t* src, int w, int h, int sstride,
                              uint8_t* dst, int dstride, int top, int bottom, int left, int right);

/* GaussianBlur(Size(7,7), 2, 2, BORDER_REFLECT_101) for 8UC1: Q8.8 taps {18,34,48,56,48,34,18}. */
void cvl_gaussian7x7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);

/* FAST-9/16 with optional 3x3 non-maximum suppression; returns the number of corners (row-major
 * order); writes at most cap of them. */
int cvl_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms,
                 cvl_kp* out, int cap);

/* Threshold-free FAST score V(p) = max(A,B)-1 (0 within 3 px of the edge); corner_t <=> V >= t. */
void cvl_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* score, int score_stride);

float cvl_fast_atan2(float y, float x);
/* small float cv::Mat algebra of the pose arithmetic (Frame.cc:285, :312-320; ORBmatcher.cc:790-799, :1556-1563) */
void cvl_gemm3_f32(const float* A, const float* x, const float* c, float* out);
void cvl_gemm3t_neg_f32(const float* A, const float* x, float* out);
double cvl_norm3_f32(const float* v);
double cvl_dot3_f32(const float* a, const float* b);

#ifdef __cplusplus
}
#endif
#endif

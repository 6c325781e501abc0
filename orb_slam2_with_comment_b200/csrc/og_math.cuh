// og_math.cuh — per-element arithmetic of the ORB front-end, shared by every kernel.
// Each routine states the reference / OpenCV semantics it reproduces bit for bit.
// Functions are __host__ __device__ so that tests/host_model can run the very same code on the CPU;
// on the host the translation unit must be built with -ffp-contract=off.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OGM_HD __host__ __device__ __forceinline__
#else
#define OGM_HD inline
#endif

namespace og {

// Separately rounded float ops: nvcc contracts a*b+c into an FMA by default, the reference's arithmetic
// (OpenCV fastAtan2, the rounded rotation at ORBextractor.cc:118-120) is plain IEEE mul/add.
OGM_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
OGM_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
OGM_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
OGM_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
// cvRound(float): round half to even
OGM_HD int round_rn(float v) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)__builtin_lrintf(v);
#endif
}

// BORDER_REFLECT_101 for an index at most one period out of range (|overshoot| < n-1).
OGM_HD int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p;
}

// ---- cv::resize INTER_LINEAR, 8UC1 (Appendix A.1) ------------------------------------------------------
// h0/h1: horizontal interpolations of the two source rows (S[sx]*w0 + S[sx+1]*w1, 11-bit weights);
// b0/b1: vertical weights.
OGM_HD int resize_hpass(int s0, int s1, int w0, int w1) { return s0 * w0 + s1 * w1; }
OGM_HD uint8_t resize_vpass(int h0, int h1, int b0, int b1) {
    return (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

// ---- cv::GaussianBlur 7x7 sigma 2, 8UC1 (Appendix A.2): Q8.8 taps, Q16.16 accumulate, round -----------
#define OG_G0 18
#define OG_G1 34
#define OG_G2 48
#define OG_G3 56
OGM_HD int blur_tap7(int a, int b, int c, int d, int e, int f, int g) {
    return OG_G0 * (a + g) + OG_G1 * (b + f) + OG_G2 * (c + e) + OG_G3 * d;
}
OGM_HD uint8_t blur_finish(uint32_t acc) { return (uint8_t)((acc + 32768u) >> 16); }

OGM_HD int imin(int a, int b) { return a < b ? a : b; }
OGM_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- cv::fastAtan2 (Appendix A.4); p1..p7 are the float products computed once on the host ---------------
struct AtanCoef {
    float p1, p3, p5, p7, eps;
};
OGM_HD float fast_atan2(float y, float x, const AtanCoef& k) {
    const float ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
    float a, c, c2;
    if (ax >= ay) {
        c = fdiv(ay, fadd(ax, k.eps));
        c2 = fmul(c, c);
        a = fmul(fadd(fmul(fadd(fmul(fadd(fmul(k.p7, c2), k.p5), c2), k.p3), c2), k.p1), c);
    } else {
        c = fdiv(ax, fadd(ay, k.eps));
        c2 = fmul(c, c);
        a = fsub(90.f, fmul(fadd(fmul(fadd(fmul(fadd(fmul(k.p7, c2), k.p5), c2), k.p3), c2), k.p1), c));
    }
    if (x < 0) a = fsub(180.f, a);
    if (y < 0) a = fsub(360.f, a);
    return a;
}

// ---- rotated BRIEF sample offset (ORBextractor.cc:118-120) ---------------------------------------------
// row = cvRound(x*b + y*a), col = cvRound(x*a - y*b) with a = cos, b = sin of the keypoint angle.
OGM_HD void brief_offset(int px, int py, float a, float b, int* row, int* col) {
    const float fx = (float)px, fy = (float)py;
    *row = round_rn(fadd(fmul(fx, b), fmul(fy, a)));
    *col = round_rn(fsub(fmul(fx, a), fmul(fy, b)));
}

// The same with float pattern coordinates and cvRound done on the FMA pipe: for |v| < 2^22, v + 1.5 * 2^23 rounds v to
// an integer (round-half-even, the FADD's own rounding = cvRound's) in the low mantissa bits.
OGM_HD int round_rn_small(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000;
#else
    return (int)__builtin_lrintf(v);
#endif
}
OGM_HD void brief_offset_f(float fx, float fy, float a, float b, int* row, int* col) {
    *row = round_rn_small(fadd(fmul(fx, b), fmul(fy, a)));
    *col = round_rn_small(fsub(fmul(fx, a), fmul(fy, b)));
}

// ---- sincosf as glibc computes it (the reference's `(float)cos(angle), (float)sin(angle)` on a float argument compiles to one
// sincosf call, ORBextractor.cc:113): argument widened to double, reduced by quadrants of pi/2 with the prescaled 2/pi, one
// degree-8 cosine and degree-7 sine polynomial in double, rounded to float once.  This restatement equals glibc 2.39's
// sincosf bit for bit on every float in [2^-20, 2*pi] (checked exhaustively on the CPU, 189 M arguments, with and without
// FMA contraction), so the rotated-BRIEF sampling offsets — and with them the descriptors — are bit-exact, and it is an
// order of magnitude cheaper than the double-precision cos() + sin() it replaces.  Valid for |y| < 120.
OGM_HD void sincosf_glibc(float y, float* sinp, float* cosp) {
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    uint32_t bits;
#if defined(__CUDA_ARCH__)
    bits = __float_as_uint(y);
#else
    memcpy(&bits, &y, 4);
#endif
    const uint32_t top = (bits >> 20) & 0x7ffu;
    double x = (double)y;
    int n = 0;
    double cs = 1.0;   // the second table entry negates the cosine polynomial
    if (top < 0x3f4u) {                    // abstop12(y) < abstop12(pi/4)
        if (top < 0x398u) {                // abstop12(y) < abstop12(2^-12)
            *sinp = y;
            *cosp = 1.0f;
            return;
        }
    } else {
        const double r = x * hpi_inv;
        n = ((int)r + 0x800000) >> 24;     // (int32_t)r truncates toward zero
        x = fma(-(double)n, hpi, x);
        const int q = n & 3;
        if (q == 1 || q == 2) x = -x;      // sign[n & 3] = {1, -1, -1, 1}
        if (n & 2) cs = -1.0;
    }
    const double x2 = x * x, x4 = x2 * x2, x3 = x2 * x;
    const double c2 = fma(x2, cs * C4, cs * C3), s1 = fma(x2, S3, S2);
    const double c1 = fma(x2, cs * C1, cs * C0), x5 = x3 * x2, x6 = x4 * x2;
    const double sn = fma(x3, S1, x), c = fma(x4, cs * C2, c1);
    const float rs = (float)fma(x5, s1, sn), rc = (float)fma(x6, c2, c);
    if (n & 1) { *sinp = rc; *cosp = rs; } else { *sinp = rs; *cosp = rc; }
}

// ---- ORBmatcher::DescriptorDistance (ORBmatcher.cc:1901-1917): 256-bit Hamming distance ----------------
OGM_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

}  // namespace og

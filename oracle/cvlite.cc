// cvlite — TEST INFRASTRUCTURE (CPU oracle), not product code.  See cvlite.h for what each routine
// restates and how it is pinned (cv2 4.13.0 golden vectors).  Nothing under orb_slam2_with_comment_b200/ links this.
#include "cvlite.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline int round_half_even(float v) { return (int)lrintf(v); }

inline int reflect101(int p, int n) {
    // gfedcb|abcdefgh|gfedcba ; n >= 2 assumed for out-of-range p
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}

// ---- resize ---------------------------------------------------------------------------------
struct LinTab {
    std::vector<int> ofs;       // source index of the first tap
    std::vector<short> w0, w1;  // 11-bit weights
};

// Coefficient table for one axis.  `clamp_weights`: the horizontal axis zeroes the fraction when the
// first tap is clamped to an end (the second tap is then never read); the vertical axis keeps the
// weights and clips the two row indices independently.
LinTab make_tab(int sn, int dn, bool clamp_weights) {
    LinTab t;
    t.ofs.resize(dn);
    t.w0.resize(dn);
    t.w1.resize(dn);
    const double scale = (double)sn / dn;
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (clamp_weights) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        }
        t.ofs[d] = s;
        t.w0[d] = (short)round_half_even((1.f - f) * 2048.f);
        t.w1[d] = (short)round_half_even(f * 2048.f);
    }
    return t;
}

// ---- FAST -----------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline void ring_offsets(int stride, int off[16]) {
    for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * stride + kRingDx[k];
}

// V = max(A,B)-1 with A = max over the 16 arcs of 9 contiguous ring pixels of min(I(p)-I(q)),
// B the same for I(q)-I(p).
inline int fast_score(const uint8_t* p, const int off[16]) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int A = -256, B = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            if (d[k + j] < mn) mn = d[k + j];
            if (d[k + j] > mx) mx = d[k + j];
        }
        if (mn > A) A = mn;
        if (-mx > B) B = -mx;
    }
    return (A > B ? A : B) - 1;
}

}  // namespace

extern "C" {

void cvl_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                          int dstride) {
    if (dw <= 0 || dh <= 0) return;
    const LinTab tx = make_tab(sw, dw, true);
    const LinTab ty = make_tab(sh, dh, false);
    std::vector<int> row0(dw), row1(dw);
    int cached0 = -1, cached1 = -1;
    auto hpass = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = tx.ofs[dx];
            const int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            out[dx] = S[sx] * tx.w0[dx] + S[sx1] * tx.w1[dx];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = ty.ofs[dy], sy1 = ty.ofs[dy] + 1;
        sy0 = sy0 < 0 ? 0 : (sy0 > sh - 1 ? sh - 1 : sy0);
        sy1 = sy1 < 0 ? 0 : (sy1 > sh - 1 ? sh - 1 : sy1);
        if (sy0 == cached1) { row0.swap(row1); std::swap(cached0, cached1); }
        if (sy0 != cached0) { hpass(sy0, row0); cached0 = sy0; }
        if (sy1 == sy0) { row1 = row0; cached1 = sy1; }
        else if (sy1 != cached1) { hpass(sy1, row1); cached1 = sy1; }
        const int b0 = ty.w0[dy], b1 = ty.w1[dy];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

void cvl_border_reflect101_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int top,
                              int bottom, int left, int right) {
    uint8_t* inner = dst + (size_t)top * dstride + left;
    if (inner != src) {
        if (inner < src)
            for (int y = 0; y < h; ++y) memmove(inner + (size_t)y * dstride, src + (size_t)y * sstride, w);
        else
            for (int y = h - 1; y >= 0; --y) memmove(inner + (size_t)y * dstride, src + (size_t)y * sstride, w);
    }
    for (int y = 0; y < h; ++y) {
        uint8_t* row = inner + (size_t)y * dstride;
        for (int x = 1; x <= left; ++x) row[-x] = row[reflect101(-x, w)];
        for (int x = 0; x < right; ++x) row[w + x] = row[reflect101(w + x, w)];
    }
    const int W = w + left + right;
    for (int y = 1; y <= top; ++y)
        memcpy(dst + (size_t)(top - y) * dstride, dst + (size_t)(top + reflect101(-y, h)) * dstride, W);
    for (int y = 0; y < bottom; ++y)
        memcpy(dst + (size_t)(top + h + y) * dstride, dst + (size_t)(top + reflect101(h + y, h)) * dstride, W);
}

void cvl_gaussian7x7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* T = tmp.data() + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= 3 && x < w - 3) {
                for (int i = 0; i < 7; ++i) acc += K[i] * S[x + i - 3];
            } else {
                for (int i = 0; i < 7; ++i) acc += K[i] * S[reflect101(x + i - 3, w)];
            }
            T[x] = (uint16_t)acc;
        }
    }
    std::vector<uint8_t> out((size_t)w * h);  // src and dst may alias (in-place blur of a clone, :1086)
    for (int y = 0; y < h; ++y) {
        const uint16_t* R[7];
        for (int j = 0; j < 7; ++j) R[j] = tmp.data() + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t* D = out.data() + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int j = 0; j < 7; ++j) acc += (uint32_t)K[j] * R[j][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    for (int y = 0; y < h; ++y) memcpy(dst + (size_t)y * dstride, out.data() + (size_t)y * w, w);
}

void cvl_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* score, int score_stride) {
    int off[16];
    ring_offsets(stride, off);
    for (int y = 0; y < h; ++y) {
        uint8_t* srow = score + (size_t)y * score_stride;
        memset(srow, 0, w);
        if (y < 3 || y >= h - 3) continue;
        for (int x = 3; x < w - 3; ++x) {
            int v = fast_score(img + (size_t)y * stride + x, off);
            srow[x] = (uint8_t)(v < 0 ? 0 : v);
        }
    }
}

int cvl_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms, cvl_kp* out, int cap) {
    if (w < 7 || h < 7) return 0;
    int off[16];
    ring_offsets(stride, off);
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;

    // class table: bit 0 = ring pixel darker than v-t, bit 1 = brighter than v+t
    uint8_t tab[512];
    for (int i = -255; i <= 255; ++i) tab[i + 255] = (uint8_t)(i < -threshold ? 1 : (i > threshold ? 2 : 0));

    static thread_local std::vector<int> sbuf;
    sbuf.assign((size_t)w * h, 0);
    int* S = sbuf.data();

    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = row + x;
            const uint8_t* t = tab + 255 - p[0];  // t[q] = class of ring value q
            int c = t[p[off[0]]] | t[p[off[8]]];
            if (!c) continue;
            c &= t[p[off[2]]] | t[p[off[10]]];
            c &= t[p[off[4]]] | t[p[off[12]]];
            c &= t[p[off[6]]] | t[p[off[14]]];
            if (!c) continue;
            c &= t[p[off[1]]] | t[p[off[9]]];
            c &= t[p[off[3]]] | t[p[off[11]]];
            c &= t[p[off[5]]] | t[p[off[13]]];
            c &= t[p[off[7]]] | t[p[off[15]]];
            if (!c) continue;
            bool corner = false;
            for (int bit = 1; bit <= 2 && !corner; bit <<= 1) {
                if (!(c & bit)) continue;
                int run = 0;
                for (int k = 0; k < 25; ++k) {
                    if (t[p[off[k & 15]]] & bit) {
                        if (++run > 8) { corner = true; break; }
                    } else
                        run = 0;
                }
            }
            if (corner) S[(size_t)y * w + x] = nms ? fast_score(p, off) : 1;
        }
    }

    int n = 0;
    for (int y = 3; y < h - 3; ++y) {
        const int* r0 = S + (size_t)(y - 1) * w;
        const int* r1 = S + (size_t)y * w;
        const int* r2 = S + (size_t)(y + 1) * w;
        for (int x = 3; x < w - 3; ++x) {
            const int s = r1[x];
            if (!s) continue;
            if (nms) {
                // s>0 is guaranteed for threshold>=1; threshold 0 corners with score 0 never survive
                // the strict comparison below, matching the zero-initialised score rows of OpenCV.
                if (!(s > r1[x - 1] && s > r1[x + 1] && s > r0[x - 1] && s > r0[x] && s > r0[x + 1] &&
                      s > r2[x - 1] && s > r2[x] && s > r2[x + 1]))
                    continue;
            }
            if (n < cap) {
                out[n].x = x;
                out[n].y = y;
                out[n].score = nms ? s : 0;
            }
            ++n;
        }
    }
    return n;
}

float cvl_fast_atan2(float y, float x) {
    // OpenCV's polynomial (degrees); every operation is a separately rounded float op (build with
    // -ffp-contract=off).
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---- small float cv::Mat algebra as the reference's pose arithmetic uses it (pinned to cv2 4.13 by tests/golden/cvsmall_golden.npz) ----
// d = A * x + c for a 3x3 and 3x1 CV_32F (cv::gemm, alpha = beta = 1, no transpose flags: the small-matrix path): the dot
// products are accumulated in float, left to right; the addition of c happens in double and is rounded once.
void cvl_gemm3_f32(const float* A, const float* x, const float* c, float* out) {
    for (int i = 0; i < 3; ++i) {
        const float t = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
        out[i] = (float)((double)t * 1.0 + (double)(c ? c[i] : 0.f) * 1.0);
    }
}
// d = -A^T * x (cv::gemm with GEMM_1_T, alpha = -1: the general path): products and sums in double, rounded once.
void cvl_gemm3t_neg_f32(const float* A, const float* x, float* out) {
    for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
        for (int j = 0; j < 3; ++j) acc += (double)A[3 * j + i] * (double)x[j];
        out[i] = (float)(-acc);
    }
}
// cv::norm(v) of a 3x1 CV_32F (NORM_L2): squares summed in double, sqrt in double.
double cvl_norm3_f32(const float* v) {
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += (double)v[i] * (double)v[i];
    return std::sqrt(s);
}
// a.dot(b) of 3x1 CV_32F: products and sum in double.
double cvl_dot3_f32(const float* a, const float* b) {
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += (double)a[i] * (double)b[i];
    return s;
}

}  // extern "C"

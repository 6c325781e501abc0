// og_extract.cu — CUDA kernels (sm_100a) of ORBextractor::operator() (ORBextractor.cc:1043-1132).
//
//   k_level0        the input image into the interior of level 0                     (:1127)
//   k_resize_tma    cv::resize INTER_LINEAR level l-1 -> l: 192 x 32 output tiles, source box by TMA (:1120)
//   k_resize4_pp / k_resize4   the same with the source words in registers (levels whose tiles do not fit the box; scale <= 2)
//   k_resize        the same, any scale, one pass incl. frame (fallback for scale > 2)
//   k_border_sides / k_border_caps   BORDER_REFLECT_101 frame of every level         (:1122-1128)
//   k_fast_seg      per-cell FAST-9/16 + 3x3 NMS + iniTh/minTh fallback (TMA tiles)  (:789-829)
//   k_octree        DistributeOctTree, one CTA per (frame, level)                    (:539-763)
//   k_blur_tma      GaussianBlur 7x7 sigma 2 (TMA tiles, DP4A)                       (:1085-1086)
//   k_orient_desc   IC_Angle + rotated BRIEF + final KeyPoint fields                 (:77-147, :837-847, :1095-1103)
//
// Integer stencil / compaction / popcount work: no tensor cores.  All kernels are batched over frames.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/orbgpu_pattern.inc"
#include "og_octree.cuh"
#include "og_octree2.cuh"
#include "og_tma.cuh"
#include "og_types.h"

namespace og {

__device__ __forceinline__ uint8_t* level_ptr(uint8_t* base, const Level& L, int frame) {
    return base + L.base + (long long)frame * L.frame_stride;
}

// ------------------------------------------------------------------------------------------------------------
// Level 0: copy of the input image into the interior of the level-0 buffer (the frame is written by the border kernels).
// One thread moves 16 pixels: the source row may start at any byte address, so it is read as five aligned words
// realigned with funnel shifts, and stored as one 16-byte vector (the interior starts 16-byte aligned).
// The last vector of a row is copied byte-wise so that nothing past the row's end is ever read.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_level0(const __grid_constant__ ExtractParams P, const uint8_t* __restrict__ images, long long row_stride,
                                                long long frame_stride, int nvec, uint32_t nvec_magic, int n_items) {
    const Level& L = P.lv[0];
    const int frame = P.frame0 + blockIdx.y;
    const int id = blockIdx.x * 256 + threadIdx.x;   // one item = 16 pixels; nvec items per row (no idle threads at row ends)
    if (id >= n_items) return;
    const int y = (int)__umulhi((uint32_t)id, nvec_magic);
    const int x = (id - y * nvec) * 16;
    const uint8_t* src = images + (long long)frame * frame_stride + (long long)y * row_stride + x;
    uint8_t* dst = level_ptr(P.pyr, L, frame) + (long long)(kEdge + y) * L.pitch + kXPad + x;
    if (x + 20 <= L.w) {   // the five words reach at most byte x + 19 of the row
        const int a = (int)(reinterpret_cast<uintptr_t>(src) & 3);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(src - a);
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
        const int sh = 8 * a;
        *reinterpret_cast<uint4*>(dst) = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                                                    __funnelshift_r(w3, w4, sh));
    } else {
        for (int k = 0; k < 16 && x + k < L.w; ++k) dst[k] = __ldg(src + k);
    }
}

// ------------------------------------------------------------------------------------------------------------
// cvtColor(im, mImGray, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) of Tracking::GrabImage* (Tracking.cc:173-198,
// :214-228): OpenCV's 8-bit fixed point, Y = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (pinned to cv2 4.13 by
// tests/golden/cvtcolor_golden.npz).  One thread converts 4 pixels into one aligned word of the packed gray image.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cvt_gray(const uint8_t* __restrict__ src, long long row_stride, long long frame_stride, int channels,
                                                  int r_first, int width, int height, uint8_t* __restrict__ gray) {
    const int frame = blockIdx.z, y = blockIdx.y;
    const int x = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x >= width) return;
    const uint8_t* s = src + (long long)frame * frame_stride + (long long)y * row_stride + (long long)x * channels;
    uint8_t* d = gray + ((long long)frame * height + y) * width + x;
    const int cr = r_first ? 9798 : 3735, cb = r_first ? 3735 : 9798;   // weight of channel 0 / channel 2
    uint32_t out = 0;
    const int n = min(4, width - x);
    for (int k = 0; k < n; ++k) {
        const int c0 = __ldg(s + k * channels), c1 = __ldg(s + k * channels + 1), c2 = __ldg(s + k * channels + 2);
        out |= (uint32_t)((c0 * cr + c1 * 19235 + c2 * cb + (1 << 14)) >> 15) << (8 * k);
    }
    if (n == 4 && ((width & 3) == 0)) {
        *reinterpret_cast<uint32_t*>(d) = out;   // rows of a packed image whose width is a multiple of 4 stay word aligned
    } else {
        for (int k = 0; k < n; ++k) d[k] = (uint8_t)(out >> (8 * k));
    }
}

// ------------------------------------------------------------------------------------------------------------
// Level l from level l-1 (chained, :1120) including the border of the new level (:1122-1123): a border pixel is
// the resize output at its reflected interior coordinate, so one pass writes the whole padded row.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_resize(const __grid_constant__ ExtractParams P, int level) {
    const Level& L = P.lv[level];
    const Level& S = P.lv[level - 1];
    const int frame = P.frame0 + blockIdx.z;
    const int Y = blockIdx.y;
    const int X4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (X4 >= L.pitch) return;
    const int y = reflect101(Y - kEdge, L.h);
    const Tap ty = L.yt[y];
    const uint8_t* src = level_ptr(P.pyr, S, frame) + (long long)kEdge * S.pitch + kXPad;
    const uint8_t* r0 = src + (long long)ty.s0 * S.pitch;
    const uint8_t* r1 = src + (long long)ty.s1 * S.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int x = X4 + k - kXPad;
        x = x < -kEdge ? -kEdge : (x > L.w + kEdge - 1 ? L.w + kEdge - 1 : x);
        x = reflect101(x, L.w);
        const Tap tx = L.xt[x];
        const int h0 = resize_hpass(r0[tx.s0], r0[tx.s1], tx.w0, tx.w1);
        const int h1 = resize_hpass(r1[tx.s0], r1[tx.s1], tx.w0, tx.w1);
        out |= (uint32_t)resize_vpass(h0, h1, ty.w0, ty.w1) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(level_ptr(P.pyr, L, frame) + (long long)Y * L.pitch + X4) = out;
}

// ------------------------------------------------------------------------------------------------------------
// Level l from level l-1, interior only, vectorised: a thread produces 4 consecutive pixels (one 32-bit store) of
// kResizeRows consecutive rows.  The 4 outputs read at most 12 consecutive source bytes (scale <= 2), fetched as
// three aligned words per source row; the byte pair of every output is cut out with a funnel shift whose word
// choice and shift depend only on x and are computed once.  Consecutive output rows share a source row about
// every other step (scale 1.2): the horizontal pass of the shared row is reused.
// The 19-px frames of all levels are written afterwards by the border kernels.
// ------------------------------------------------------------------------------------------------------------
constexpr int kResizeRows = 8, kResizeThreads = 128;

__global__ void __launch_bounds__(kResizeThreads) k_resize4(const __grid_constant__ ExtractParams P, int level, int nwx, uint32_t nwx_magic,
                                                            int n_items) {
    const Level& L = P.lv[level];
    const Level& S = P.lv[level - 1];
    const int frame = P.frame0 + blockIdx.y;
    const int id = blockIdx.x * kResizeThreads + threadIdx.x;
    if (id >= n_items) return;
    const int band = (int)__umulhi((uint32_t)id, nwx_magic), wx = id - band * nwx;
    const int x = 4 * wx, y0 = band * kResizeRows;
    // x taps of the 4 outputs (the table is padded to a multiple of 4 entries)
    const uint4 ta = __ldg(reinterpret_cast<const uint4*>(L.xt + x)), tb = __ldg(reinterpret_cast<const uint4*>(L.xt + x) + 1);
    const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};   // per output: (s0 | s1 << 16), (w0 | w1 << 16)
    const int base = (int)(tw[0] & 0xffffu) & ~3;
    int sel[4], sh[4];
    uint32_t wq[4];   // (w0 | w1 << 16): the two 11-bit weights as DP2A's 16-bit operand pair
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = (int)(tw[2 * j] & 0xffffu) - base;   // 0..9
        sel[j] = o >> 2;
        sh[j] = 8 * (o & 3);
        wq[j] = tw[2 * j + 1];
    }
    const uint8_t* src = level_ptr(P.pyr, S, frame) + (long long)kEdge * S.pitch + kXPad + base;
    uint8_t* dst = level_ptr(P.pyr, L, frame) + (long long)kEdge * L.pitch + kXPad + x;
    // horizontal pass of one source row for the 4 outputs: g = (S[sx]*w0 + S[sx+1]*w1) >> 4, one DP2A per output
    auto hrow = [&](int sy, uint32_t (&g)[4]) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(src + (long long)sy * S.pitch);
        const uint32_t W0 = __ldg(r), W1 = __ldg(r + 1), W2 = __ldg(r + 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = sel[j] == 0 ? W0 : (sel[j] == 1 ? W1 : W2), hi = sel[j] == 0 ? W1 : W2;
            const uint32_t pr = __funnelshift_r(lo, hi, sh[j]);   // bytes 0,1 = the two source pixels
            g[j] = __dp2a_lo(wq[j], pr, 0u) >> 4;
        }
    };
    uint32_t g0[4], g1[4];
    int prev_s1 = -1;
#pragma unroll 1
    for (int r = 0; r < kResizeRows; ++r) {
        const int y = y0 + r;
        if (y >= L.h) break;
        const Tap ty = L.yt[y];
        if (ty.s0 == prev_s1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) g0[j] = g1[j];
        } else {
            hrow(ty.s0, g0);
        }
        if (ty.s1 == ty.s0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) g1[j] = g0[j];
        } else {
            hrow(ty.s1, g1);
        }
        prev_s1 = ty.s1;
        // vertical pass: ((b0 * g0) >> 16) + ((b1 * g1) >> 16) + 2 >> 2, the products' high halves taken with one IMAD.HI each
        const uint32_t b0 = (uint32_t)ty.w0 << 16, b1 = (uint32_t)ty.w1 << 16;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (__umulhi(b0, g0[j]) + __umulhi(b1, g1[j]) + 2u) >> 2;
        *reinterpret_cast<uint32_t*>(dst + (long long)y * L.pitch) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    }
}

// The same with all loads up front: the 8 output rows of a thread read at most kResizeSpan consecutive source rows (true for
// scale <= 1.25, checked on the host per level).  All 3 x kResizeSpan word loads are issued back to back (the row loop above
// waits for every row's three loads before it can start the next).  The byte pairs of outputs 0,1 lie inside the first two of
// the three words and those of outputs 2,3 inside two adjacent ones, so ONE byte permute per output pair gathers the source
// bytes [p(s0) p(s0+1) p(s0') p(s0'+1)]; the two permuted words of every source row (8 bytes) are parked in shared memory (own
// slots only: no barrier), and the vertical loop picks the two rows an output row blends and runs the horizontal pass on them
// (low / high DP2A apply the 16-bit weight pairs).  Parking the gathered bytes instead of the four 32-bit horizontal sums halves
// the shared-memory traffic (88 + 128 bytes per thread instead of 176 + 256) for 16 instead of 11 horizontal passes:
// 0.99 -> 0.94 ms per 1024 KITTI frames.
constexpr int kResizeSpan = 11;
#ifndef OG_RESIZE_MINB
#define OG_RESIZE_MINB 8
#endif
__global__ void __launch_bounds__(kResizeThreads, OG_RESIZE_MINB) k_resize4_pp(const __grid_constant__ ExtractParams P, int level, int nwx, uint32_t nwx_magic,
                                                               int n_items) {
    __shared__ uint2 sp[kResizeSpan][kResizeThreads];
    const Level& L = P.lv[level];
    const Level& S = P.lv[level - 1];
    const int frame = P.frame0 + blockIdx.y;
    const int id = blockIdx.x * kResizeThreads + threadIdx.x;
    if (id >= n_items) return;
    const int band = (int)__umulhi((uint32_t)id, nwx_magic), wx = id - band * nwx;
    const int x = 4 * wx, y0 = band * kResizeRows;
    const uint4 ta = __ldg(reinterpret_cast<const uint4*>(L.xt + x)), tb = __ldg(reinterpret_cast<const uint4*>(L.xt + x) + 1);
    const int s0 = (int)(ta.x & 0xffffu), base = s0 & ~3;
    const int o0 = s0 - base, o1 = (int)(ta.z & 0xffffu) - base;
    int o2 = (int)(tb.x & 0xffffu) - base, o3 = (int)(tb.z & 0xffffu) - base;
    const bool hiB = o2 >= 4;
    if (hiB) { o2 -= 4; o3 -= 4; }
    const uint32_t selA = (uint32_t)(o0 | ((o0 + 1) << 4) | (o1 << 8) | ((o1 + 1) << 12));
    const uint32_t selB = (uint32_t)(o2 | ((o2 + 1) << 4) | (o3 << 8) | ((o3 + 1) << 12));
    const uint32_t wq0 = ta.y, wq1 = ta.w, wq2 = tb.y, wq3 = tb.w;
    const int sfirst = L.yt[y0].s0;
    // 32-bit byte offsets from the CTA-uniform level base (a level image is far below 4 GB): one integer add per address
    const uint8_t* sbase = level_ptr(P.pyr, S, frame);
    const uint32_t soff = (uint32_t)(kEdge + sfirst) * (uint32_t)S.pitch + (uint32_t)(kXPad + base);
    // Rows sfirst + k past the level's last row (k <= 10, bottom band only) are never selected by the taps; their addresses lie in
    // the level's own bottom frame rows (kEdge = 19 of them), so no clamp is needed: the row address is one IMAD.WIDE.
    static_assert(kResizeSpan - 1 <= kEdge, "unclamped source rows stay inside the level's frame rows");
    const uint32_t spitch = (uint32_t)S.pitch;
    uint32_t W[kResizeSpan][3];
#pragma unroll
    for (int k = 0; k < kResizeSpan; ++k) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(sbase + (soff + spitch * (uint32_t)k));
        W[k][0] = __ldg(r); W[k][1] = __ldg(r + 1); W[k][2] = __ldg(r + 2);
    }
#pragma unroll
    for (int k = 0; k < kResizeSpan; ++k)
        sp[k][threadIdx.x] = make_uint2(__byte_perm(W[k][0], W[k][1], selA), __byte_perm(hiB ? W[k][1] : W[k][0], hiB ? W[k][2] : W[k][1], selB));
    uint8_t* dbase = level_ptr(P.pyr, L, frame);
    const uint32_t dpitch = (uint32_t)L.pitch;
    uint32_t doff = (uint32_t)(kEdge + y0) * dpitch + (uint32_t)(kXPad + x);
    const int nrows = min(kResizeRows, L.h - y0);
    const Tap* ytp = L.yt + y0;
#pragma unroll 2
    for (int r = 0; r < nrows; ++r, doff += dpitch) {
        const Tap ty = ytp[r];
        const uint2 a = sp[ty.s0 - sfirst][threadIdx.x], b = sp[ty.s1 - sfirst][threadIdx.x];   // own slots only: no barrier needed
        const uint32_t b0 = (uint32_t)ty.w0 << 16, b1 = (uint32_t)ty.w1 << 16;
        const uint32_t q0 = (__umulhi(b0, __dp2a_lo(wq0, a.x, 0u) >> 4) + __umulhi(b1, __dp2a_lo(wq0, b.x, 0u) >> 4) + 2u) >> 2;
        const uint32_t q1 = (__umulhi(b0, __dp2a_hi(wq1, a.x, 0u) >> 4) + __umulhi(b1, __dp2a_hi(wq1, b.x, 0u) >> 4) + 2u) >> 2;
        const uint32_t q2 = (__umulhi(b0, __dp2a_lo(wq2, a.y, 0u) >> 4) + __umulhi(b1, __dp2a_lo(wq2, b.y, 0u) >> 4) + 2u) >> 2;
        const uint32_t q3 = (__umulhi(b0, __dp2a_hi(wq3, a.y, 0u) >> 4) + __umulhi(b1, __dp2a_hi(wq3, b.y, 0u) >> 4) + 2u) >> 2;
        *reinterpret_cast<uint32_t*>(dbase + doff) = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
    }
}

// The same resize with the source staged by TMA: one CTA = 192 x 32 outputs of one (level, frame); the source pixels they
// blend (scale <= 1.25: at most 242 columns x 42 rows, checked per level on the host) arrive as ONE 256 x 42 box of level l-1
// (UTMALDG, 16-byte aligned start column at or left of the first source column).  A thread owns 4 output columns x 8 rows: per
// output row it reads the three words around its source bytes in the two source rows from shared memory, gathers the byte
// pairs with one permute per output pair and applies the weights by DP2A as above.  No global load instruction, no tag
// look-ups, nothing parked: 33 registers instead of 64.
// tmaps[4 K + l]: 256 x kRtBoxH boxes over level l (K = kMaxLevels).
constexpr int kRtW = 192, kRtH = 32, kRtBoxH = 42, kRtRows = 8, kRtThreads = (kRtW / 4) * (kRtH / kRtRows);
__global__ void __launch_bounds__(kRtThreads) k_resize_tma(const __grid_constant__ ExtractParams P, int level, int ntx, const CUtensorMap* __restrict__ tmaps) {
    __shared__ __align__(128) uint8_t tile[kRtBoxH * 256 + 16];
    __shared__ uint64_t mbar;
    const Level& L = P.lv[level];
    const int frame = P.frame0 + blockIdx.y, t = threadIdx.x;
    const int ty_ = blockIdx.x / ntx, tx_ = blockIdx.x - ty_ * ntx;
    const int X0 = kRtW * tx_, Y0 = kRtH * ty_;
    const int sy0 = L.yt[Y0].s0;                                    // first source row of the tile
    const int gx = (kXPad + (int)L.xt[X0].s0) & ~15;                // box start column (buffer coordinates)
    if (t == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(&mbar, kRtBoxH * 256);
        tma_load_3d(tile, tmaps + 4 * kMaxLevels + (level - 1), gx, kEdge + sy0, frame, &mbar);
    }
    const int g = t / (kRtW / 4), q = t - g * (kRtW / 4);
    const int x = X0 + 4 * q, y0 = Y0 + kRtRows * g;
    const bool active = x < L.w && y0 < L.h;
    uint32_t selA = 0, selB = 0, wq0 = 0, wq1 = 0, wq2 = 0, wq3 = 0;
    int base = 0;
    bool hiB = false;
    if (active) {   // x taps of the 4 outputs, while the box is in flight (the table is padded to a multiple of 4 entries)
        const uint4 ta = __ldg(reinterpret_cast<const uint4*>(L.xt + x)), tb = __ldg(reinterpret_cast<const uint4*>(L.xt + x) + 1);
        const int c0 = kXPad - gx;                                   // source column -> byte of a box row
        const int b0 = (int)(ta.x & 0xffffu) + c0;
        base = b0 & ~3;
        const int o0 = b0 - base, o1 = (int)(ta.z & 0xffffu) + c0 - base;
        int o2 = (int)(tb.x & 0xffffu) + c0 - base, o3 = (int)(tb.z & 0xffffu) + c0 - base;
        hiB = o2 >= 4;
        if (hiB) { o2 -= 4; o3 -= 4; }
        selA = (uint32_t)(o0 | ((o0 + 1) << 4) | (o1 << 8) | ((o1 + 1) << 12));
        selB = (uint32_t)(o2 | ((o2 + 1) << 4) | (o3 << 8) | ((o3 + 1) << 12));
        wq0 = ta.y; wq1 = ta.w; wq2 = tb.y; wq3 = tb.w;
    }
    mbar_wait(&mbar, 0);
    if (!active) return;
    const uint8_t* col = tile + base - sy0 * 256;   // the decoded taps carry source row * 256
    const size_t dpitch = (size_t)L.pitch;          // widened once: the row pointer advances by one 64-bit add per row
    uint8_t* drow = level_ptr(P.pyr, L, frame) + (size_t)(kEdge + y0) * dpitch + (size_t)(kXPad + x);
    const int nrows = min(kRtRows, L.h - y0);
    const uint4* ytp = L.ytw + y0;
    auto row = [&](int r) {
        const uint4 ty = __ldg(ytp + r);   // {s0 * 256, s1 * 256, w0 << 16, w1 << 16}
        const uint32_t* ra = reinterpret_cast<const uint32_t*>(col + ty.x);
        const uint32_t* rb = reinterpret_cast<const uint32_t*>(col + ty.y);
        const uint32_t a0 = ra[0], a1 = ra[1], a2 = ra[2], c0 = rb[0], c1 = rb[1], c2 = rb[2];
        const uint32_t ax = __byte_perm(a0, a1, selA), ay = __byte_perm(hiB ? a1 : a0, hiB ? a2 : a1, selB);
        const uint32_t bx = __byte_perm(c0, c1, selA), by = __byte_perm(hiB ? c1 : c0, hiB ? c2 : c1, selB);
        const uint32_t b0 = ty.z, b1 = ty.w;
        const uint32_t q0 = (__umulhi(b0, __dp2a_lo(wq0, ax, 0u) >> 4) + __umulhi(b1, __dp2a_lo(wq0, bx, 0u) >> 4) + 2u) >> 2;
        const uint32_t q1 = (__umulhi(b0, __dp2a_hi(wq1, ax, 0u) >> 4) + __umulhi(b1, __dp2a_hi(wq1, bx, 0u) >> 4) + 2u) >> 2;
        const uint32_t q2 = (__umulhi(b0, __dp2a_lo(wq2, ay, 0u) >> 4) + __umulhi(b1, __dp2a_lo(wq2, by, 0u) >> 4) + 2u) >> 2;
        const uint32_t q3 = (__umulhi(b0, __dp2a_hi(wq3, ay, 0u) >> 4) + __umulhi(b1, __dp2a_hi(wq3, by, 0u) >> 4) + 2u) >> 2;
        *reinterpret_cast<uint32_t*>(drow) = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
        drow += dpitch;
    };
    if (nrows == kRtRows) {   // all but a level's last rows: no per-row checks
#pragma unroll
        for (int r = 0; r < kRtRows; ++r) row(r);
    } else {
        for (int r = 0; r < nrows; ++r) row(r);
    }
}

// The BORDER_REFLECT_101 frame of every level (:1122-1123, :1127), in two kernels.
//   k_border_sides  one thread per (row, side): the 19 frame bytes left or right of an interior row are the mirrored
//                   interior bytes 1..19 (w-2..w-20).  They are read as aligned words and written as aligned words: a
//                   frame word is four consecutive interior bytes in reverse order = funnel shift + PRMT.  The words that
//                   straddle the frame's ends also cover padding bytes of the row (nobody reads those) or, on the right,
//                   the last interior bytes (rewritten with their own values).
//   k_border_caps   then the 19 rows above and below are whole-row copies of padded rows (the corners come along), moved as
//                   16-byte vectors.
// grid = (items, level, frame): blocks beyond a level's item count exit at once.
// ------------------------------------------------------------------------------------------------------------
constexpr int kBorderThreads = 128;
__global__ void __launch_bounds__(kBorderThreads) k_border_sides(const __grid_constant__ ExtractParams P) {
    const Level& L = P.lv[blockIdx.y];
    const int it = blockIdx.x * kBorderThreads + threadIdx.x;
    if (it >= 2 * L.h) return;
    const int frame = P.frame0 + blockIdx.z;
    const int y = it >> 1;
    uint8_t* row = level_ptr(P.pyr, L, frame) + (long long)(kEdge + y) * L.pitch;   // 128-byte aligned
    uint32_t* rw = reinterpret_cast<uint32_t*>(row);
    if ((it & 1) == 0) {
        // left: frame byte 32 - k = interior byte 32 + k (k = 1..19); the word at byte a takes bytes 64-a, 63-a, 62-a, 61-a
        const uint32_t w8 = rw[8], w9 = rw[9], w10 = rw[10], w11 = rw[11], w12 = rw[12], w13 = rw[13];
        rw[3] = __byte_perm(w12, w13, 0x1234);   // bytes 12..15 (12 is padding)
        *reinterpret_cast<uint4*>(row + 16) = make_uint4(__byte_perm(w11, w12, 0x1234), __byte_perm(w10, w11, 0x1234),
                                                         __byte_perm(w9, w10, 0x1234), __byte_perm(w8, w9, 0x1234));
    } else {
        // right: frame byte E + j = byte E - 2 - j (j = 0..18), E = 32 + w.  The word at byte a takes bytes s, s-1, s-2, s-3
        // with s = 2E - 2 - a; in the first word the bytes below E keep their interior values.
        const int E = kXPad + L.w, a0 = E & ~3, nw = ((E + kEdge - 1) >> 2) - (E >> 2) + 1;
        const uint32_t keep = (1u << (8 * (E & 3))) - 1u;   // 0 when E is word aligned
        for (int k = 0; k < nw; ++k) {
            const int a = a0 + 4 * k, lo = 2 * E - 5 - a;   // source bytes lo .. lo + 3
            const uint32_t* sw = rw + (lo >> 2);
            const uint32_t x = __funnelshift_r(sw[0], sw[1], 8 * (lo & 3));
            uint32_t out = __byte_perm(x, 0u, 0x0123);
            if (k == 0 && keep) out = (rw[a0 >> 2] & keep) | (out & ~keep);
            rw[a >> 2] = out;
        }
    }
}

__global__ void __launch_bounds__(kBorderThreads) k_border_caps(const __grid_constant__ ExtractParams P) {
    const Level& L = P.lv[blockIdx.y];
    const int rr = blockIdx.x;                                   // 0..18 rows above, 19..37 rows below
    const int frame = P.frame0 + blockIdx.z;
    const int Y = rr < kEdge ? rr : L.h + rr;                    // padded row index
    const int y = reflect101(Y - kEdge, L.h);
    uint8_t* img = level_ptr(P.pyr, L, frame);
    const uint4* srow = reinterpret_cast<const uint4*>(img + (long long)(kEdge + y) * L.pitch);
    uint4* drow = reinterpret_cast<uint4*>(img + (long long)Y * L.pitch);
    const int nvec = (kXPad + L.w + kEdge + 15) >> 4;            // through the last frame byte (<= pitch / 16)
    for (int v = threadIdx.x; v < nvec; v += kBorderThreads) drow[v] = srow[v];
}

// ------------------------------------------------------------------------------------------------------------
// DistributeOctTree: one CTA per (level, frame).  Gathers the level's candidates from the per-cell slots in
// emission order (cell-row-major, row-major inside a cell), then runs the block-cooperative state machine.
// ------------------------------------------------------------------------------------------------------------
#ifndef OG_OCT_THREADS
#define OG_OCT_THREADS 128
#endif
constexpr int kOctThreads = OG_OCT_THREADS;
constexpr int kOctSmem = 160 * 1024;   // shared-memory workspace budget of one octree CTA in latency mode (small batches)
constexpr int kOctSmemMaxBatch = 16;
constexpr int kOctLatThreads = 512;      // CTA size in latency mode (per-thread key chunks shrink 4x; occupancy is irrelevant there)
constexpr int kOctMaxThreads = 1024;     // sizes the per-thread scan scratch
#ifndef OG_OT2_BUDGET
#define OG_OT2_BUDGET 16384
#endif
constexpr int kOctDirectSmem = 74 * 1024;    // shared memory of one CTA of the pass-free octree (three CTAs per SM)
constexpr int kOt2Budget = OG_OT2_BUDGET;   // cells of the deepest histogram level of the pass-free octree (2 bytes each)

__device__ __forceinline__ OtWork carve_work(uint8_t* ws, int cap, int node_cap) {
    OtWork W;
    auto take = [&](size_t bytes) { uint8_t* p = ws; ws += (bytes + 15) & ~size_t(15); return p; };
    W.kxy[0] = (uint32_t*)take((size_t)cap * 4);
    W.kxy[1] = (uint32_t*)take((size_t)cap * 4);
    W.knode[0] = (uint16_t*)take((size_t)cap * 2);
    W.knode[1] = (uint16_t*)take((size_t)cap * 2);
    W.kresp[0] = (uint8_t*)take((size_t)cap);
    W.kresp[1] = (uint8_t*)take((size_t)cap);
    W.nodes[0] = (OtNode*)take((size_t)node_cap * sizeof(OtNode));
    W.nodes[1] = (OtNode*)take((size_t)node_cap * sizeof(OtNode));
    W.tmp = (OtTmp*)take((size_t)node_cap * sizeof(OtTmp));
    W.R[0] = (int32_t*)take((size_t)node_cap * 4);
    W.R[1] = (int32_t*)take((size_t)node_cap * 4);
    W.ordv = (int32_t*)take((size_t)node_cap * 4);
    W.ordv2 = (int32_t*)take((size_t)node_cap * 4);
    W.surv = (int32_t*)take((size_t)node_cap * 4);
    W.thr = (int32_t*)take((size_t)kOctMaxThreads * 4 * 4);
    W.cap = cap;
    W.node_cap = node_cap;
    return W;
}

// Workspace bytes of one (level, frame) for `cap` keys and `node_cap` nodes (same carving as carve_work).
__host__ __device__ inline size_t octree_ws_bytes_dev(int cap, int node_cap) {
    size_t b = 0;
    auto take = [&](size_t bytes) { b += (bytes + 15) & ~size_t(15); };
    take((size_t)cap * 4); take((size_t)cap * 4);
    take((size_t)cap * 2); take((size_t)cap * 2);
    take((size_t)cap); take((size_t)cap);
    take((size_t)node_cap * sizeof(OtNode)); take((size_t)node_cap * sizeof(OtNode));
    take((size_t)node_cap * sizeof(OtTmp));
    for (int i = 0; i < 5; ++i) take((size_t)node_cap * 4);
    take((size_t)kOctMaxThreads * 16);
    return b;
}

// DistributeOctTree of one (level, frame): reads the level's candidates from the per-cell slots in emission order, writes the
// selected keys.  Called by all threads of a CTA.
// smem_budget > 0 (small batches, where the kernel's own latency is what counts): the workspace of the general path lives
// in shared memory when it fits — data written in one phase is read in the next, which from global memory is an L2 round
// trip (stores do not allocate in L1).  Large batches keep that workspace in HBM.
// direct_mem != nullptr: shared memory for the pass-free construction (og_octree2.cuh), tried first: ot2_smem_bytes() followed
// by 4 * n_cells bytes for the emission-order offsets of the cells.  It reads the candidates where k_fast_seg left them.
__device__ __forceinline__ void octree_level(const ExtractParams& P, int level, int frame, OtShared& sh, uint8_t* oct_smem, int smem_budget,
                                             uint8_t* direct_mem = nullptr, Ot2Shared* s2 = nullptr, int direct_kcap = 0) {
    const Level& L = P.lv[level];
    const int THREADS = (int)blockDim.x;
    const int32_t* ccount = P.cell_count + (long long)frame * P.total_cells + L.cell_base;
    const uint32_t* cxy = P.cand_xy + (long long)frame * P.total_cand_cap + L.cand_base;
    const uint8_t* crr = P.cand_resp + (long long)frame * P.total_cand_cap + L.cand_base;
    const Cell* cells = P.cells + L.cell_base;
    uint32_t* oxy = P.sel_xy + (long long)frame * P.total_sel_cap + L.sel_base;
    uint8_t* orr = P.sel_resp + (long long)frame * P.total_sel_cap + L.sel_base;
    int n = -1;
    if (direct_mem) {
        const int Dh = ot2_depth(L.n_ini, kOt2Budget), small_cap = max(L.node_cap, THREADS);
        int32_t* coff = reinterpret_cast<int32_t*>(direct_mem + ot2_smem_bytes(L.n_ini, Dh, small_cap, direct_kcap));
        OG_FOR(i, L.n_cells) coff[i] = ccount[i];
        OG_SYNC();
        block_exscan(coff, L.n_cells, &sh);
        const Ot2CellKeys keys{coff, cells, cxy, crr, L.n_cells};
        n = ot_run_direct(keys, sh.scan_total, direct_mem, Dh, small_cap, direct_kcap, &sh, s2, L.n_ini, L.hx, L.det_h, L.quota, oxy, orr, L.sel_cap);
    }
    if (n < 0) {
        // the general path: division passes over a workspace of keys
        bool in_smem = false;
        int cap_s = 0;
        if (smem_budget > 0) {
            // number of candidates of this level = sum of the per-cell counts
            int part = 0;
            for (int i = threadIdx.x; i < L.n_cells; i += THREADS) part += ccount[i];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            if ((threadIdx.x & 31) == 0) sh.warp_sums[threadIdx.x >> 5] = part;
            __syncthreads();
            int Mtot = 0;
            for (int w = 0; w < THREADS / 32; ++w) Mtot += sh.warp_sums[w];
            __syncthreads();
            cap_s = max(max(Mtot, L.n_cells), 1);   // kxy[0] doubles as the cell-offset scan buffer
            in_smem = octree_ws_bytes_dev(cap_s, L.node_cap) <= (size_t)smem_budget;
        }
        OtWork W = in_smem ? carve_work(oct_smem, cap_s, L.node_cap)
                           : carve_work(P.ot_ws + (long long)frame * P.ot_frame_bytes + L.ot_base, L.cand_cap, L.node_cap);
        // exclusive scan of the per-cell counts -> emission-order offsets (cells may exceed node_cap, so the scan lives in
        // kxy[0], which is free until the root partition)
        int32_t* coff = (int32_t*)W.kxy[0];
        OG_FOR(i, L.n_cells) coff[i] = ccount[i];
        OG_SYNC();
        block_exscan(coff, L.n_cells, &sh);
        const int M = sh.scan_total;
        // one thread per cell copies the cell's run (a handful of candidates)
        for (int ci = threadIdx.x; ci < L.n_cells; ci += THREADS) {
            const int nk = ccount[ci], dst = coff[ci], src = cells[ci].slot;
            for (int k = 0; k < nk; ++k) {
                W.kxy[1][dst + k] = cxy[src + k];
                W.kresp[1][dst + k] = crr[src + k];
            }
        }
        OG_SYNC();
        n = ot_run(W, &sh, M, L.n_ini, L.hx, L.det_h, L.quota, oxy, orr, L.sel_cap);
    }
    if (threadIdx.x == 0) P.sel_count[frame * P.n_levels + level] = n;
}

// Dynamic shared memory: [direct_bytes of the pass-free construction][smem_budget of latency-mode workspace].
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_octree(const __grid_constant__ ExtractParams P, int smem_budget, int direct_bytes, int direct_kcap) {
    extern __shared__ __align__(16) uint8_t oct_smem[];
    __shared__ OtShared sh;
    __shared__ Ot2Shared s2;
    octree_level(P, blockIdx.x, P.frame0 + blockIdx.y, sh, oct_smem + direct_bytes, smem_budget, direct_bytes ? oct_smem : nullptr, &s2, direct_kcap);
}

// Stand-alone octree on caller-provided candidates (stage-level parity tests, orbgpu_octree).  direct_bytes > 0: the pass-free
// construction is tried first (dynamic shared memory), as in the extraction path; 0 runs the division-pass state machine alone.
__global__ void __launch_bounds__(kOctThreads) k_octree_single(uint8_t* ws, int cap, int node_cap, const uint32_t* xy,
                                                               const uint8_t* resp, int M, int n_ini, float hx, int height,
                                                               int N, uint32_t* out_xy, uint8_t* out_resp, int out_cap,
                                                               int* out_n, int direct_bytes, int direct_kcap) {
    extern __shared__ __align__(16) uint8_t oct_smem[];
    __shared__ OtShared sh;
    __shared__ Ot2Shared s2;
    OtWork W = carve_work(ws, cap, node_cap);
    OG_FOR(i, M) { W.kxy[1][i] = xy[i]; W.kresp[1][i] = resp[i]; }
    OG_SYNC();
    int n = -1;
    if (direct_bytes)
        n = ot_run_direct(Ot2CompactKeys{W.kxy[1], W.kresp[1]}, M, oct_smem, ot2_depth(n_ini, kOt2Budget), max(node_cap, kOctThreads), direct_kcap, &sh,
                          &s2, n_ini, hx, height, N, out_xy, out_resp, out_cap);
    if (threadIdx.x == 0) out_n[1] = n >= 0;   // which path produced the result (tests)
    if (n < 0) n = ot_run(W, &sh, M, n_ini, hx, height, N, out_xy, out_resp, out_cap);
    if (threadIdx.x == 0) *out_n = n;
}

// ------------------------------------------------------------------------------------------------------------
// FAST per cell (:789-829): one CTA per (segment, frame), a segment being up to 8 consecutive cells of one
// cell row.  The tile arrives by TMA.  The work is organised so that only the cheap rejection test touches every
// pixel and everything expensive runs over a compacted queue:
//   1. SWAR rejection test, 4 pixels per 32-bit word: a FAST-9 arc always covers four consecutive of the eight even
//      ring pixels (compass points + diagonals), so a corner has four consecutive of them with |I - Ic| > t.
//      VABSDIFF4 + a carry-trick compare; 4-bit result per word -> mask bytes.
//   2. ordered-free compaction of the surviving pixels into a shared-memory queue (warp scan + one atomic).
//   3. per queued pixel: sign-aware compass test, then the exact score.  V = max(A,B)-1 with A/B the max over the
//      16 arcs of the min of d / -d: both halves are evaluated at once on packed 16-bit pairs
//      (256+d | 256-d << 16) with the 3-input VIMNMX3.U16x2 (two levels of min3 give the 9-wide arc minimum).
//   4. 3x3 NMS over the queue only (cell borders count as 0, :809 runs FAST per cell) -> survivor bitmaps, one for
//      V >= minTh and one for V >= iniTh.
//   5. one warp per cell: if the iniTh bitmap has any bit in the cell it is used, else the minTh one (:809-816);
//      row-major emission order comes from popcount prefix sums over the bitmap rows.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSegThreads = 256;
constexpr int kBmWords = 10;   // 256-bit bitmap row + 2 words so a 64-bit window can be read at any offset

__host__ __device__ inline int fast_seg_smem_bytes(int hbox, int th_max) {
    return hbox * kSegPitch + (th_max + 2) * kSegPitch + th_max * 64 + 2 * th_max * kBmWords * 4 + kSegMaxTw * th_max * 2 + 256 + 128;
}

// atomicAdd on a shared-memory counter as ONE instruction (ATOMS.ADD).  With the intrinsic the compiler wraps every call in a
// warp-aggregation loop (match / elect / shuffle, ~150 SASS instructions per site) although the callers already aggregate:
// that was 5.9 % of k_fast_seg's executed instructions (profiles/r2_fast_seg_v0_lines.txt).
__device__ __forceinline__ int smem_atomic_add(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ uint32_t gt_bytes(uint32_t x, uint32_t kadd, bool big) {
    // 0x80 in every byte of x that is > t.  t < 127: kadd = (0x7f - t) per byte, carry from the low 7 bits or bit 7 itself;
    // t >= 127: needs bit 7 and the low 7 bits > t - 128: kadd = (0xff - t) per byte.
    const uint32_t s = (x & 0x7f7f7f7fu) + kadd;
    return big ? (s & x & 0x80808080u) : ((s | x) & 0x80808080u);
}

__global__ void __launch_bounds__(kSegThreads, 5) k_fast_seg(const __grid_constant__ ExtractParams P, const CUtensorMap* __restrict__ tmaps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ int qn;

    const Segment sg = P.segs[blockIdx.x];
    const int frame = P.frame0 + blockIdx.y;
    const Level& L = P.lv[sg.level];
    const int t = threadIdx.x, lane = t & 31, wi = t >> 5;
    const int th = sg.th, tw = sg.tw, hbox = L.hbox;
    // TMA needs a 16-byte aligned start column: the tile begins at gx, tested pixel px sits at tile column ox + px
    const int gx = (kXPad + sg.x0 - 3) & ~15, ox = kXPad + sg.x0 - gx;   // 3 <= ox <= 18

    uint8_t* tile = smem;                                                   // [hbox][256]
    uint8_t* score = tile + hbox * kSegPitch;                               // [th+2][256], pixel (px,py) at [py+1][px+4]
    uint8_t* mk = score + (th + 2) * kSegPitch;                             // [th][64] 4-bit rejection results per word
    uint32_t* bm_min = reinterpret_cast<uint32_t*>(mk + th * 64);           // [th][kBmWords]
    uint32_t* bm_ini = bm_min + th * kBmWords;
    uint16_t* queue = reinterpret_cast<uint16_t*>(bm_ini + th * kBmWords);  // [tw*th] py << 8 | px
    uint8_t* lut = reinterpret_cast<uint8_t*>(queue + tw * th);             // [256] bit0: first column of its cell, bit1: last

    if (t == 0) {
        mbar_init(&mbar, 1);
        qn = 0;
    }
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(&mbar, (uint32_t)(hbox * kSegPitch));
        tma_load_3d(tile, tmaps + sg.level, gx, kEdge + sg.y0 - 3, frame, &mbar);
    }
    // while the tile is in flight: clear score map, masks, bitmaps; build the cell-column table
    {
        uint4* z = reinterpret_cast<uint4*>(score);
        const int nz = ((th + 2) * kSegPitch + th * 64 + 2 * th * kBmWords * 4) / 16;
        for (int i = t; i < nz; i += kSegThreads) z[i] = make_uint4(0, 0, 0, 0);
        const int wc = L.wcell;
        for (int px = t; px < 256; px += kSegThreads) {
            const int j = (px * L.wcell_magic) >> 16, lo = j * wc, hi = min(lo + wc, tw) - 1;
            lut[px] = (uint8_t)((px == lo ? 1 : 0) | (px == hi ? 2 : 0));
        }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);

    const int min_th = P.min_th, ini_th = P.ini_th;
    __shared__ uint32_t colmask[8];   // pass 1: tile columns of the cells that need the minTh pass (256 bits)
    __shared__ uint32_t vmask[8];     // tile columns of the tested pixels
    {
        const unsigned bal = __ballot_sync(0xffffffffu, t >= ox && t < ox + tw);
        if (lane == 0) vmask[wi] = bal;   // read after the barrier behind the rejection test
    }
    __shared__ int need[8], nneed;
    const int wcell = L.wcell;
    // Pass 0 works at iniTh for the whole segment: FAST(cell, iniTh) only involves pixels with V >= iniTh (a corner's
    // neighbours below the threshold cannot beat it), and far fewer pixels pass the rejection test at 20 than at 7.
    // Only cells left without a survivor (:812) are redone at minTh in pass 1 (a few percent of the cells).
    for (int pass = 0; pass < 2; ++pass) {
        const int tcur = pass == 0 ? ini_th : min_th;
        uint32_t* bm_out = pass == 0 ? bm_ini : bm_min;
        // ---- 1. rejection test, 8 pixels (two words) per step -> one byte of the row's 256-bit candidate bitmap -----
        {
            const uint32_t* T = reinterpret_cast<const uint32_t*>(tile);
            const int b0 = ox >> 3, nb = ((ox + tw - 1) >> 3) - b0 + 1, total = th * nb;
            const uint32_t kadd = (uint32_t)(tcur >= 127 ? 0xff - tcur : 0x7f - tcur) * 0x01010101u;
            // An arc of 9 contiguous ring pixels contains two neighbouring compass points — one of N / S and one of E / W — and two
            // neighbouring (+-2, +-2) diagonal points — one of NE / SW and one of SE / NW.  So a corner has
            //     (N | S) & (E | W) & (NE | SW) & (SE | NW),    X = |I_X - Ic| > t,
            // which on the bench frames lets through exactly as many pixels as "four consecutive of the eight" (10.2 %; the plain
            // compass test 17 - 18 %; true corners 2.7 %) for a fifth of its AND / OR network, and everything downstream —
            // compaction, exact score — scales with that count.
            // far2(): bit 7 of every byte says that one of the two ring pixels differs from the centre by more than t; the other bits
            // are garbage and are masked once, after the AND (t < 127: carry out of the low 7 bits, or bit 7 of the difference
            // itself; t >= 127: bit 7 and the carry).
            // The threshold class is a compile-time constant of the loop (one LOP3 per test instead of a predicated pair).
            auto reject = [&](auto big_c) {
            constexpr bool big = decltype(big_c)::value;
            // t < 127: the sum runs on the unmasked differences.  A byte >= 129 + t then carries into its left neighbour, whose
            // bit 7 — if its own bit 7 and carry are both clear — comes out set for a difference of t already instead of t + 1:
            // a rare extra candidate for the exact score to turn down, never a missed one (bit 7 of d itself covers d >= 128, and
            // a carry can only raise a byte's sum), so the two mask instructions per pair are not needed.
            auto far2 = [&](uint32_t c, uint32_t xa, uint32_t xb) {
                const uint32_t da = __vabsdiffu4(c, xa), db = __vabsdiffu4(c, xb);
                if (big) {
                    const uint32_t sa = (da & 0x7f7f7f7fu) + kadd, sb = (db & 0x7f7f7f7fu) + kadd;
                    return (sa & da) | (sb & db);
                }
                return (da + kadd) | da | (db + kadd) | db;
            };
            auto flags = [&](uint32_t ns, uint32_t ew, uint32_t d1, uint32_t d2) {
                const uint32_t g = ns & ew & d1 & d2 & 0x80808080u;
                return ((g >> 7) * 0x01020408u) >> 24;   // bits 7,15,23,31 -> 4-bit value
            };
            const uint8_t* cmb = reinterpret_cast<const uint8_t*>(colmask);
            for (int u = t; u < total; u += kSegThreads) {
                const int r = (int)__umulhi((uint32_t)u, sg.nw_magic);
                const int br = u - r * nb, B = b0 + br;
                uint32_t m = 0;
                if (pass == 0 || cmb[B]) {
                    const uint32_t* row = T + (r + 3) * 64 + 2 * B;
                    const uint32_t c0 = row[0], c1 = row[1], lw = row[-1], rw = row[2];
                    const uint32_t* ru = row - 2 * 64;   // two rows up / down: the diagonal ring pixels sit two columns left / right
                    const uint32_t* rd = row + 2 * 64;
                    const uint32_t ua = ru[-1], ub = ru[0], uc = ru[1], ud = ru[2], da = rd[-1], db = rd[0], dc = rd[1], dd = rd[2];
                    const uint32_t u_l = __funnelshift_r(ua, ub, 16), u_m = __funnelshift_r(ub, uc, 16), u_r = __funnelshift_r(uc, ud, 16);
                    const uint32_t d_l = __funnelshift_r(da, db, 16), d_m = __funnelshift_r(db, dc, 16), d_r = __funnelshift_r(dc, dd, 16);
                    // pairs (N, S), (E, W), (NE, SW), (SE, NW); the columns outside the tested range are masked by the compaction
                    const uint32_t n0 = flags(far2(c0, row[-3 * 64], row[3 * 64]), far2(c0, __funnelshift_r(c0, c1, 24), __funnelshift_r(lw, c0, 8)),
                                              far2(c0, u_m, d_l), far2(c0, d_m, u_l));
                    const uint32_t n1 = flags(far2(c1, row[-3 * 64 + 1], row[3 * 64 + 1]), far2(c1, __funnelshift_r(c1, rw, 24), __funnelshift_r(c0, c1, 8)),
                                              far2(c1, u_r, d_m), far2(c1, d_r, u_m));
                    m = n0 | (n1 << 4);
                }
                mk[r * 32 + B] = (uint8_t)m;
            }
            };
            if (tcur >= 127) reject(std::true_type{}); else reject(std::false_type{});
        }
        __syncthreads();
        // ---- 2. compaction of the surviving pixels into the queue (entry = tile row << 8 | tile column) ----------
        // a thread takes 32 pixels (one word of the row bitmaps): one warp scan per 1024 pixels
        {
            const uint32_t* M = reinterpret_cast<const uint32_t*>(mk);
            const uint32_t* vm = pass == 0 ? vmask : colmask;
            const int nwords = th * 8;
            for (int i0 = 0; i0 < nwords; i0 += kSegThreads) {
                const int i = i0 + t;
                uint32_t v = i < nwords ? (M[i] & vm[i & 7]) : 0u;   // only the tested columns (pass 1: of the cells that need it)
                const int cnt = __popc(v);
                int inc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += o;
                }
                int base = 0;
                if (lane == 31 && inc) base = smem_atomic_add(&qn, inc);
                base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
                const int e0 = i << 5;   // 8 words per tile row: row * 256 + 32 * (i & 7)
                while (v) {
                    const int b = __ffs(v) - 1;
                    v &= v - 1;
                    queue[base++] = (uint16_t)(e0 + b);
                }
            }
        }
        __syncthreads();
        const int nq = qn;   // written again only by "prepare pass 1", several barriers from here
        if (t == 0) nneed = 0;
        if (t < 8) need[t] = 0;
        // ---- 3. exact score of the queued pixels ------------------------------------------------------------------------------
        // A warp takes a contiguous share of the queue and moves the corners it finds to the front of that share (ballot + running
        // count in a register: no atomics, the write position never passes the read position), so the NMS below sees corners only.
        const int per = ((nq + kSegThreads - 1) / kSegThreads) * 32;   // queue entries per warp
        const int qlo = wi * per, qhi = min(qlo + per, nq);
        int ncorner = 0;                                                // corners of this warp (same in all lanes)
        for (int q0 = qlo; q0 < qhi; q0 += 32) {
            const int q = q0 + lane;
            bool corner = false;
            int e = 0;
            if (q < qhi) {
                e = queue[q];
                const uint8_t* p = tile + 3 * kSegPitch + e;
                const uint32_t v = p[0];
                const uint32_t bias = ((256u - v) << 16) | (256u + v);
                uint32_t E[16];
                E[0] = p[3 * kSegPitch] * 0xFFFFu + bias;  E[4] = p[3] * 0xFFFFu + bias;   // lo: 256 + d_k, hi: 256 - d_k
                E[8] = p[-3 * kSegPitch] * 0xFFFFu + bias; E[12] = p[-3] * 0xFFFFu + bias;
                E[1] = p[3 * kSegPitch + 1] * 0xFFFFu + bias;   E[2] = p[2 * kSegPitch + 2] * 0xFFFFu + bias;   E[3] = p[kSegPitch + 3] * 0xFFFFu + bias;
                E[5] = p[-kSegPitch + 3] * 0xFFFFu + bias;      E[6] = p[-2 * kSegPitch + 2] * 0xFFFFu + bias;  E[7] = p[-3 * kSegPitch + 1] * 0xFFFFu + bias;
                E[9] = p[-3 * kSegPitch - 1] * 0xFFFFu + bias;  E[10] = p[-2 * kSegPitch - 2] * 0xFFFFu + bias; E[11] = p[-kSegPitch - 3] * 0xFFFFu + bias;
                E[13] = p[kSegPitch - 3] * 0xFFFFu + bias;      E[14] = p[2 * kSegPitch - 2] * 0xFFFFu + bias;  E[15] = p[3 * kSegPitch - 1] * 0xFFFFu + bias;
                uint32_t m3[16], m9[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) m3[k] = __vimin3_u16x2(E[k], E[(k + 1) & 15], E[(k + 2) & 15]);
#pragma unroll
                for (int k = 0; k < 16; ++k) m9[k] = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                uint32_t a = __vimax3_u16x2(m9[0], m9[1], m9[2]), b = __vimax3_u16x2(m9[3], m9[4], m9[5]);
                uint32_t c = __vimax3_u16x2(m9[6], m9[7], m9[8]), d = __vimax3_u16x2(m9[9], m9[10], m9[11]);
                uint32_t f = __vimax3_u16x2(m9[12], m9[13], m9[14]);
                a = __vimax3_u16x2(a, b, c);
                d = __vimax3_u16x2(d, f, m9[15]);
                a = __vmaxu2(a, d);
                const int V = max((int)(a & 0xffffu), (int)(a >> 16)) - 257;
                if (V >= tcur) {
                    score[e + kSegPitch + 4 - ox] = (uint8_t)V;   // (row + 1) * 256 + px + 4
                    corner = true;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, corner);   // every lane has read its entry of this chunk by now
            if (corner) queue[qlo + ncorner + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)e;
            ncorner += __popc(bal);
        }
        __syncthreads();
        // ---- 4. NMS over the corners (each warp walks its own list) ----------------------------------------------------------------
        for (int q = qlo + lane; q < qlo + ncorner; q += 32) {
            const int e = queue[q], r = e >> 8, px = (e & 255) - ox;
            const uint8_t* s = score + e + kSegPitch + 4 - ox;
            const int v = s[0];
            const int fl = lut[px];
            // in pass 0 the score map only holds V >= iniTh; in pass 1 a cell's map holds everything >= minTh
            // all eight neighbours are loaded up front (the map has a zero margin all round); a cell's first / last column ignores the
            // neighbours that belong to the next cell
            const int n0 = s[-kSegPitch], n1 = s[kSegPitch], l0 = s[-1], l1 = s[-kSegPitch - 1], l2 = s[kSegPitch - 1];
            const int r0 = s[1], r1 = s[-kSegPitch + 1], r2 = s[kSegPitch + 1];
            const bool kl = (fl & 1) | ((v > l0) & (v > l1) & (v > l2)), kr = ((fl >> 1) & 1) | ((v > r0) & (v > r1) & (v > r2));
            const bool k = (v > n0) & (v > n1) & kl & kr;
            if (k) {
                // plain shared-memory reduction (atomicOr makes ptxas build a warp-aggregation loop that costs more than it saves)
                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(smem_u32(&bm_out[r * kBmWords + (px >> 5)])), "r"(1u << (px & 31)) : "memory");
            }
        }
        __syncthreads();
        if (pass == 1) break;
        // ---- which cells have no iniTh survivor? (one warp per cell) ----------------------------------------------
        for (int j = wi; j < sg.ncells; j += kSegThreads / 32) {
            const int cx = j * wcell, cw = P.cells[sg.first_cell + j].tw;
            const uint64_t wmask = cw >= 64 ? ~0ull : ((1ull << cw) - 1);
            uint64_t any = 0;
            for (int row = lane; row < th; row += 32) {
                const uint32_t* bw = bm_ini + row * kBmWords + (cx >> 5);
                const int sh = cx & 31;
                any |= (((uint64_t)__funnelshift_r(bw[1], bw[2], sh) << 32) | __funnelshift_r(bw[0], bw[1], sh)) & wmask;
            }
            if (!__any_sync(0xffffffffu, any != 0) && lane == 0) { need[j] = 1; smem_atomic_add(&nneed, 1); }
        }
        __syncthreads();
        if (nneed == 0) break;
        // ---- prepare pass 1: column mask of the needy cells, empty queue (the tile is still intact) ---------------------
        {
            const int px = t - ox;
            const bool on = px >= 0 && px < tw && need[(px * L.wcell_magic) >> 16];
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (lane == 0) colmask[wi] = bal;
            if (t == 0) qn = 0;
        }
        __syncthreads();
    }
    // ---- 5. per-cell threshold vote and ordered emission ------------------------------------------------------
    const int wc = L.wcell;
    for (int j = wi; j < sg.ncells; j += kSegThreads / 32) {
        const int ci = sg.first_cell + j;
        const Cell c = P.cells[ci];
        const int cx = j * wc, cw = c.tw;
        uint32_t* oxy = P.cand_xy + (long long)frame * P.total_cand_cap + L.cand_base + c.slot;
        uint8_t* orr = P.cand_resp + (long long)frame * P.total_cand_cap + L.cand_base + c.slot;
        if (th <= 32 && cw <= 32) {
            // the usual cell (about 30 x 30 tested pixels): one bitmap word per row, one row per lane
            const uint32_t wmask = cw >= 32 ? ~0u : ((1u << cw) - 1u);
            uint32_t mm = 0, mi = 0;
            if (lane < th) {
                const uint32_t* a = bm_min + lane * kBmWords + (cx >> 5);
                const uint32_t* b = bm_ini + lane * kBmWords + (cx >> 5);
                const int sh = cx & 31;
                mm = __funnelshift_r(a[0], a[1], sh) & wmask;
                mi = __funnelshift_r(b[0], b[1], sh) & wmask;
            }
            if (__any_sync(0xffffffffu, mi != 0)) mm = mi;
            const int c0 = __popc(mm);   // 3x3 maxima: at most 16 in a 32-pixel row
            int i0 = c0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o0 = __shfl_up_sync(0xffffffffu, i0, d);
                if (lane >= d) i0 += o0;
            }
            const int tot = __shfl_sync(0xffffffffu, i0, 31);
            int pos = i0 - c0;
            while (mm) {
                const int b = __ffs((int)mm) - 1;
                mm &= mm - 1;
                // coordinates relative to (minBorderX, minBorderY) = (16,16) as the reference stores them (:822-823)
                oxy[pos] = ((uint32_t)(c.y0 + lane - 16) << 16) | (uint32_t)(c.x0 + b - 16);
                orr[pos] = score[(lane + 1) * kSegPitch + cx + b + 4];
                ++pos;
            }
            if (lane == 0) P.cell_count[(long long)frame * P.total_cells + ci] = tot;
            continue;
        }
        const uint64_t wmask = cw >= 64 ? ~0ull : ((1ull << cw) - 1);
        uint64_t mm[2], mi[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = lane + 32 * h;
            mm[h] = 0; mi[h] = 0;
            if (row < th) {
                const uint32_t* a = bm_min + row * kBmWords + (cx >> 5);
                const uint32_t* b = bm_ini + row * kBmWords + (cx >> 5);
                const int sh = cx & 31;
                mm[h] = (((uint64_t)__funnelshift_r(a[1], a[2], sh) << 32) | __funnelshift_r(a[0], a[1], sh)) & wmask;
                mi[h] = (((uint64_t)__funnelshift_r(b[1], b[2], sh) << 32) | __funnelshift_r(b[0], b[1], sh)) & wmask;
            }
        }
        const bool any_ini = __any_sync(0xffffffffu, (mi[0] | mi[1]) != 0);
        if (any_ini) { mm[0] = mi[0]; mm[1] = mi[1]; }
        const int c0 = __popcll(mm[0]), c1 = __popcll(mm[1]);
        int i0 = c0, i1 = c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o0 = __shfl_up_sync(0xffffffffu, i0, d), o1 = __shfl_up_sync(0xffffffffu, i1, d);
            if (lane >= d) { i0 += o0; i1 += o1; }
        }
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = lane + 32 * h;
            int pos = h == 0 ? i0 - c0 : tot0 + i1 - c1;
            uint64_t m = mm[h];
            while (m) {
                const int b = __ffsll((long long)m) - 1;
                m &= m - 1;
                oxy[pos] = ((uint32_t)(c.y0 + row - 16) << 16) | (uint32_t)(c.x0 + b - 16);
                orr[pos] = score[(row + 1) * kSegPitch + cx + b + 4];
                ++pos;
            }
        }
        if (lane == 0) P.cell_count[(long long)frame * P.total_cells + ci] = tot0 + tot1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// GaussianBlur 7x7 (separable, fixed point, Appendix A.2).  BORDER_REFLECT_101 of the isolated clone (:1085-1086) reaches 3 px
// outside the level: the tiles at a level's edges mirror those pixels inside shared memory, the stencil itself has no
// border logic.  One CTA = 224 x 128 outputs of one (level, frame); the 256 x 134 input tile arrives by
// TMA (16-byte aligned origin: 16 spare columns on each side).  A thread owns one 32-bit word (4 columns) and walks
// down 32 rows: the horizontal pass is two DP4A per pixel on byte windows cut with funnel shifts, the vertical pass
// runs on a register window of row pairs (period 6: the row loop is rolled in steps of 6, so the window rotates by renaming).
// ------------------------------------------------------------------------------------------------------------
// kBlurBands bands of 32 rows, kBlurW / 4 = 56 threads each: 224 threads = 7 full warps (with 64 threads per band an eighth of the
// lanes had no column to work on)
constexpr int kBlurW = 224, kBlurBands = 4, kBlurH = 32 * kBlurBands, kBlurBox = kBlurH + 6, kBlurThreads = kBlurBands * (kBlurW / 4);

struct BlurTile {
    int16_t level, tx, ty, pad;
};

// no register cap: 53 registers = 4 CTAs per SM measured faster (0.80 ms per 1024 frames) than 48 (5 CTAs, 0.82) or 40 (6 CTAs, 0.90)
__global__ void __launch_bounds__(kBlurThreads) k_blur_tma(
    const __grid_constant__ ExtractParams P, const BlurTile* __restrict__ tiles,
                                                           const CUtensorMap* __restrict__ tmaps) {
    __shared__ __align__(1024) uint8_t tile[kBlurBox * 256];
    __shared__ uint64_t mbar;
    const BlurTile bt = tiles[blockIdx.x];
    const Level& L = P.lv[bt.level];
    const int frame = P.frame0 + blockIdx.y, t = threadIdx.x;
    if (t == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(&mbar, kBlurBox * 256);
        // tile column 16 = interior column 224*tx (global column kXPad + 224*tx), tile row 3 = interior row kBlurH*ty
        tma_load_3d(tile, tmaps + kMaxLevels + bt.level, kXPad - 16 + kBlurW * bt.tx, kEdge - 3 + kBlurH * bt.ty, frame, &mbar);
    }
    const int band = t / (kBlurW / 4), c = t - band * (kBlurW / 4);
    const int x = kBlurW * bt.tx + 4 * c;          // first of this thread's 4 interior columns
    const int ybase = kBlurH * bt.ty + 32 * band;  // first output row of this band
    mbar_wait(&mbar, 0);
    // BORDER_REFLECT_101 of the isolated level (:1085-1086) for the 3 pixels the taps reach outside it: written into the tile here,
    // so the kernel does not depend on the level's 19-px frame in HBM (which is materialised only on demand).  Columns first
    // (every tile row), then rows as whole tile rows, like copyMakeBorder.
    {
        const int ex_ = L.w - 1 - kBlurW * bt.tx, ey = L.h - 1 - kBlurH * bt.ty;   // last image column / row in tile-local output coordinates
        const bool fix_l = bt.tx == 0, fix_r = ex_ <= kBlurW + 2, fix_t = bt.ty == 0, fix_b = ey <= kBlurH + 2;
        if (fix_l || fix_r) {
            for (int r = t; r < kBlurBox; r += kBlurThreads) {
                uint8_t* row = tile + r * 256 + 16;
                if (fix_l) { row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3]; }
                if (fix_r) {
#pragma unroll
                    for (int k = 1; k <= 3; ++k)
                        if (16 + ex_ + k < 256) row[ex_ + k] = row[ex_ - k];
                }
            }
            __syncthreads();
        }
        if (fix_t || fix_b) {
            uint32_t* Tw = reinterpret_cast<uint32_t*>(tile);
            for (int i = t; i < 3 * 64; i += kBlurThreads) {
                const int k = i / 64 + 1, wd = i & 63;
                if (fix_t) Tw[(3 - k) * 64 + wd] = Tw[(3 + k) * 64 + wd];
                if (fix_b && 3 + ey + k < kBlurBox) Tw[(3 + ey + k) * 64 + wd] = Tw[(3 + ey - k) * 64 + wd];
            }
            __syncthreads();
        }
    }
    if (x >= L.w || ybase >= L.h) return;
    const uint32_t* T = reinterpret_cast<const uint32_t*>(tile) + (32 * band) * 64 + 4 + c;
    const uint32_t KLO = OG_G0 | (OG_G1 << 8) | (OG_G2 << 16) | (OG_G3 << 24), KHI = OG_G2 | (OG_G1 << 8) | (OG_G0 << 16);
    const size_t dpitch = (size_t)L.pitch;   // widened once: the row pointer advances by one 64-bit add per row
    uint8_t* dst = level_ptr(P.blur, L, frame) + (size_t)(kEdge + ybase) * dpitch + (size_t)(kXPad + x);
    const int nrows = min(32, L.h - ybase);
    // Vertical pass on row PAIRS: a horizontal sum is at most 255 * 256 < 2^16, so two vertically adjacent sums share a word
    // (pr[k] = row k | row k+1 << 16, one PRMT) and one DP2A applies two taps: 3 DP2A + 1 IMAD per pixel instead of 7 multiply-adds.
    // The pair window has period 6, so the row loop is rolled in steps of 6 rows (window slots stay compile-time): the fully
    // unrolled 38 rows were 37 KB of straight-line code that every warp ran through once, and instruction fetch was the kernel's
    // first stall reason (ncu: no_instruction 3.0 warps per issue).
    uint8_t* drow = dst;
    uint32_t pr[6][4], hprev[4] = {0u, 0u, 0u, 0u};
    const uint32_t K01 = OG_G0 | (OG_G1 << 8), K23 = OG_G2 | (OG_G3 << 8), K21 = OG_G2 | (OG_G1 << 8);
    auto hpass = [&](const uint32_t* R, uint32_t (&h)[4]) {   // horizontal pass of one box row
        const uint32_t pw = R[-1], cw = R[0], nw = R[1];
        const uint32_t m3 = __funnelshift_r(pw, cw, 8), m2 = __funnelshift_r(pw, cw, 16), m1 = __funnelshift_r(pw, cw, 24);
        const uint32_t p1 = __funnelshift_r(cw, nw, 8), p2 = __funnelshift_r(cw, nw, 16), p3 = __funnelshift_r(cw, nw, 24);
        h[0] = __dp4a(m3, KLO, __dp4a(p1, KHI, 0u));
        h[1] = __dp4a(m2, KLO, __dp4a(p2, KHI, 0u));
        h[2] = __dp4a(m1, KLO, __dp4a(p3, KHI, 0u));
        h[3] = __dp4a(cw, KLO, __dp4a(nw, KHI, 0u));
    };
    // box rows 0..5: fill the window (pairs 0..4)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        uint32_t h[4];
        hpass(T + i * 64, h);
        if (i >= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pr[i - 1][j] = __byte_perm(hprev[j], h[j], 0x5410);   // rows i-1, i
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) hprev[j] = h[j];
    }
    // box rows 6 + 6 g + k produce output rows r = 6 g + k: window rows r..r+6 = pairs r, r+2, r+4 (slots k, k+2, k+4 mod 6) and row r+6
    const uint32_t* R = T + 6 * 64;
    auto row = [&](auto k_c) {   // box row R -> output row 6 g + k
        constexpr int k = decltype(k_c)::value;
        uint32_t h[4];
        hpass(R, h);
        R += 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[(k + 5) % 6][j] = __byte_perm(hprev[j], h[j], 0x5410);   // rows r+5, r+6
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t acc = OG_G0 * h[j] + 32768u;
            acc = __dp2a_lo(pr[k][j], K01, acc);
            acc = __dp2a_lo(pr[(k + 2) % 6][j], K23, acc);
            acc = __dp2a_lo(pr[(k + 4) % 6][j], K21, acc);
            o[j] = acc;   // result byte = bits 16..23
        }
        const uint32_t lo = __byte_perm(o[0], o[1], 0x0062), hi = __byte_perm(o[2], o[3], 0x0062);
        *reinterpret_cast<uint32_t*>(drow) = __byte_perm(lo, hi, 0x5410);
        drow += dpitch;
#pragma unroll
        for (int j = 0; j < 4; ++j) hprev[j] = h[j];
    };
    int r0 = 0;
#pragma unroll 1
    for (; r0 + 6 <= nrows; r0 += 6) {   // whole groups: no per-row checks
        row(std::integral_constant<int, 0>{}); row(std::integral_constant<int, 1>{}); row(std::integral_constant<int, 2>{});
        row(std::integral_constant<int, 3>{}); row(std::integral_constant<int, 4>{}); row(std::integral_constant<int, 5>{});
    }
    // the last 32 - 30 = 2 rows of a full band, or what a bottom tile has left before the level's last row
    const int left = nrows - r0;
    if (left > 0) row(std::integral_constant<int, 0>{});
    if (left > 1) row(std::integral_constant<int, 1>{});
    if (left > 2) row(std::integral_constant<int, 2>{});
    if (left > 3) row(std::integral_constant<int, 3>{});
    if (left > 4) row(std::integral_constant<int, 4>{});
}

// ------------------------------------------------------------------------------------------------------------
// Orientation + descriptor + final KeyPoint record: one warp per keypoint, kDescPerWarp keypoints per warp.
//   IC_Angle (:77-104): integer moments over the radius-15 disc of the UN-blurred level.  The disc rows are read as
//   aligned 32-bit words (9 per row); a host-built table gives, per (alignment, row, word), the four signed column
//   weights u (0 outside the disc) and the four row weights v, so one word contributes with two DP4A:
//   m10 += dp4a(pixels, u weights), m01 += dp4a(pixels, v weights).  Then cv::fastAtan2's polynomial.
//   computeOrbDescriptor (:108-147): the 37 x 37 neighbourhood of the BLURRED level is staged in shared memory with
//   coalesced word loads; lane i produces descriptor byte i from pattern points 16i .. 16i+15.  cos/sin of the float
//   angle are evaluated in double and rounded to float.
// ------------------------------------------------------------------------------------------------------------
#ifndef OG_DESC_WARPS
#define OG_DESC_WARPS 6
#define OG_DESC_PER_WARP 16
#define OG_DESC_MINB 6
#endif
constexpr int kDescWarps = OG_DESC_WARPS, kDescPerWarp = OG_DESC_PER_WARP;   // measured on B200 (device / host-path k frames/s): 8x4 96.3 / 82.2, 4x16 96.9 / 86.9, 6x16 97.8 / 87.1
constexpr int kIcWords = 9, kIcRows = 31;           // table [4 alignments][31 rows][9 words][2]
constexpr int kIcBoxW = 48, kBlurBoxW = 64, kPatchRows = 37;   // TMA boxes: 48 x 31 of the level, 64 x 37 of its blur
#ifndef OG_DESC_SWIZZLE
#define OG_DESC_SWIZZLE 1   // 0: unswizzled windows (development; 1.35 vs 1.23 ms per 1024 frames)
#endif
// box bytes rounded up to the 128 B a TMA destination wants (a 64-byte-swizzled destination: to the 512 B its pattern spans)
constexpr int kIcSlot = 1536, kBlurSlot = OG_DESC_SWIZZLE ? 2560 : 2432;
__constant__ int8_t c_pat_x[512] = {ORB_PATTERN_X_INIT};
__constant__ int8_t c_pat_y[512] = {ORB_PATTERN_Y_INIT};

// cvRound(x * b + y * a) and cvRound(x * a - y * b) of computeOrbDescriptor (:115-120) as the raw bits of round_rn_small's sum
// (0x4B400000 + the rounded value); one byte from a shared-memory address.
__device__ __forceinline__ uint32_t brief_row_bits(float fx, float fy, float a, float b) {
    return __float_as_uint(__fadd_rn(fadd(fmul(fx, b), fmul(fy, a)), 12582912.f));
}
__device__ __forceinline__ uint32_t brief_col_bits(float fx, float fy, float a, float b) {
    return __float_as_uint(__fadd_rn(fsub(fmul(fx, a), fmul(fy, b)), 12582912.f));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {   // unsigned bytes of a times signed bytes of b
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Both neighbourhoods of a key point arrive by TMA (one elected lane, double-buffered boxes, one mbarrier per buffer and warp):
// the 31 x 31 disc of the level for IC_Angle and the 37 x 37 window of the blurred level for the descriptor.  A box starts at a
// 16-byte aligned column at or left of the window (TMA's alignment rule), so it is 48 / 64 bytes wide.  tmaps: [2K + l] level
// boxes, [3K + l] blur boxes (K = kMaxLevels).
// A warp owns kDescPerWarp key points and works in phases, so that everything that is ONE value per key point is computed by
// one lane per key point instead of by all 32 lanes for one key point at a time:
//   0. lane n looks up key point n (level, position, response);
//   1. moments: the discs stream through two buffers, all lanes sum one disc, lane n keeps m10 / m01 of key point n;
//   2. lane n: fastAtan2, cos / sin, the cv::KeyPoint record;
//   3. descriptors: the blurred windows stream through the same two buffers, cos / sin of key point n come by shuffle.
__global__ void __launch_bounds__(kDescWarps * 32, OG_DESC_MINB) k_orient_desc(const __grid_constant__ ExtractParams P, const CUtensorMap* __restrict__ tmaps,
                                                                    KeyPoint* __restrict__ kp_out, uint8_t* __restrict__ desc_out,
                                                                    int32_t* __restrict__ counts) {
    static_assert(kDescPerWarp <= 32 && 2 * kIcSlot <= 2 * kBlurSlot, "one lane per key point; the disc buffers fit into the window buffers");
    // pattern point j of lane i (= bit_pattern_31_ point 16 i + j) at spat[j * 32 + i]: staged once per CTA so that every lane
    // picks up its 16 points with conflict-free loads (a lane-indexed read of the __constant__ table would serialise)
    __shared__ __half2 spat[512];
    __shared__ __align__(1024) uint8_t patch[kDescWarps][2 * kBlurSlot];
    __shared__ uint64_t bars[kDescWarps][2];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) spat[(i & 15) * 32 + (i >> 4)] = __floats2half2_rn((float)c_pat_x[i], (float)c_pat_y[i]);
    const int frame = P.frame0 + blockIdx.y;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    if (lane == 0) {
        mbar_init(&bars[wi][0], 1);
        mbar_init(&bars[wi][1], 1);
    }
    __syncthreads();
    // keypoints are concatenated in level order (:1076-1103): level starts from the per-level counts
    const int32_t* sc = P.sel_count + frame * P.n_levels;
    int lstart[kMaxLevels + 1];
    lstart[0] = 0;
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) lstart[l + 1] = lstart[l] + (l < P.n_levels ? sc[l] : 0);
    const int total = min(lstart[kMaxLevels], P.kp_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[frame] = total;
    const unsigned full = 0xffffffffu;

    // ---- 0. lane n: key point n of this warp -------------------------------------------------------------------------------
    const int my_idx = (blockIdx.x * kDescPerWarp + lane) * kDescWarps + wi;   // index among the frame's keypoints
    const bool mine = lane < kDescPerWarp && my_idx < total;
    int my_level = 0, my_cx = 0, my_cy = 0, my_resp = 0;
    uint32_t my_xy = 0;   // my_cy << 16 | my_cx: one shuffle hands both to the warp
    if (mine) {
        int first = 0;
#pragma unroll
        for (int l = 1; l < kMaxLevels; ++l)
            if (my_idx >= lstart[l]) { my_level = l; first = lstart[l]; }
        const Level& L = P.lv[my_level];
        const long long o = (long long)frame * P.total_sel_cap + L.sel_base + (my_idx - first);
        const uint32_t xy = P.sel_xy[o];
        my_resp = P.sel_resp[o];
        my_cx = (int)(xy & 0xffffu) + 16;   // level coordinates (:840-841)
        my_cy = (int)(xy >> 16) + 16;
        my_xy = ((uint32_t)my_cy << 16) | (uint32_t)my_cx;
    }
    const int nk = __popc(__ballot_sync(full, mine));   // the warp's key points are lanes 0 .. nk-1
    if (nk == 0) return;
    uint8_t* buf0 = patch[wi];
    uint32_t use0 = 0, use1 = 0;   // completed phases of the two barriers

    // ---- 1. IC_Angle moments (:77-104) ------------------------------------------------------------------------------------------
    const int ic_r = lane / kIcWords, ic_j = lane - ic_r * kIcWords;     // lanes 0..26: 3 disc rows x 9 words per step
    // issue_*(n) returns key point n's packed position, which the loop keeps for iteration n (one position shuffle per key point)
    auto issue_ic = [&](int n) {
        const uint32_t xy = __shfl_sync(full, my_xy, n);
        const int lv = __shfl_sync(full, my_level, n), cx = (int)(xy & 0xffffu), cy = (int)(xy >> 16);
        if (lane == 0) {
            uint64_t* bar = &bars[wi][n & 1];
            mbar_expect_tx(bar, kIcBoxW * kIcRows);
            tma_load_3d(buf0 + (n & 1) * kIcSlot, tmaps + 2 * kMaxLevels + lv, (kXPad + cx - kHalfPatch) & ~15, kEdge + cy - kHalfPatch, frame, bar);
        }
        return xy;
    };
    int my_m10 = 0, my_m01 = 0;
    uint32_t xy_next = issue_ic(0);
    for (int n = 0; n < nk; ++n) {
        const uint32_t xy_cur = xy_next;
        if (n + 1 < nk) xy_next = issue_ic(n + 1);   // its buffer was read two iterations ago (the __syncwarp below)
        const int icx = kXPad + (int)(xy_cur & 0xffffu) - kHalfPatch;   // first column of the disc
        int m10 = 0, m01 = 0;
        if (n & 1) { mbar_wait(&bars[wi][1], use1 & 1); ++use1; } else { mbar_wait(&bars[wi][0], use0 & 1); ++use0; }
        {
            const int A = icx & 3;
            const uint32_t* base = reinterpret_cast<const uint32_t*>(buf0 + (n & 1) * kIcSlot + ((icx & 15) & ~3)) + ic_j;   // aligned word holding the disc's first column
            const uint2* tab = reinterpret_cast<const uint2*>(P.ic_tab) + (A * kIcRows) * kIcWords + ic_j;
            if (lane < 3 * kIcWords) {
#pragma unroll
                for (int it = 0; it < (kIcRows + 2) / 3; ++it) {
                    const int row = 3 * it + ic_r;
                    if (row < kIcRows) {
                        const uint32_t px = base[row * (kIcBoxW / 4)];
                        const uint2 w = __ldg(tab + row * kIcWords);
                        m10 = dp4a_us(px, w.x, m10);
                        m01 = dp4a_us(px, w.y, m01);
                    }
                }
            }
        }
        // warp sums in one instruction each (REDUX) instead of five shuffle + add steps
        m10 = __reduce_add_sync(full, m10);
        m01 = __reduce_add_sync(full, m01);
        if (lane == n) { my_m10 = m10; my_m01 = m01; }
        __syncwarp();   // every lane is done with this buffer
    }

    // ---- 2. lane n: orientation, cos / sin, the KeyPoint record (:837-847, :1095-1103) ----------------------------------------------------
    float my_a = 0.f, my_b = 0.f;
    if (mine) {
        const float angle = fast_atan2((float)my_m01, (float)my_m10, P.atan);
        sincosf_glibc(fmul(angle, P.factor_pi), &my_b, &my_a);   // a = cos, b = sin, as glibc's sincosf returns them to the reference (:113)
        const Level& L = P.lv[my_level];
        KeyPoint kp;
        const float fx = (float)my_cx, fy = (float)my_cy;
        kp.x = my_level ? fmul(fx, L.scale) : fx;   // keypoint->pt *= scale for level != 0 (:1095-1101)
        kp.y = my_level ? fmul(fy, L.scale) : fy;
        kp.size = L.kp_size;
        kp.angle = angle;
        kp.response = (float)my_resp;
        kp.octave = my_level;
        kp.class_id = -1;
        kp_out[(long long)frame * P.kp_cap + my_idx] = kp;
    }

    // ---- 3. rotated BRIEF on the blurred windows (:108-147) ----------------------------------------------------------------------------------
    auto issue_bl = [&](int n) {
        const uint32_t xy = __shfl_sync(full, my_xy, n);
        const int lv = __shfl_sync(full, my_level, n), cx = (int)(xy & 0xffffu), cy = (int)(xy >> 16);
        if (lane == 0) {
            uint64_t* bar = &bars[wi][n & 1];
            mbar_expect_tx(bar, kBlurBoxW * kPatchRows);
            tma_load_3d(buf0 + (n & 1) * kBlurSlot, tmaps + 3 * kMaxLevels + lv, (kXPad + cx - 18) & ~15, kEdge + cy - 18, frame, bar);
        }
        return xy;
    };
    xy_next = issue_bl(0);
    // The lane's 16 pattern points are the same for every key point of the warp: they live in 16 registers as half pairs
    // (|coordinate| <= 15 is exact in half precision; two conversions per point in the loop) instead of being re-read from shared
    // memory per key point, which leaves the shared-memory pipe to the byte gathers: 1.22 -> 1.17 ms per 1024 frames at the same
    // 56 registers / 6 CTAs per SM (as float pairs they need 64 registers / 5 CTAs: 1.18 ms; 80 / 4 CTAs: 1.24 ms).
    __half2 preg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) preg[j] = spat[j * 32 + lane];
    for (int n = 0; n < nk; ++n) {
        const uint32_t xy_cur = xy_next;
        if (n + 1 < nk) xy_next = issue_bl(n + 1);
        const float a = __shfl_sync(full, my_a, n), b = __shfl_sync(full, my_b, n);
        const int blx = kXPad + (int)(xy_cur & 0xffffu) - 18;   // first column of the 37-wide window
        if (n & 1) { mbar_wait(&bars[wi][1], use1 & 1); ++use1; } else { mbar_wait(&bars[wi][0], use0 & 1); ++use0; }
#if OG_DESC_SWIZZLE
        // 64-byte swizzle: the 16-byte chunk index (address bits 4-5) of a box row is XORed with bits 7-8 (= row / 2 mod 4), so the
        // gathers of a warp spread over all 16 banks of their row parity instead of the 9 the window's columns cover
        // The sample address is formed on the raw bits of the rounding (round_rn_small: float + 1.5 * 2^23 holds the integer in its
        // low mantissa bits, C = 0x4B400000 more than it): (ir - C) * 64 + (iq - C) + window centre + buffer address
        // = ir * 64 + (iq + kabs), so one add and one shift-add per point instead of five integer instructions.  The buffers
        // start on multiples of 512 bytes, so the swizzle bits of the absolute address are those of the offset.
        static_assert(kBlurSlot % 512 == 0 && (2 * kBlurSlot) % 1024 == 0, "swizzle bits of the absolute shared-memory address");
        const uint32_t kabs = smem_u32(buf0 + (n & 1) * kBlurSlot) + (uint32_t)(18 * kBlurBoxW + 18 + (blx & 15)) - 65u * 0x4B400000u;
#else
        const uint8_t* b0 = buf0 + (n & 1) * kBlurSlot + 18 * kBlurBoxW + 18 + (blx & 15);
#endif
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 p0 = __half22float2(preg[2 * k]), p1 = __half22float2(preg[2 * k + 1]);
#if OG_DESC_SWIZZLE
            static_assert(kBlurBoxW == 64, "row pitch of the window = the shift below");
            const uint32_t a0 = brief_row_bits(p0.x, p0.y, a, b) * 64u + (brief_col_bits(p0.x, p0.y, a, b) + kabs);
            const uint32_t a1 = brief_row_bits(p1.x, p1.y, a, b) * 64u + (brief_col_bits(p1.x, p1.y, a, b) + kabs);
            const uint32_t t0 = lds_u8(a0 ^ ((a0 >> 3) & 0x30u)), t1 = lds_u8(a1 ^ ((a1 >> 3) & 0x30u));
#else
            int r0, q0, r1, q1;
            brief_offset_f(p0.x, p0.y, a, b, &r0, &q0);
            brief_offset_f(p1.x, p1.y, a, b, &r1, &q1);
            const int t0 = b0[r0 * kBlurBoxW + q0], t1 = b0[r1 * kBlurBoxW + q1];
#endif
            val |= (t0 < t1) << k;
        }
        const int idx = (blockIdx.x * kDescPerWarp + n) * kDescWarps + wi;
        desc_out[((long long)frame * P.kp_cap + idx) * 32 + lane] = (uint8_t)val;
        __syncwarp();   // every lane is done with this buffer
    }
}

}  // namespace og

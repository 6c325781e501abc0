// og_stereo.cuh — Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:501-675) on the device, fed directly by the
// key points, descriptors and pyramids the left and right extractor instances left in HBM (no pyramid download).
//
//   k_stereo_rows     vRowIndices (:508-523): CSR of the right key points per level-0 image row (band of +-2*scale rows)
//   k_stereo_match    one warp per left key point: best right candidate by Hamming distance among the row's candidates
//                     (octave +-1, uR in [uL - maxD, uL]; first minimum = smallest right index), then the 11 x 11 SAD over
//                     11 offsets on the level images, parabola fit, disparity / depth (:532-655)
//   k_stereo_filter   per frame: median of the accepted SAD values, matches with SAD >= 1.5*1.4*median removed (:659-674)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "og_match.cuh"
#include "og_types.h"

namespace og {

struct StereoArgs {
    ExtractParams PL, PR;          // geometry + pyramid base of the two extractors (same geometry)
    const KeyPoint *kpL, *kpR;     // [batch][kp_stride]
    const uint8_t *descL, *descR;  // [batch][kp_stride][32]
    const int32_t *cntL, *cntR;    // [batch]
    int kp_stride;
    int n_rows;                    // rows of level 0
    float max_d, mbf;              // mbf / mb, mbf
    float inv_scale[kMaxLevels];
    int32_t* row_start;            // [batch][n_rows + 1]
    int32_t* row_items;            // [batch][items_cap]
    int items_cap;
    float *u_right, *depth;        // [batch][out_stride]
    int32_t* sad;                  // [batch][kp_stride] scratch: accepted SAD or -1
    int out_stride;
};

constexpr int kStereoThreads = 256;

__global__ void __launch_bounds__(kStereoThreads) k_stereo_rows(const __grid_constant__ StereoArgs A) {
    extern __shared__ int s_cnt[];   // [n_rows + 1]
    __shared__ int wsum[kStereoThreads / 32];
    const int f = blockIdx.x, t = threadIdx.x;
    const int nr = A.cntR[f];
    const KeyPoint* kp = A.kpR + (long long)f * A.kp_stride;
    for (int i = t; i <= A.n_rows; i += kStereoThreads) s_cnt[i] = 0;
    __syncthreads();
    auto band = [&](int i, int& lo, int& hi) {
        const float y = kp[i].y, r = __fmul_rn(2.0f, A.PR.lv[kp[i].octave].scale);   // :516-519
        hi = min((int)ceilf(__fadd_rn(y, r)), A.n_rows - 1);
        lo = max((int)floorf(__fsub_rn(y, r)), 0);
    };
    for (int i = t; i < nr; i += kStereoThreads) {
        int lo, hi;
        band(i, lo, hi);
        for (int y = lo; y <= hi; ++y) atomicAdd(&s_cnt[y], 1);
    }
    __syncthreads();
    // exclusive scan over the rows: contiguous chunk per thread
    const int chunk = (A.n_rows + kStereoThreads - 1) / kStereoThreads;
    const int lo = min(t * chunk, A.n_rows), hi = min(lo + chunk, A.n_rows);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += s_cnt[i];
    const int lane = t & 31, w = t >> 5;
    int v = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    int run = v - s;
    for (int k = 0; k < w; ++k) run += wsum[k];
    int32_t* rs = A.row_start + (long long)f * (A.n_rows + 1);
    for (int i = lo; i < hi; ++i) {
        const int c = s_cnt[i];
        rs[i] = run;
        s_cnt[i] = run;   // fill cursor
        run += c;
    }
    if (t == kStereoThreads - 1) rs[A.n_rows] = run;
    __syncthreads();
    int32_t* items = A.row_items + (long long)f * A.items_cap;
    for (int i = t; i < nr; i += kStereoThreads) {
        int blo, bhi;
        band(i, blo, bhi);
        for (int y = blo; y <= bhi; ++y) {
            const int pos = atomicAdd(&s_cnt[y], 1);
            if (pos < A.items_cap) items[pos] = i;   // order inside a row is irrelevant: the match key carries the index
        }
    }
}

__global__ void __launch_bounds__(kStereoThreads) k_stereo_match(const __grid_constant__ StereoArgs A) {
    __shared__ int s_sad[kStereoThreads / 32][12];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int iL = blockIdx.x * (kStereoThreads / 32) + wi;
    const int nl = A.cntL[f];
    if (iL >= nl) {
        // slots beyond the frame's key points read -1 as well (the host variant copies whole rows)
        if (lane == 0 && iL < A.kp_stride) {
            A.u_right[(long long)f * A.out_stride + iL] = -1.0f;
            A.depth[(long long)f * A.out_stride + iL] = -1.0f;
        }
        return;
    }
    const KeyPoint kL = A.kpL[(long long)f * A.kp_stride + iL];
    float out_u = -1.0f, out_d = -1.0f;
    int out_sad = -1;
    const int row = (int)kL.y;
    const float minU = __fsub_rn(kL.x, A.max_d), maxU = kL.x;   // minD = 0 (:526, :547-548)
    uint32_t best = kEmptyKey;
    if (row >= 0 && row < A.n_rows && !(maxU < 0.f)) {
        const int32_t* rs = A.row_start + (long long)f * (A.n_rows + 1);
        const int c0 = rs[row], c1 = min(rs[row + 1], A.items_cap);
        const int32_t* items = A.row_items + (long long)f * A.items_cap;
        const KeyPoint* kR = A.kpR + (long long)f * A.kp_stride;
        const Desc dq = load_desc(A.descL, (long long)f * A.kp_stride + iL);
        for (int c = c0 + lane; c < c1; c += 32) {
            const int iR = items[c];
            const KeyPoint k = kR[iR];
            if (k.octave < kL.octave - 1 || k.octave > kL.octave + 1) continue;
            if (!(k.x >= minU && k.x <= maxU)) continue;
            const int dist = hamming256(dq, load_desc(A.descR, (long long)f * A.kp_stride + iR));
            best = min(best, ((uint32_t)dist << kPosBits) | (uint32_t)iR);   // strict '<' over ascending iR = smallest index among ties
        }
        best = __reduce_min_sync(0xffffffffu, best);
    }
    const int bestDist = best == kEmptyKey ? ORBGPU_TH_HIGH : (int)(best >> kPosBits);
    if (bestDist < ORBGPU_TH_HIGH && bestDist < (ORBGPU_TH_HIGH + ORBGPU_TH_LOW) / 2) {
        const int bestR = (int)(best & kPosMask);
        const float uR0 = A.kpR[(long long)f * A.kp_stride + bestR].x;
        const float sfac = A.inv_scale[kL.octave];
        const int su = (int)roundf(__fmul_rn(kL.x, sfac)), sv = (int)roundf(__fmul_rn(kL.y, sfac)), sr = (int)roundf(__fmul_rn(uR0, sfac));
        const Level& L = A.PL.lv[kL.octave];
        const int w = 5, LL = 5;
        // iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1 (:597-600)
        if (!(sr + LL - w < 0 || sr + LL + w + 1 >= L.w)) {
            const uint8_t* IL = A.PL.pyr + L.base + (long long)(A.PL.frame0 + f) * L.frame_stride + (long long)(kEdge + sv) * L.pitch + kXPad + su;
            const Level& LR = A.PR.lv[kL.octave];
            const uint8_t* IR = A.PR.pyr + LR.base + (long long)(A.PR.frame0 + f) * LR.frame_stride + (long long)(kEdge + sv) * LR.pitch + kXPad + sr;
            if (lane < 11) s_sad[wi][lane] = 0;
            __syncwarp();
            const int cL = IL[0];
            // 121 (row, offset) pairs over the lanes; every pair sums 11 columns.  Integer arithmetic is exact (the reference's
            // float patches hold integers in [-255, 255], cv::norm accumulates in double)
            for (int p = lane; p < 121; p += 32) {
                const int dy = p / 11 - w, inc = p % 11 - LL;
                const int cR = IR[inc];
                const uint8_t* a = IL + (long long)dy * L.pitch;
                const uint8_t* b = IR + (long long)dy * LR.pitch + inc;
                int acc = 0;
#pragma unroll
                for (int dx = -5; dx <= 5; ++dx) acc += abs(((int)a[dx] - cL) - ((int)b[dx] - cR));
                atomicAdd(&s_sad[wi][inc + LL], acc);
            }
            __syncwarp();
            if (lane == 0) {
                int bestSad = 0x7fffffff, bestinc = 0;
                for (int k = 0; k < 11; ++k)
                    if (s_sad[wi][k] < bestSad) { bestSad = s_sad[wi][k]; bestinc = k - LL; }
                if (!(bestinc == -LL || bestinc == LL)) {
                    const float d1 = (float)s_sad[wi][LL + bestinc - 1], d2 = (float)s_sad[wi][LL + bestinc], d3 = (float)s_sad[wi][LL + bestinc + 1];
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                    if (!(deltaR < -1.f || deltaR > 1.f)) {
                        float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn((float)sr, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(kL.x, bestuR);
                        if (disparity >= 0.f && disparity < A.max_d) {
                            if (disparity <= 0.f) {
                                disparity = 0.01f;
                                bestuR = (float)((double)kL.x - 0.01);
                            }
                            out_d = __fdiv_rn(A.mbf, disparity);
                            out_u = bestuR;
                            out_sad = bestSad;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) {
        A.u_right[(long long)f * A.out_stride + iL] = out_u;
        A.depth[(long long)f * A.out_stride + iL] = out_d;
        A.sad[(long long)f * A.kp_stride + iL] = out_sad;
    }
}

// median of the accepted (SAD, index) pairs in the order of std::sort on pair<int,int>, then the removal sweep
__global__ void __launch_bounds__(kStereoThreads) k_stereo_filter(const __grid_constant__ StereoArgs A) {
    __shared__ int s_n, s_med;
    __shared__ int wtot[kStereoThreads / 32];
    const int f = blockIdx.x, t = threadIdx.x;
    const int nl = A.cntL[f];
    const int32_t* sad = A.sad + (long long)f * A.kp_stride;
    // number of accepted matches
    int c = 0;
    for (int i = t; i < nl; i += kStereoThreads) c += sad[i] >= 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((t & 31) == 0) wtot[t >> 5] = c;
    __syncthreads();
    if (t == 0) {
        int n = 0;
        for (int k = 0; k < kStereoThreads / 32; ++k) n += wtot[k];
        s_n = n;
        s_med = -1;
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;   // the reference indexes an empty vector here (undefined); nothing to remove
    // vDistIdx[size / 2].first after std::sort (:659-661): only the VALUE of the element of rank n/2 matters, i.e. the largest v with
    // |{sad < v}| <= n/2 — found bit by bit with block-wide counts (17 rounds of n/256 compares per thread instead of the n^2
    // rank counting this kernel started with: 0.52 ms -> a few microseconds per 256 frames).
    const int target = n / 2;
    int v = 0;
    for (int bit = 16; bit >= 0; --bit) {   // SAD <= 121 * 2 * 255 < 2^17
        const int cand = v + (1 << bit);
        int below = 0;
        for (int i = t; i < nl; i += kStereoThreads) {
            const int si = sad[i];
            below += si >= 0 && si < cand;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) below += __shfl_xor_sync(0xffffffffu, below, d);
        __syncthreads();   // wtot of the previous round has been read by everyone
        if ((t & 31) == 0) wtot[t >> 5] = below;
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int k = 0; k < kStereoThreads / 32; ++k) tot += wtot[k];
        if (tot <= target) v = cand;
    }
    if (t == 0) s_med = v;
    __syncthreads();
    const float th = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_med);
    for (int i = t; i < nl; i += kStereoThreads) {
        if (sad[i] >= 0 && !((float)sad[i] < th)) {
            A.u_right[(long long)f * A.out_stride + i] = -1.0f;
            A.depth[(long long)f * A.out_stride + i] = -1.0f;
        }
    }
}

}  // namespace og

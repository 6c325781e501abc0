// og_match.cuh — device code of the Hamming path of ORBmatcher (ORBmatcher.cc), sm_100a.
//
// Every search function of the reference is "for each query, scan a candidate list, keep best / second best,
// apply an acceptance rule, and (in most of them) mark the winner as taken so later queries skip it".  The
// distance evaluations are independent; only the "taken" state is sequential.  So every search is split in two:
//
//   phase A (parallel, all the popcount work): for every query the K smallest candidates by (distance, scan
//           position) among the statically eligible ones                  -> top-K lists in HBM
//   phase B (one warp per frame / frame pair, walks the queries in the reference's order): best and second best
//           = the first two entries of the list that are not taken yet; if the list is exhausted before two are
//           found and it was truncated, the warp rescans that query's candidates with the taken mask (exact).
//
// Keys are dist << 20 | position, so an unsigned minimum is "smallest distance, first encountered", the
// reference's strict `<` update rule (Appendix C of SURVEY.md).  SearchForTriangulation has no taken state and
// lets equal distances replace (:882), so its key is dist << 20 | (0xFFFFF - position) and needs no phase B walk.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "og_types.h"

namespace og {

constexpr int kTopK = 4;
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;
constexpr int kPosBits = 20;
constexpr uint32_t kPosMask = (1u << kPosBits) - 1;
constexpr int kGridCols = 64, kGridRows = 48;  // FRAME_GRID_COLS / FRAME_GRID_ROWS, Frame.h:37-38
constexpr int kHisto = 30;                     // HISTO_LENGTH, ORBmatcher.cc:39

// Device view of a frame set (see orbgpu_frame_set in include/orbgpu.h).
struct FrameSetView {
    const int32_t* kp_off;
    const KeyPoint* keys;
    const uint8_t* desc;
    const float* u_right;   // may be null
    const uint8_t* flags;   // may be null
    const float* grid;      // may be null
    const int32_t* node_off;
    const int32_t* node_id;
    const int32_t* feat_off;
    const int32_t* feat;
};

struct MapPointView {
    const int32_t* mp_off;
    const float *proj_x, *proj_y, *proj_xr, *view_cos;
    const int32_t* level;
    const uint8_t* flags;
    const uint8_t* desc;
};

struct Desc {
    uint4 lo, hi;
};

__device__ __forceinline__ Desc load_desc(const uint8_t* base, long long row) {
    const uint4* p = reinterpret_cast<const uint4*>(base + row * 32);
    Desc d;
    d.lo = __ldg(p);
    d.hi = __ldg(p + 1);
    return d;
}

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1901-1917): the reference's SWAR bit count over eight 32-bit
// words is a population count; POPC does it in one instruction per word.
__device__ __forceinline__ int hamming256(const Desc& a, const Desc& b) {
    return __popc(a.lo.x ^ b.lo.x) + __popc(a.lo.y ^ b.lo.y) + __popc(a.lo.z ^ b.lo.z) + __popc(a.lo.w ^ b.lo.w) +
           __popc(a.hi.x ^ b.hi.x) + __popc(a.hi.y ^ b.hi.y) + __popc(a.hi.z ^ b.hi.z) + __popc(a.hi.w ^ b.hi.w);
}

// The same distance with half the POPCs: POPC issues on the XU pipe at 16 lanes/clk/SM and binds the brute-force scan, the
// LOP3 pipe has room.  Carry-save adders (sum = a^b^c, carry = maj(a,b,c), one LOP3 each) compress the eight XOR words
// into one word of weight 1 twice (s2, x7), one of weight 2 (s3) and one of weight 4 (c3): 4 POPC + 16 LOP3 instead of 8 + 8.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int hamming256_csa(const Desc& a, const Desc& b) {
    const uint32_t x0 = a.lo.x ^ b.lo.x, x1 = a.lo.y ^ b.lo.y, x2 = a.lo.z ^ b.lo.z, x3 = a.lo.w ^ b.lo.w;
    const uint32_t x4 = a.hi.x ^ b.hi.x, x5 = a.hi.y ^ b.hi.y, x6 = a.hi.z ^ b.hi.z, x7 = a.hi.w ^ b.hi.w;
    const uint32_t s0 = lop3_xor3(x0, x1, x2), c0 = lop3_maj(x0, x1, x2);
    const uint32_t s1 = lop3_xor3(x3, x4, x5), c1 = lop3_maj(x3, x4, x5);
    const uint32_t s2 = lop3_xor3(s0, s1, x6), c2 = lop3_maj(s0, s1, x6);
    const uint32_t s3 = lop3_xor3(c0, c1, c2), c3 = lop3_maj(c0, c1, c2);
    return __popc(s2) + __popc(x7) + 2 * __popc(s3) + 4 * __popc(c3);
}

// sorted insertion into t[0] <= t[1] <= t[2] <= t[3]; precondition key < t[3]
__device__ __forceinline__ void topk_insert(uint32_t (&t)[kTopK], uint32_t key) {
    t[3] = key;
    uint32_t a = min(t[2], t[3]), b = max(t[2], t[3]);
    t[2] = a; t[3] = b;
    a = min(t[1], t[2]); b = max(t[1], t[2]);
    t[1] = a; t[2] = b;
    a = min(t[0], t[1]); b = max(t[0], t[1]);
    t[0] = a; t[1] = b;
}

__device__ __forceinline__ void topk_insert2(uint32_t (&t)[kTopK], int32_t (&v)[kTopK], uint32_t key, int32_t val) {
    t[3] = key; v[3] = val;
#pragma unroll
    for (int k = 3; k > 0; --k) {
        if (t[k] < t[k - 1]) {
            const uint32_t a = t[k]; t[k] = t[k - 1]; t[k - 1] = a;
            const int32_t c = v[k]; v[k] = v[k - 1]; v[k - 1] = c;
        }
    }
}

// The K smallest keys held anywhere in the warp (each lane's list sorted, keys unique except kEmptyKey).
__device__ __forceinline__ void warp_topk_merge(uint32_t (&t)[kTopK], uint32_t (&out)[kTopK]) {
#pragma unroll
    for (int k = 0; k < kTopK; ++k) {
        const uint32_t m = __reduce_min_sync(0xffffffffu, t[0]);
        out[k] = m;
        if (t[0] == m && m != kEmptyKey) { t[0] = t[1]; t[1] = t[2]; t[2] = t[3]; t[3] = kEmptyKey; }
    }
}

__device__ __forceinline__ void warp_topk_merge2(uint32_t (&t)[kTopK], int32_t (&v)[kTopK], uint32_t (&out)[kTopK],
                                                 int32_t (&outv)[kTopK]) {
#pragma unroll
    for (int k = 0; k < kTopK; ++k) {
        const uint32_t m = __reduce_min_sync(0xffffffffu, t[0]);
        const unsigned own = __ballot_sync(0xffffffffu, t[0] == m);
        const int src = __ffs(own) - 1;
        const int32_t val = __shfl_sync(0xffffffffu, v[0], src);
        out[k] = m;
        outv[k] = m == kEmptyKey ? -1 : val;
        if (t[0] == m && m != kEmptyKey) {
            t[0] = t[1]; t[1] = t[2]; t[2] = t[3]; t[3] = kEmptyKey;
            v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = -1;
        }
    }
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int lo, int hi, int key) {  // first i in [lo,hi) with a[i] >= key
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// last i in [lo, hi) with a[i] <= key (a ascending, a[lo] <= key)
__device__ __forceinline__ int upper_slot_i32(const int32_t* a, int lo, int hi, int key) {
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_slot_i64(const long long* a, int lo, int hi, long long key) {
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

// the rotation-histogram bin shared by the BoW and triangulation searches (e.g. ORBmatcher.cc:718-728):
// rot = a1 - a2; if (rot < 0.0) rot += 360.0f; bin = round(rot * (1.0f / HISTO_LENGTH)); 30 -> 0
__device__ __forceinline__ int rot_bin(float a1, float a2) {
    const float factor = 1.0f / kHisto;
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == kHisto) bin = 0;
    return bin;
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1854-1895) on bin sizes.
__device__ __forceinline__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

// One pair of frames of a node-scan search, resolved from the per-pair control arrays.
struct PairCtx {
    int fa, fb;          // frame indices in set 1 / set 2
    int ka, kb;          // first keypoint of the frames
    int na, nb;          // keypoints per frame
    int a0, a1;          // node range of frame fa
    int b0, b1;          // node range of frame fb
    int fbase;           // first fv_feat entry of frame fa
};

__device__ __forceinline__ PairCtx pair_ctx(const FrameSetView& S1, const FrameSetView& S2, const int32_t* idx1,
                                            const int32_t* idx2, int p) {
    PairCtx c;
    c.fa = idx1[p]; c.fb = idx2[p];
    c.ka = S1.kp_off[c.fa]; c.na = S1.kp_off[c.fa + 1] - c.ka;
    c.kb = S2.kp_off[c.fb]; c.nb = S2.kp_off[c.fb + 1] - c.kb;
    c.a0 = S1.node_off[c.fa]; c.a1 = S1.node_off[c.fa + 1];
    c.b0 = S2.node_off[c.fb]; c.b1 = S2.node_off[c.fb + 1];
    c.fbase = S1.feat_off[c.a0];
    return c;
}

// Scan entry e of a pair = the e-th fv_feat entry of frame fa (nodes ascending, entries in vector order): this IS
// the reference's query order (the merge walk visits equal node ids in ascending order, :664-750).  Resolves the
// owning node a and the node b of frame fb with the same id (-1 when frame fb has no such node).
__device__ __forceinline__ void entry_nodes(const FrameSetView& S1, const FrameSetView& S2, const PairCtx& c, int e, int& a,
                                            int& b) {
    a = upper_slot_i32(S1.feat_off, c.a0, c.a1, c.fbase + e);
    const int id = S1.node_id[a];
    const int j = lower_bound_i32(S2.node_id, c.b0, c.b1, id);
    b = (j < c.b1 && S2.node_id[j] == id) ? j : -1;
}

}  // namespace og

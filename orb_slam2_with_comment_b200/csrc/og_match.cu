// og_match.cu — CUDA kernels (sm_100a) and C ABI of the Hamming path of ORBmatcher (ORBmatcher.cc).
//
//   k_hamming_pairs     DescriptorDistance for n pairs                                     (:1901-1917)
//   k_bow_plan          finds the node pairs large enough for the register-tiled scan
//   k_bow_topk_tile     phase A of the BoW-node scans, register-tiled: a thread owns RQ queries, candidates are
//                       staged in shared memory, 8 POPC per 256-bit pair                   (:664-709, :240-282)
//   k_bow_topk_warp     phase A, one warp per query (small nodes)
//   k_bow_select        phase B: greedy walk in the reference's query order, ratio test, rotation histogram
//                                                                                          (:711-765, :284-341)
//   k_tri_scan          SearchForTriangulation, one warp per query (no greedy state)       (:840-926)
//   k_tri_finalize      vMatches12 + rotation histogram                                    (:908-972)
//   k_grid_build        Frame::AssignFeaturesToGrid                                        (Frame.cc:232-247)
//   k_sbp_topk          SearchByProjection phase A: window -> candidates -> top-K          (:66-139, Frame.cc:353-410)
//   k_sbp_select        phase B: greedy walk over the map points in vector order           (:108-110, :142-151)
//
// See og_match.cuh for the two-phase scheme.  Integer/popcount work: no tensor cores.
#include "og_nvtx.h"
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/orbgpu.h"
#include "og_match.cuh"

int og_fail(int code, const std::string& msg);  // og_capi.cu (thread-local last error)

namespace og {

// ------------------------------------------------------------------------------------------------------------
__global__ void k_hamming_pairs(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = hamming256(load_desc(a, i), load_desc(b, i));
}

// ------------------------------------------------------------------------------------------------------------
// BoW-node scans
// ------------------------------------------------------------------------------------------------------------
struct BowArgs {
    FrameSetView S1, S2;
    int n_pairs;
    const int32_t *idx1, *idx2;
    const long long* entry_off;   // [n_pairs+1] scan entries (queries) before pair p
    const long long* match_off;   // [n_pairs]
    int require_mp2;
    int dmax;                     // distances above it cannot change any accept decision (see bow_dev): phase A drops them
    int dense_q, dense_c;         // a node pair with n1 >= dense_q && n2 >= dense_c goes to the tiled kernel (dense_q<=0: never)
    uint32_t* topk;               // [total entries][kTopK]
    int4* items;                  // tiled-kernel work list
    int* n_items;
    int* n_sparse;                // node pairs left to the warp kernel (null: unknown, scan everything)
    int items_cap;
    unsigned long long* evals;
};

__device__ __forceinline__ bool is_dense(const BowArgs& A, int n1, int n2) {
    return A.dense_q > 0 && n1 >= A.dense_q && n2 >= A.dense_c;
}

constexpr int kTileThreads = 256;
constexpr int kTileChunk = 256;   // candidates staged per round (8 KB)

template <int RQ>
__global__ void k_bow_plan(const __grid_constant__ BowArgs A) {
    const int p = blockIdx.x;
    const PairCtx c = pair_ctx(A.S1, A.S2, A.idx1, A.idx2, p);
    const int a = c.a0 + blockIdx.y * blockDim.x + threadIdx.x;
    if (a >= c.a1) return;
    const int id = A.S1.node_id[a];
    const int j = lower_bound_i32(A.S2.node_id, c.b0, c.b1, id);
    const int n1 = A.S1.feat_off[a + 1] - A.S1.feat_off[a];
    const bool matched = j < c.b1 && A.S2.node_id[j] == id;
    const int n2 = matched ? A.S2.feat_off[j + 1] - A.S2.feat_off[j] : 0;
    if (!matched || !is_dense(A, n1, n2)) {
        // the warp kernel owns this node's queries (it also writes the empty lists of nodes frame 2 does not have)
        if (n1 > 0) atomicAdd(A.n_sparse, 1);
        return;
    }
    const int QT = kTileThreads * RQ;
    const int nt = (n1 + QT - 1) / QT;
    const int base = atomicAdd(A.n_items, nt);
    for (int t = 0; t < nt && base + t < A.items_cap; ++t) A.items[base + t] = make_int4(p, a, j, t);
}

template <int RQ>
__global__ void __launch_bounds__(kTileThreads) k_bow_topk_tile(const __grid_constant__ BowArgs A) {
    __shared__ uint4 cd[kTileChunk * 2];
    __shared__ int cpos[kTileChunk];
    __shared__ int s_nv, s_ne;
    const int tid = threadIdx.x;
    const int dmax = A.dmax;
    const int n_items = min(*A.n_items, A.items_cap);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int4 it = A.items[w];
        const int p = it.x, a = it.y, b = it.z, tile = it.w;
        const PairCtx c = pair_ctx(A.S1, A.S2, A.idx1, A.idx2, p);
        const int a_off = A.S1.feat_off[a], n1 = A.S1.feat_off[a + 1] - a_off;
        const int b_off = A.S2.feat_off[b], n2 = A.S2.feat_off[b + 1] - b_off;

        Desc q[RQ];
        bool valid[RQ];
        uint32_t t[RQ][kTopK];
        int nvalid = 0;
#pragma unroll
        for (int r = 0; r < RQ; ++r) {
            const int qi = (tile * RQ + r) * kTileThreads + tid;
            valid[r] = false;
            q[r].lo = make_uint4(0, 0, 0, 0);
            q[r].hi = make_uint4(0, 0, 0, 0);
            if (qi < n1) {
                const int i1 = A.S1.feat[a_off + qi];
                // queries need a valid MapPoint (:673-677, :233-238)
                valid[r] = A.S1.flags && (A.S1.flags[c.ka + i1] & 1);
                if (valid[r]) q[r] = load_desc(A.S1.desc, c.ka + i1);
            }
            nvalid += valid[r];
#pragma unroll
            for (int k = 0; k < kTopK; ++k) t[r][k] = kEmptyKey;
        }
        int nelig = 0;
        for (int c0 = 0; c0 < n2; c0 += kTileChunk) {
            __syncthreads();
            {
                const int j = c0 + tid;
                int pos = -1;
                if (j < n2) {
                    const int i2 = A.S2.feat[b_off + j];
                    const bool ok = !A.require_mp2 || (A.S2.flags && (A.S2.flags[c.kb + i2] & 1));
                    const Desc d = load_desc(A.S2.desc, c.kb + i2);
                    cd[2 * tid] = d.lo;
                    cd[2 * tid + 1] = d.hi;
                    pos = ok ? j : -1;
                }
                cpos[tid] = pos;
                nelig += pos >= 0;
            }
            __syncthreads();
            const int m = min(kTileChunk, n2 - c0);
#pragma unroll 2
            for (int jj = 0; jj < m; ++jj) {
                const int pos = cpos[jj];
                if (pos < 0) continue;   // warp-uniform
                Desc d;
                d.lo = cd[2 * jj];
                d.hi = cd[2 * jj + 1];
#pragma unroll
                for (int r = 0; r < RQ; ++r) {
                    const int dist = hamming256_csa(q[r], d);
                    if (dist <= dmax) {
                        const uint32_t key = ((uint32_t)dist << kPosBits) | (uint32_t)pos;
                        if (key < t[r][kTopK - 1]) topk_insert(t[r], key);
                    }
                }
            }
        }
        const long long g0 = A.entry_off[p] + (a_off - c.fbase);
#pragma unroll
        for (int r = 0; r < RQ; ++r) {
            const int qi = (tile * RQ + r) * kTileThreads + tid;
            if (qi < n1) {
                uint4 o = valid[r] ? make_uint4(t[r][0], t[r][1], t[r][2], t[r][3]) : make_uint4(kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey);
                *reinterpret_cast<uint4*>(A.topk + (g0 + qi) * kTopK) = o;
            }
        }
        // distance evaluations of this work item (also the barrier before the next item's staging)
        __syncthreads();
        if (tid == 0) { s_nv = 0; s_ne = 0; }
        __syncthreads();
        if (nvalid) atomicAdd(&s_nv, nvalid);
        if (nelig) atomicAdd(&s_ne, nelig);
        __syncthreads();
        if (tid == 0) atomicAdd(A.evals, (unsigned long long)s_nv * (unsigned long long)s_ne);
    }
}

// One warp scans the candidates of node b for one query; `taken` (bitmask over frame fb's keypoints, may be null)
// removes candidates already matched.  Returns the K smallest keys, identical in every lane.
__device__ __forceinline__ int bow_scan_warp(const FrameSetView& S2, int kb, int b, const Desc& dq, int require_mp2,
                                             const uint32_t* taken, int dmax, uint32_t (&out)[kTopK]) {
    const int lane = threadIdx.x & 31;
    const int b_off = S2.feat_off[b], n2 = S2.feat_off[b + 1] - b_off;
    uint32_t t[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; ++k) t[k] = kEmptyKey;
    int evals = 0;
    for (int j = lane; j < n2; j += 32) {
        const int i2 = S2.feat[b_off + j];
        if (require_mp2 && !(S2.flags && (S2.flags[kb + i2] & 1))) continue;
        if (taken && ((taken[i2 >> 5] >> (i2 & 31)) & 1u)) continue;
        const int dist = hamming256(dq, load_desc(S2.desc, kb + i2));
        ++evals;
        if (dist > dmax) continue;
        const uint32_t key = ((uint32_t)dist << kPosBits) | (uint32_t)j;
        if (key < t[kTopK - 1]) topk_insert(t, key);
    }
    warp_topk_merge(t, out);
    return __reduce_add_sync(0xffffffffu, evals);   // warp total, identical in every lane
}

constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_bow_topk_warp(const __grid_constant__ BowArgs A, long long total_entries) {
    __shared__ int s_evals;
    if (A.n_sparse && *A.n_sparse == 0) return;   // every matched node pair went to the tiled kernel (which also wrote the empty lists)
    if (threadIdx.x == 0) s_evals = 0;
    __syncthreads();
    int evals = 0;
    // grid-stride over the queries: the grid is a few waves of CTAs, not one CTA per 8 queries (a call with a million CTAs
    // that all exit at once still cost 0.5 ms of launch work)
    for (long long g = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); g < total_entries; g += (long long)gridDim.x * kWarpsPerBlock) {
        const int p = upper_slot_i64(A.entry_off, 0, A.n_pairs, g);
        const PairCtx c = pair_ctx(A.S1, A.S2, A.idx1, A.idx2, p);
        const int e = (int)(g - A.entry_off[p]);
        int a, b;
        entry_nodes(A.S1, A.S2, c, e, a, b);
        bool dense = false;
        if (b >= 0) dense = is_dense(A, A.S1.feat_off[a + 1] - A.S1.feat_off[a], A.S2.feat_off[b + 1] - A.S2.feat_off[b]);
        if (!dense) {
            uint32_t out[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
            const int i1 = A.S1.feat[c.fbase + e];
            const bool ok = b >= 0 && A.S1.flags && (A.S1.flags[c.ka + i1] & 1);
            if (ok) evals += bow_scan_warp(A.S2, c.kb, b, load_desc(A.S1.desc, c.ka + i1), A.require_mp2, nullptr, A.dmax, out);
            if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint4*>(A.topk + g * kTopK) = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
    if (evals && (threadIdx.x & 31) == 0) atomicAdd(&s_evals, evals);
    __syncthreads();
    if (threadIdx.x == 0 && s_evals) atomicAdd(A.evals, (unsigned long long)s_evals);
}

struct SelectArgs {
    float nnratio;
    int check_orientation, th_low, th_inclusive;
    int8_t* entry_bin;     // [total entries]
    int32_t *match12, *match_dist, *nmatches;
    int taken_words;       // shared-memory words of the taken mask
};

// Phase B, one warp per pair, kSelectWarps pairs per CTA (the walk is sequential by definition, so throughput comes from
// resident warps: a 32-thread CTA capped the SM at 32 pairs, 3 % occupancy in the r1 profile).
constexpr int kSelectWarps = 4;
__global__ void __launch_bounds__(kSelectWarps * 32) k_bow_select(const __grid_constant__ BowArgs A, const __grid_constant__ SelectArgs Z) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    uint32_t* taken = smem + (size_t)wi * (Z.taken_words + 32);
    int* hist = (int*)(taken + Z.taken_words);
    const int p = blockIdx.x * kSelectWarps + wi;
    if (p >= A.n_pairs) return;
    const PairCtx c = pair_ctx(A.S1, A.S2, A.idx1, A.idx2, p);
    for (int i = lane; i < (c.nb + 31) / 32; i += 32) taken[i] = 0;
    hist[lane] = 0;
    const long long mo = A.match_off[p];
    for (int i = lane; i < c.na; i += 32) {
        Z.match12[mo + i] = -1;
        if (Z.match_dist) Z.match_dist[mo + i] = -1;
    }
    __syncwarp();
    const long long g0 = A.entry_off[p];
    const int E = (int)(A.entry_off[p + 1] - g0);
    const int th = Z.th_low;
    auto pass_th = [&](int d) { return Z.th_inclusive ? d <= th : d < th; };
    int nacc = 0, evals = 0;
    for (int base = 0; base < E; base += 32) {
        // lane-parallel prefetch of this chunk's lists
        const int e = base + lane;
        uint32_t k[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
        int c2[kTopK] = {-1, -1, -1, -1};
        float ang2[kTopK] = {0.f, 0.f, 0.f, 0.f}, ang1 = 0.f;   // angles for the rotation histogram, fetched lane-parallel
        int i1 = -1, nb_node = -1;
        if (e < E) {
            const uint4 v = *reinterpret_cast<const uint4*>(A.topk + (g0 + e) * kTopK);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
            Z.entry_bin[g0 + e] = -1;
            if (k[0] != kEmptyKey && pass_th((int)(k[0] >> kPosBits))) {
                int a;
                entry_nodes(A.S1, A.S2, c, e, a, nb_node);
                i1 = A.S1.feat[c.fbase + e];
                const int b_off = A.S2.feat_off[nb_node];
#pragma unroll
                for (int j = 0; j < kTopK; ++j)
                    if (k[j] != kEmptyKey) c2[j] = A.S2.feat[b_off + (int)(k[j] & kPosMask)];
                if (Z.check_orientation) {
                    ang1 = A.S1.keys[c.ka + i1].angle;
#pragma unroll
                    for (int j = 0; j < kTopK; ++j)
                        if (c2[j] >= 0) ang2[j] = A.S2.keys[c.kb + c2[j]].angle;
                }
            }
        }
        // Speculative walk (see k_sbp_select): every pending lane takes best / second among its list entries not taken so far; a
        // lane must wait if an accepted lane before it takes its best or second candidate, or if its truncated list ran dry (the
        // rescan needs everything before it committed); the conflict-free prefix commits at once.
        unsigned pend = __ballot_sync(0xffffffffu, i1 >= 0);
        while (pend) {
            const bool mine = (pend >> lane) & 1u;
            int cnt = 0, best_c = -1, second_c = -1, d1 = 256, d2 = 256;
            float best_a = 0.f;
            bool complete = false;
            if (mine) {
#pragma unroll
                for (int j = 0; j < kTopK; ++j) {
                    if (cnt == 2 || complete) break;
                    if (k[j] == kEmptyKey) { complete = true; break; }
                    const int i2 = c2[j];
                    if ((taken[i2 >> 5] >> (i2 & 31)) & 1u) continue;
                    if (cnt == 0) { best_c = i2; best_a = ang2[j]; d1 = (int)(k[j] >> kPosBits); cnt = 1; }
                    else { second_c = i2; d2 = (int)(k[j] >> kPosBits); cnt = 2; }
                }
                if (cnt == 2) complete = true;
            }
            const bool rescan = mine && !complete && (cnt == 0 || pass_th(d1));
            const int first = __ffs(pend) - 1;
            if (__shfl_sync(0xffffffffu, (int)rescan, first)) {
                // the first pending lane's truncated list ran out: exact rescan of its node with the taken mask, by the whole warp
                const int qi1 = __shfl_sync(0xffffffffu, i1, first), qb = __shfl_sync(0xffffffffu, nb_node, first);
                const float qa1 = __shfl_sync(0xffffffffu, ang1, first);
                uint32_t out[kTopK];
                evals += bow_scan_warp(A.S2, c.kb, qb, load_desc(A.S1.desc, c.ka + qi1), A.require_mp2, taken, 256, out);
                int bc = -1, e1 = 256, e2 = 256;
                if (out[0] != kEmptyKey) {
                    bc = A.S2.feat[A.S2.feat_off[qb] + (int)(out[0] & kPosMask)];
                    e1 = (int)(out[0] >> kPosBits);
                }
                if (out[1] != kEmptyKey) e2 = (int)(out[1] >> kPosBits);
                if (bc >= 0 && pass_th(e1) && (float)e1 < __fmul_rn(Z.nnratio, (float)e2)) {
                    int bin = 0;
                    if (Z.check_orientation) bin = rot_bin(qa1, A.S2.keys[c.kb + bc].angle);
                    if (lane == 0) {
                        taken[bc >> 5] |= 1u << (bc & 31);
                        Z.match12[mo + qi1] = bc;
                        if (Z.match_dist) Z.match_dist[mo + qi1] = e1;
                        Z.entry_bin[g0 + base + first] = (int8_t)bin;
                        hist[bin] += 1;
                    }
                    ++nacc;
                }
                pend &= ~(1u << first);
                __syncwarp();
                continue;
            }
            const bool accept = mine && best_c >= 0 && pass_th(d1) && (float)d1 < __fmul_rn(Z.nnratio, (float)d2);
            bool stop = rescan;
            unsigned cl = __ballot_sync(0xffffffffu, accept);
            while (cl) {
                const int i = __ffs(cl) - 1;
                cl &= cl - 1;
                const int cb = __shfl_sync(0xffffffffu, best_c, i);
                if (mine && lane > i && (cb == best_c || cb == second_c)) stop = true;
            }
            const unsigned sb = __ballot_sync(0xffffffffu, stop);
            const unsigned done = sb ? (pend & ((1u << (__ffs(sb) - 1)) - 1u)) : pend;
            if (((done >> lane) & 1u) && accept) {
                int bin = 0;
                if (Z.check_orientation) bin = rot_bin(ang1, best_a);
                atomicOr(&taken[best_c >> 5], 1u << (best_c & 31));
                Z.match12[mo + i1] = best_c;
                if (Z.match_dist) Z.match_dist[mo + i1] = d1;
                Z.entry_bin[g0 + e] = (int8_t)bin;
                atomicAdd(&hist[bin], 1);
            }
            nacc += __popc(__ballot_sync(0xffffffffu, accept) & done);
            pend &= ~done;
            __syncwarp();
        }
    }
    __syncwarp();
    int removed = 0;
    if (Z.check_orientation) {
        int ind1, ind2, ind3;
        three_maxima(hist, kHisto, ind1, ind2, ind3);
        for (int e = lane; e < E; e += 32) {
            const int bin = Z.entry_bin[g0 + e];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) {
                const int i1 = A.S1.feat[c.fbase + e];
                Z.match12[mo + i1] = -1;
                if (Z.match_dist) Z.match_dist[mo + i1] = -1;
                ++removed;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
    }
    if (lane == 0) {
        if (Z.nmatches) Z.nmatches[p] = nacc - removed;
        if (evals) atomicAdd(A.evals, (unsigned long long)evals);
    }
}

// ------------------------------------------------------------------------------------------------------------
// SearchForTriangulation
// ------------------------------------------------------------------------------------------------------------
struct TriArgs {
    FrameSetView S1, S2;
    int n_pairs;
    const int32_t *idx1, *idx2;
    const long long* entry_off;
    const long long* match_off;
    const float* f12;        // [n_pairs][9]
    const float* epipole;    // [n_pairs][2]
    const float* scale;      // mvScaleFactors
    const float* sigma2;     // mvLevelSigma2
    int only_stereo, check_orientation;
    int32_t* res_idx2;       // [total entries]
    int32_t* res_dist;
    int32_t *match12, *match_dist, *nmatches;
    unsigned long long* evals;
};

// ORBmatcher::CheckDistEpipolarLine (ORBmatcher.cc:173-196), F12 row major, plain IEEE mul/add (no FMA)
__device__ __forceinline__ bool check_epipolar(float x1, float y1, float x2, float y2, const float* F, float sig2) {
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[0]), __fmul_rn(y1, F[3])), F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[1]), __fmul_rn(y1, F[4])), F[7]);
    const float cc = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[2]), __fmul_rn(y1, F[5])), F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, x2), __fmul_rn(b, y2)), cc);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0.f) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < 3.84 * (double)sig2;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_tri_scan(const __grid_constant__ TriArgs A, long long total_entries) {
    __shared__ int s_evals;
    if (threadIdx.x == 0) s_evals = 0;
    __syncthreads();
    const long long g = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    int evals = 0;
    if (g < total_entries) {
        const int p = upper_slot_i64(A.entry_off, 0, A.n_pairs, g);
        const PairCtx c = pair_ctx(A.S1, A.S2, A.idx1, A.idx2, p);
        const int e = (int)(g - A.entry_off[p]);
        int a, b;
        entry_nodes(A.S1, A.S2, c, e, a, b);
        const int i1 = A.S1.feat[c.fbase + e];
        bool ok = b >= 0 && !(A.S1.flags && (A.S1.flags[c.ka + i1] & 1));   // pMP1 set -> skip (:846)
        const bool stereo1 = A.S1.u_right && A.S1.u_right[c.ka + i1] >= 0.f;
        if (A.only_stereo && !stereo1) ok = false;
        uint32_t best = kEmptyKey;
        if (ok) {
            const Desc dq = load_desc(A.S1.desc, c.ka + i1);
            const float x1 = A.S1.keys[c.ka + i1].x, y1 = A.S1.keys[c.ka + i1].y;
            const float* F = A.f12 + 9 * (long long)p;
            const float ex = A.epipole[2 * p], ey = A.epipole[2 * p + 1];
            const int b_off = A.S2.feat_off[b], n2 = A.S2.feat_off[b + 1] - b_off;
            for (int j = lane; j < n2; j += 32) {
                const int i2 = A.S2.feat[b_off + j];
                if (A.S2.flags && (A.S2.flags[c.kb + i2] & 1)) continue;     // pMP2 set (:868)
                const bool stereo2 = A.S2.u_right && A.S2.u_right[c.kb + i2] >= 0.f;
                if (A.only_stereo && !stereo2) continue;
                const int dist = hamming256(dq, load_desc(A.S2.desc, c.kb + i2));
                ++evals;
                if (dist > ORBGPU_TH_LOW) continue;                            // (:882)
                const KeyPoint k2 = A.S2.keys[c.kb + i2];
                if (!stereo1 && !stereo2) {
                    const float dx = __fsub_rn(ex, k2.x), dy = __fsub_rn(ey, k2.y);
                    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, A.scale[k2.octave])) continue;
                }
                if (!check_epipolar(x1, y1, k2.x, k2.y, F, A.sigma2[k2.octave])) continue;
                // equal distances replace (:882 `dist>bestDist` -> continue), so the LAST minimum wins
                const uint32_t key = ((uint32_t)dist << kPosBits) | (kPosMask - (uint32_t)j);
                best = min(best, key);
            }
            best = __reduce_min_sync(0xffffffffu, best);
        }
        if (lane == 0) {
            int r2 = -1, rd = -1;
            if (best != kEmptyKey) {
                r2 = A.S2.feat[A.S2.feat_off[b] + (int)(kPosMask - (best & kPosMask))];
                rd = (int)(best >> kPosBits);
            }
            A.res_idx2[g] = r2;
            A.res_dist[g] = rd;
        }
    }
    if (evals) atomicAdd(&s_evals, evals);
    __syncthreads();
    if (threadIdx.x == 0 && s_evals) atomicAdd(A.evals, (unsigned long long)s_evals);
}

__global__ void __launch_bounds__(32) k_tri_finalize(const __grid_constant__ TriArgs A) {
    __shared__ int hist[32];
    const int lane = threadIdx.x, p = blockIdx.x;
    const PairCtx c = pair_ctx(A.S1, A.S2, A.idx1, A.idx2, p);
    hist[lane] = 0;
    const long long mo = A.match_off[p];
    for (int i = lane; i < c.na; i += 32) {
        A.match12[mo + i] = -1;
        if (A.match_dist) A.match_dist[mo + i] = -1;
    }
    __syncwarp();
    const long long g0 = A.entry_off[p];
    const int E = (int)(A.entry_off[p + 1] - g0);
    int n = 0;
    for (int e = lane; e < E; e += 32) {
        const int r2 = A.res_idx2[g0 + e];
        if (r2 >= 0) {
            const int i1 = A.S1.feat[c.fbase + e];
            A.match12[mo + i1] = r2;
            if (A.match_dist) A.match_dist[mo + i1] = A.res_dist[g0 + e];
            ++n;
            if (A.check_orientation) atomicAdd(&hist[rot_bin(A.S1.keys[c.ka + i1].angle, A.S2.keys[c.kb + r2].angle)], 1);
        }
    }
    __syncwarp();
    if (A.check_orientation) {
        int ind1, ind2, ind3;
        three_maxima(hist, kHisto, ind1, ind2, ind3);
        for (int e = lane; e < E; e += 32) {
            const int r2 = A.res_idx2[g0 + e];
            if (r2 >= 0) {
                const int i1 = A.S1.feat[c.fbase + e];
                const int bin = rot_bin(A.S1.keys[c.ka + i1].angle, A.S2.keys[c.kb + r2].angle);
                if (bin != ind1 && bin != ind2 && bin != ind3) {
                    A.match12[mo + i1] = -1;
                    if (A.match_dist) A.match_dist[mo + i1] = -1;
                    --n;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    if (lane == 0 && A.nmatches) A.nmatches[p] = n;
}

// ------------------------------------------------------------------------------------------------------------
// SearchByProjection(Frame&, vector<MapPoint*>&, th)
// ------------------------------------------------------------------------------------------------------------
constexpr int kGridCells = kGridCols * kGridRows;
constexpr int kGridThreads = 256;

// Queries of the generic windowed search (orbgpu_window_query_set): explicit window and level range per query.
struct WindowQueryView {
    const int32_t* q_off;
    const float *u, *v, *radius, *ur, *angle;
    const int32_t *min_level, *max_level;
    const uint8_t* flags;
    const uint8_t* desc;
};

struct SbpArgs {
    FrameSetView F;
    MapPointView M;
    WindowQueryView W;
    int generic;           // 0: SearchByProjection(Frame&, vector<MapPoint*>&, th) queries from M; 1: queries from W
    int skip_any;          // generic: candidates holding ANY MapPoint are skipped (:1776), else only those with observations (:1619-1621)
    int th_dist;           // generic: accept when best <= th_dist
    int check_orientation; // generic: rotation histogram
    int8_t* q_bin;         // generic: [total queries] histogram bin of the accepted query, -1 otherwise
    const float* inv_sigma2;  // generic, best-only searches: mvInvLevelSigma2 for the chi-square gate of Fuse (:1077-1102); null = no gate
    int no_xr_window;      // generic: the |ur - mvuRight| <= radius test (:1624-1630) is off (Fuse / SearchBySim3 have none)
    int n_frames;
    const float* scale;
    float th, nnratio;
    int32_t* cell_start;   // [n_frames][kGridCells + 1], offsets into the frame's slice of cell_items
    int32_t* cell_items;   // [total keypoints] keypoint indices grouped by cell, ascending inside a cell
    uint32_t* topk_key;    // [total map points][kTopK]
    int32_t* topk_idx;
    int32_t *kp_match, *mp_best_idx, *mp_best_dist, *mp_second_dist, *nmatches;
    int blocked_words;
    unsigned long long* evals;
};

// Frame::AssignFeaturesToGrid + PosInGrid (Frame.cc:232-247, :412-422): one CTA per frame.
__global__ void __launch_bounds__(kGridThreads) k_grid_build(const __grid_constant__ SbpArgs A) {
    __shared__ int cnt[kGridCells];
    __shared__ int wsum[kGridThreads / 32];
    const int f = blockIdx.x, t = threadIdx.x;
    const int k0 = A.F.kp_off[f], n = A.F.kp_off[f + 1] - k0;
    const float* g = A.F.grid + 4 * f;
    const float minx = g[0], miny = g[1], iw = g[2], ih = g[3];
    for (int i = t; i < kGridCells; i += kGridThreads) cnt[i] = 0;
    __syncthreads();
    auto cell_of = [&](int i) {
        const KeyPoint kp = A.F.keys[k0 + i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, minx), iw));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, miny), ih));
        if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) return -1;
        return px * kGridRows + py;
    };
    for (int i = t; i < n; i += kGridThreads) {
        const int c = cell_of(i);
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan of the 3072 counts: 12 consecutive cells per thread
    constexpr int per = kGridCells / kGridThreads;
    int loc[per], s = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) { loc[k] = cnt[t * per + k]; s += loc[k]; }
    const int lane = t & 31, w = t >> 5;
    int v = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += wsum[k];
    int run = woff + v - s;
    int32_t* cs = A.cell_start + (long long)f * (kGridCells + 1);
#pragma unroll
    for (int k = 0; k < per; ++k) {
        cs[t * per + k] = run;
        cnt[t * per + k] = run;   // becomes the fill cursor
        run += loc[k];
    }
    if (t == kGridThreads - 1) cs[kGridCells] = run;
    __syncthreads();
    int32_t* items = A.cell_items + k0;
    for (int i = t; i < n; i += kGridThreads) {
        const int c = cell_of(i);
        if (c >= 0) items[atomicAdd(&cnt[c], 1)] = i;
    }
    __syncthreads();
    // push_back order = ascending keypoint index: sort every cell's short run
    for (int c = t; c < kGridCells; c += kGridThreads) {
        const int b = cs[c], e = cnt[c];
        for (int i = b + 1; i < e; ++i) {
            const int x = items[i];
            int j = i - 1;
            while (j >= b && items[j] > x) { items[j + 1] = items[j]; --j; }
            items[j + 1] = x;
        }
    }
}

struct SbpQuery {
    bool live, use_xr;
    int f, k0, lvl, min_level, max_level;
    float x, y, rs, xr;
    int minx, maxx, miny, maxy;
};

// the per-map-point part of :66-87 and the window of Frame::GetFeaturesInArea (Frame.cc:358-377)
__device__ __forceinline__ SbpQuery sbp_query(const SbpArgs& A, int q, int f) {
    SbpQuery Q;
    Q.live = false;
    Q.f = f;
    Q.k0 = A.F.kp_off[f];
    if (A.generic) {
        if (!(A.W.flags[q] & 1)) return Q;
        Q.lvl = 0;
        Q.min_level = A.W.min_level[q];
        Q.max_level = A.W.max_level[q];
        Q.rs = A.W.radius[q];
        Q.x = A.W.u[q];
        Q.y = A.W.v[q];
        Q.use_xr = A.W.ur != nullptr;
        Q.xr = Q.use_xr ? A.W.ur[q] : 0.f;
    } else {
        const int fl = A.M.flags[q];
        if (!(fl & 1) || (fl & 2)) return Q;   // !mbTrackInView || isBad()
        Q.lvl = A.M.level[q];
        Q.min_level = Q.lvl - 1;
        Q.max_level = Q.lvl;
        float r = ((double)A.M.view_cos[q] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos (:157-163)
        if (A.th != 1.0f) r = __fmul_rn(r, A.th);
        Q.rs = __fmul_rn(r, A.scale[Q.lvl]);
        Q.x = A.M.proj_x[q];
        Q.y = A.M.proj_y[q];
        Q.use_xr = true;
        Q.xr = A.M.proj_xr ? A.M.proj_xr[q] : 0.f;
    }
    const float* g = A.F.grid + 4 * f;
    Q.minx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.x, g[0]), Q.rs), g[2])));
    if (Q.minx >= kGridCols) return Q;
    Q.maxx = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.x, g[0]), Q.rs), g[2])));
    if (Q.maxx < 0) return Q;
    Q.miny = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.y, g[1]), Q.rs), g[3])));
    if (Q.miny >= kGridRows) return Q;
    Q.maxy = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.y, g[1]), Q.rs), g[3])));
    if (Q.maxy < 0) return Q;
    Q.live = Q.maxx >= Q.minx && Q.maxy >= Q.miny;
    return Q;
}

// One warp enumerates the window's candidates in the reference's order (ix outer, iy inner, ascending index in a
// cell) and keeps the K smallest (dist, position).  `blocked` = bitmask of keypoints taken during phase B.
__device__ __forceinline__ int sbp_scan_warp(const SbpArgs& A, const SbpQuery& Q, const Desc& dq, const uint32_t* blocked,
                                             uint32_t (&out)[kTopK], int32_t (&outi)[kTopK], const uint16_t* mdist = nullptr) {
    const int lane = threadIdx.x & 31;
    uint32_t t[kTopK];
    int32_t v[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; ++k) { t[k] = kEmptyKey; v[k] = -1; }
    const int ny = Q.maxy - Q.miny + 1, ncell = (Q.maxx - Q.minx + 1) * ny;
    const int32_t* cs = A.cell_start + (long long)Q.f * (kGridCells + 1);
    const int32_t* items = A.cell_items + Q.k0;
    const int minLevel = Q.min_level, maxLevel = Q.max_level;
    const bool check_levels = (minLevel > 0) || (maxLevel >= 0);
    int base = 0, evals = 0;
    for (int c0 = 0; c0 < ncell; c0 += 32) {
        const int ci = c0 + lane;
        int b = 0, n = 0;
        if (ci < ncell) {
            const int ix = Q.minx + ci / ny, iy = Q.miny + ci % ny;
            b = cs[ix * kGridRows + iy];
            n = cs[ix * kGridRows + iy + 1] - b;
        }
        if (!__any_sync(0xffffffffu, n > 0)) continue;   // these cells hold no key point
        int inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        const int pos0 = base + inc - n;
        base += __shfl_sync(0xffffffffu, inc, 31);
        for (int j = 0; j < n; ++j) {
            const int idx = items[b + j];
            const KeyPoint kp = A.F.keys[Q.k0 + idx];
            if (check_levels) {
                if (kp.octave < minLevel) continue;
                if (maxLevel >= 0 && kp.octave > maxLevel) continue;
            }
            if (!(fabsf(__fsub_rn(kp.x, Q.x)) < Q.rs && fabsf(__fsub_rn(kp.y, Q.y)) < Q.rs)) continue;
            if (A.F.flags) {   // holds a MapPoint with observations (:108-110, :1619-1621) / any MapPoint (:1776)
                const int st = A.F.flags[Q.k0 + idx];
                if (st == 1 || (A.skip_any && st != 0)) continue;
            }
            if (blocked && ((blocked[idx >> 5] >> (idx & 31)) & 1u)) continue;
            if (A.inv_sigma2) {   // reprojection error against the candidate's key point, chi-square at 95 % (ORBmatcher.cc:1077-1102)
                const float kur = A.F.u_right ? A.F.u_right[Q.k0 + idx] : -1.f;
                const float ex = __fsub_rn(Q.x, kp.x), ey = __fsub_rn(Q.y, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                double lim = 5.99;
                if (kur >= 0.f) {
                    const float er = __fsub_rn(Q.xr, kur);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    lim = 7.8;
                }
                if ((double)__fmul_rn(e2, A.inv_sigma2[kp.octave]) > lim) continue;
            }
            if (Q.use_xr && !A.no_xr_window && A.F.u_right && A.F.u_right[Q.k0 + idx] > 0.f) {   // (:113-118, :1624-1630)
                const float er = fabsf(__fsub_rn(Q.xr, A.F.u_right[Q.k0 + idx]));
                if (er > Q.rs) continue;
            }
            const uint32_t dist = (uint32_t)hamming256(dq, load_desc(A.F.desc, Q.k0 + idx));
            ++evals;
            if (mdist && (uint32_t)mdist[idx] <= dist) continue;   // SearchForInitialization: vMatchedDistance[i2] <= dist (:546)
            const uint32_t key = (dist << kPosBits) | (uint32_t)(pos0 + j);
            if (key < t[kTopK - 1]) topk_insert2(t, v, key, idx);
        }
    }
    if (!__any_sync(0xffffffffu, t[0] != kEmptyKey)) {   // an empty window (most local map points at th = 1): nothing to merge
#pragma unroll
        for (int k = 0; k < kTopK; ++k) { out[k] = kEmptyKey; outi[k] = -1; }
        return __reduce_add_sync(0xffffffffu, evals);
    }
    warp_topk_merge2(t, v, out, outi);
    return __reduce_add_sync(0xffffffffu, evals);   // warp total, identical in every lane
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_sbp_topk(const __grid_constant__ SbpArgs A, int total_mp) {
    __shared__ int s_evals;
    if (threadIdx.x == 0) s_evals = 0;
    __syncthreads();
    const int q = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int evals = 0;
    if (q < total_mp) {
        const int f = upper_slot_i32(A.generic ? A.W.q_off : A.M.mp_off, 0, A.n_frames, q);
        const SbpQuery Q = sbp_query(A, q, f);
        uint32_t out[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
        int32_t outi[kTopK] = {-1, -1, -1, -1};
        if (Q.live) evals = sbp_scan_warp(A, Q, load_desc(A.generic ? A.W.desc : A.M.desc, q), nullptr, out, outi);
        if ((threadIdx.x & 31) == 0) {
            *reinterpret_cast<uint4*>(A.topk_key + (long long)q * kTopK) = make_uint4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<int4*>(A.topk_idx + (long long)q * kTopK) = make_int4(outi[0], outi[1], outi[2], outi[3]);
        }
    }
    if (evals && (threadIdx.x & 31) == 0) atomicAdd(&s_evals, evals);
    __syncthreads();
    if (threadIdx.x == 0 && s_evals) atomicAdd(A.evals, (unsigned long long)s_evals);
}

// Phase A with G lanes per query instead of a warp: a window at th = 1 covers a handful of grid cells (one lane each) holding a
// key point or two, so a full warp per map point left three quarters of its lanes without a cell.  The group enumerates the
// candidates in the same order (ix outer, iy inner, ascending index in a cell) and merges its lanes' lists with group-wide
// min / ballot / shuffle (all *_sync calls carry the group's own mask: groups of one warp run different trip counts).
template <int G>
__device__ __forceinline__ int sbp_scan_group(const SbpArgs& A, const SbpQuery& Q, const Desc& dq, uint32_t (&out)[kTopK], int32_t (&outi)[kTopK]) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(G - 1));
    uint32_t t[kTopK];
    int32_t v[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; ++k) { t[k] = kEmptyKey; v[k] = -1; }
    const int ny = Q.maxy - Q.miny + 1, ncell = (Q.maxx - Q.minx + 1) * ny;
    const int32_t* cs = A.cell_start + (long long)Q.f * (kGridCells + 1);
    const int32_t* items = A.cell_items + Q.k0;
    const int minLevel = Q.min_level, maxLevel = Q.max_level;
    const bool check_levels = (minLevel > 0) || (maxLevel >= 0);
    int base = 0, evals = 0;
    for (int c0 = 0; c0 < ncell; c0 += G) {
        const int ci = c0 + gl;
        int b = 0, n = 0;
        if (ci < ncell) {
            const int ix = Q.minx + ci / ny, iy = Q.miny + ci % ny;
            b = cs[ix * kGridRows + iy];
            n = cs[ix * kGridRows + iy + 1] - b;
        }
        if (__ballot_sync(gmask, n > 0) == 0) continue;   // these cells hold no key point
        int inc = n;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const int u = __shfl_up_sync(gmask, inc, d, G);
            if (gl >= d) inc += u;
        }
        const int pos0 = base + inc - n;
        base += __shfl_sync(gmask, inc, G - 1, G);
        for (int j = 0; j < n; ++j) {
            const int idx = items[b + j];
            const KeyPoint kp = A.F.keys[Q.k0 + idx];
            if (check_levels) {
                if (kp.octave < minLevel) continue;
                if (maxLevel >= 0 && kp.octave > maxLevel) continue;
            }
            if (!(fabsf(__fsub_rn(kp.x, Q.x)) < Q.rs && fabsf(__fsub_rn(kp.y, Q.y)) < Q.rs)) continue;
            if (A.F.flags) {
                const int st = A.F.flags[Q.k0 + idx];
                if (st == 1 || (A.skip_any && st != 0)) continue;
            }
            if (A.inv_sigma2) {
                const float kur = A.F.u_right ? A.F.u_right[Q.k0 + idx] : -1.f;
                const float ex = __fsub_rn(Q.x, kp.x), ey = __fsub_rn(Q.y, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                double lim = 5.99;
                if (kur >= 0.f) {
                    const float er = __fsub_rn(Q.xr, kur);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    lim = 7.8;
                }
                if ((double)__fmul_rn(e2, A.inv_sigma2[kp.octave]) > lim) continue;
            }
            if (Q.use_xr && !A.no_xr_window && A.F.u_right && A.F.u_right[Q.k0 + idx] > 0.f) {
                const float er = fabsf(__fsub_rn(Q.xr, A.F.u_right[Q.k0 + idx]));
                if (er > Q.rs) continue;
            }
            const uint32_t dist = (uint32_t)hamming256(dq, load_desc(A.F.desc, Q.k0 + idx));
            ++evals;
            const uint32_t key = (dist << kPosBits) | (uint32_t)(pos0 + j);
            if (key < t[kTopK - 1]) topk_insert2(t, v, key, idx);
        }
    }
    if (__ballot_sync(gmask, t[0] != kEmptyKey) == 0) {
#pragma unroll
        for (int k = 0; k < kTopK; ++k) { out[k] = kEmptyKey; outi[k] = -1; }
        return __reduce_add_sync(gmask, evals);
    }
#pragma unroll
    for (int k = 0; k < kTopK; ++k) {
        const uint32_t m = __reduce_min_sync(gmask, t[0]);
        const unsigned own = __ballot_sync(gmask, t[0] == m);
        const int32_t val = __shfl_sync(gmask, v[0], __ffs(own) - 1);
        out[k] = m;
        outi[k] = m == kEmptyKey ? -1 : val;
        if (t[0] == m && m != kEmptyKey) {
            t[0] = t[1]; t[1] = t[2]; t[2] = t[3]; t[3] = kEmptyKey;
            v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = -1;
        }
    }
    return __reduce_add_sync(gmask, evals);   // group total, identical in every lane of the group
}

// G lanes per query.  Measured on 256 TUM frames x 5000 map points at th = 1 (whole search): warp 1.41 ms, G = 16 1.04, 8 0.84,
// 4 0.73, 2 0.68, 1 0.68: 4 for the narrow windows of th <= 2, 8 for wider ones and the generic windowed searches.
template <int kSbpGroup>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_sbp_topk_g(const __grid_constant__ SbpArgs A, int total_mp) {
    __shared__ int s_evals;
    if (threadIdx.x == 0) s_evals = 0;
    __syncthreads();
    const int q = (blockIdx.x * kWarpsPerBlock * 32 + threadIdx.x) / kSbpGroup;
    int evals = 0;
    if (q < total_mp) {
        const int f = upper_slot_i32(A.generic ? A.W.q_off : A.M.mp_off, 0, A.n_frames, q);
        const SbpQuery Q = sbp_query(A, q, f);
        uint32_t out[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
        int32_t outi[kTopK] = {-1, -1, -1, -1};
        if (Q.live) evals = sbp_scan_group<kSbpGroup>(A, Q, load_desc(A.generic ? A.W.desc : A.M.desc, q), out, outi);
        if ((threadIdx.x & (kSbpGroup - 1)) == 0) {
            *reinterpret_cast<uint4*>(A.topk_key + (long long)q * kTopK) = make_uint4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<int4*>(A.topk_idx + (long long)q * kTopK) = make_int4(outi[0], outi[1], outi[2], outi[3]);
            if (evals) atomicAdd(&s_evals, evals);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_evals) atomicAdd(A.evals, (unsigned long long)s_evals);
}

// Phase B, one warp per frame: the map points in vector order (:66), F.mvpMapPoints[bestIdx] = pMP (:149).
__global__ void __launch_bounds__(32) k_sbp_select(const __grid_constant__ SbpArgs A) {
    // The reference walks the map points in vector order and a match marks its key point taken for the points behind it (:108-110,
    // :142-151).  Only that mark is sequential, so a chunk of 32 map points is evaluated SPECULATIVELY, one per lane, against the
    // marks committed so far; a lane's result can only change if a point before it in the chunk takes the key point the lane found
    // best or second best.  The longest prefix of lanes without such a conflict is committed at once, the first conflicting lane
    // and everything behind it are evaluated again: one or two rounds per chunk instead of one step per map point.
    extern __shared__ uint32_t blocked[];
    const int lane = threadIdx.x, f = blockIdx.x;
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    const int k0 = A.F.kp_off[f], n = A.F.kp_off[f + 1] - k0;
    for (int i = lane; i < (n + 31) / 32; i += 32) blocked[i] = 0;
    if (A.kp_match) for (int i = lane; i < n; i += 32) A.kp_match[k0 + i] = -1;
    __syncwarp();
    const int q0 = A.M.mp_off[f], q1 = A.M.mp_off[f + 1];
    int nacc = 0, evals = 0;
    for (int base = q0; base < q1; base += 32) {
        const int q = base + lane;
        uint32_t k[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
        int ci[kTopK] = {-1, -1, -1, -1}, co[kTopK] = {-1, -1, -1, -1};
        int fl = 0;
        if (q < q1) {
            const uint4 v = *reinterpret_cast<const uint4*>(A.topk_key + (long long)q * kTopK);
            const int4 vi = *reinterpret_cast<const int4*>(A.topk_idx + (long long)q * kTopK);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
            ci[0] = vi.x; ci[1] = vi.y; ci[2] = vi.z; ci[3] = vi.w;
#pragma unroll
            for (int j = 0; j < kTopK; ++j)
                if (ci[j] >= 0) co[j] = A.F.keys[k0 + ci[j]].octave;
            fl = A.M.flags[q];
            if (k[0] == kEmptyKey) {   // nothing in the window: no walk step needed, the defaults are written lane-parallel
                if (A.mp_best_idx) A.mp_best_idx[q] = -1;
                if (A.mp_best_dist) A.mp_best_dist[q] = 256;
                if (A.mp_second_dist) A.mp_second_dist[q] = 256;
            }
        }
        // only queries that have at least one candidate take part (most local map points have none)
        unsigned pend = __ballot_sync(full, k[0] != kEmptyKey);
        while (pend) {
            const bool mine = (pend >> lane) & 1u;
            // ---- every pending lane: best / second among its list entries not taken so far
            int cnt = 0, bestIdx = -1, secondIdx = -1, d1 = 256, d2 = 256, l1 = -1, l2 = -1;
            bool complete = false;
            if (mine) {
#pragma unroll
                for (int j = 0; j < kTopK; ++j) {
                    if (cnt == 2 || complete) break;
                    if (k[j] == kEmptyKey) { complete = true; break; }
                    const int idx = ci[j];
                    if ((blocked[idx >> 5] >> (idx & 31)) & 1u) continue;
                    if (cnt == 0) { bestIdx = idx; d1 = (int)(k[j] >> kPosBits); l1 = co[j]; cnt = 1; }
                    else { secondIdx = idx; d2 = (int)(k[j] >> kPosBits); l2 = co[j]; cnt = 2; }
                }
                if (cnt == 2) complete = true;
            }
            const int first = __ffs(pend) - 1;
            if (!__shfl_sync(full, (int)complete, first)) {
                // the first pending lane has exhausted its truncated list: exact rescan with the mask, by the whole warp (its state
                // is final: nothing before it is pending)
                const int qq = base + first;
                const SbpQuery Q = sbp_query(A, qq, f);
                uint32_t out[kTopK];
                int32_t outi[kTopK];
                evals += sbp_scan_warp(A, Q, load_desc(A.M.desc, qq), blocked, out, outi);
                int bI = -1, e1 = 256, e2 = 256, m1 = -1, m2 = -1;
                if (out[0] != kEmptyKey) { bI = outi[0]; e1 = (int)(out[0] >> kPosBits); m1 = A.F.keys[k0 + outi[0]].octave; }
                if (out[1] != kEmptyKey) { e2 = (int)(out[1] >> kPosBits); m2 = A.F.keys[k0 + outi[1]].octave; }
                bool acc = false;
                if (bI >= 0 && e1 <= ORBGPU_TH_HIGH) acc = !(m1 == m2 && (float)e1 > __fmul_rn(A.nnratio, (float)e2));
                const int qfl = __shfl_sync(full, fl, first);
                if (lane == 0) {
                    if (A.mp_best_idx) A.mp_best_idx[qq] = bI;
                    if (A.mp_best_dist) A.mp_best_dist[qq] = e1;
                    if (A.mp_second_dist) A.mp_second_dist[qq] = e2;
                    if (acc) {
                        if (qfl & 4) blocked[bI >> 5] |= 1u << (bI & 31);
                        if (A.kp_match) atomicMax(&A.kp_match[k0 + bI], qq - q0);
                    }
                }
                nacc += acc;
                pend &= ~(1u << first);
                __syncwarp();
                continue;
            }
            bool accept = false;
            if (mine && bestIdx >= 0 && d1 <= ORBGPU_TH_HIGH) accept = !(l1 == l2 && (float)d1 > __fmul_rn(A.nnratio, (float)d2));
            // ---- conflicts: a pending lane before me takes my best or second key point (Observations() > 0 marks it taken), or I
            // am incomplete (my rescan must see everything before me committed)
            const bool claims = mine && accept && (fl & 4);
            bool stop = mine && !complete;
            unsigned cl = __ballot_sync(full, claims);
            while (cl) {
                const int i = __ffs(cl) - 1;
                cl &= cl - 1;
                const int c = __shfl_sync(full, bestIdx, i);
                if (mine && lane > i && (c == bestIdx || c == secondIdx)) stop = true;
            }
            const unsigned sb = __ballot_sync(full, stop);
            const unsigned done = sb ? (pend & ((1u << (__ffs(sb) - 1)) - 1u)) : pend;   // pending lanes before the first that must wait
            if ((done >> lane) & 1u) {
                if (A.mp_best_idx) A.mp_best_idx[q] = bestIdx;
                if (A.mp_best_dist) A.mp_best_dist[q] = d1;
                if (A.mp_second_dist) A.mp_second_dist[q] = d2;
                if (accept) {
                    if (fl & 4) atomicOr(&blocked[bestIdx >> 5], 1u << (bestIdx & 31));   // Observations() > 0: later candidates skip it
                    if (A.kp_match) atomicMax(&A.kp_match[k0 + bestIdx], q - q0);           // the last map point in vector order keeps the key point
                }
            }
            nacc += __popc(__ballot_sync(full, accept) & done);
            pend &= ~done;
            __syncwarp();
        }
    }
    (void)lt;
    if (lane == 0) {
        if (A.nmatches) A.nmatches[f] = nacc;
        if (evals) atomicAdd(A.evals, (unsigned long long)evals);
    }
}

// Phase B of the generic windowed search, one warp per frame: queries in order, best candidate only, accept when
// best <= th_dist, CurrentFrame.mvpMapPoints[bestIdx2] = pMP (:1644, :1789), rotation histogram over the accepted queries and
// removal of the matches outside the three main bins (:1663-1682: the key point is reset to NULL -> kp_match = -2).
// Best-only searches (Fuse, SearchBySim3): queries are independent, the winner is the head of the top-K list.
__global__ void __launch_bounds__(256) k_win_best(const __grid_constant__ SbpArgs A, int nq) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    const uint32_t key = A.topk_key[(long long)q * kTopK];
    A.mp_best_idx[q] = key == kEmptyKey ? -1 : A.topk_idx[(long long)q * kTopK];
    A.mp_best_dist[q] = key == kEmptyKey ? 256 : (int)(key >> kPosBits);
}

__global__ void __launch_bounds__(32) k_win_select(const __grid_constant__ SbpArgs A) {
    extern __shared__ uint32_t smem_b[];
    uint32_t* blocked = smem_b;
    int* hist = (int*)(smem_b + A.blocked_words);
    const int lane = threadIdx.x, f = blockIdx.x;
    const int k0 = A.F.kp_off[f], n = A.F.kp_off[f + 1] - k0;
    for (int i = lane; i < (n + 31) / 32; i += 32) blocked[i] = 0;
    hist[lane] = 0;
    if (A.kp_match) for (int i = lane; i < n; i += 32) A.kp_match[k0 + i] = -1;
    __syncwarp();
    const int q0 = A.W.q_off[f], q1 = A.W.q_off[f + 1];
    int nacc = 0, evals = 0;
    for (int base = q0; base < q1; base += 32) {
        const int q = base + lane;
        uint32_t k[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
        int ci[kTopK] = {-1, -1, -1, -1};
        float ca[kTopK] = {0.f, 0.f, 0.f, 0.f}, qa = 0.f;
        int fl = 0;
        if (q < q1) {
            const uint4 v = *reinterpret_cast<const uint4*>(A.topk_key + (long long)q * kTopK);
            const int4 vi = *reinterpret_cast<const int4*>(A.topk_idx + (long long)q * kTopK);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
            ci[0] = vi.x; ci[1] = vi.y; ci[2] = vi.z; ci[3] = vi.w;
            fl = A.W.flags[q];
            A.q_bin[q] = -1;
            if (A.check_orientation) {
                qa = A.W.angle[q];
#pragma unroll
                for (int j = 0; j < kTopK; ++j)
                    if (ci[j] >= 0) ca[j] = A.F.keys[k0 + ci[j]].angle;
            }
        }
        if (q < q1 && k[0] == kEmptyKey) {   // nothing in the window: defaults lane-parallel, no walk step
            if (A.mp_best_idx) A.mp_best_idx[q] = -1;
            if (A.mp_best_dist) A.mp_best_dist[q] = 256;
        }
        // speculative walk (see k_sbp_select): best-only, so a lane waits only if an accepted lane before it takes its best key point
        unsigned pend = __ballot_sync(0xffffffffu, k[0] != kEmptyKey);
        while (pend) {
            const bool mine = (pend >> lane) & 1u;
            int bestIdx = -1, d1 = 256;
            float bang = 0.f;
            bool complete = false;
            if (mine) {
#pragma unroll
                for (int j = 0; j < kTopK; ++j) {
                    if (bestIdx >= 0 || complete) break;
                    if (k[j] == kEmptyKey) { complete = true; break; }
                    const int idx = ci[j];
                    if ((blocked[idx >> 5] >> (idx & 31)) & 1u) continue;
                    bestIdx = idx; d1 = (int)(k[j] >> kPosBits); bang = ca[j];
                }
            }
            const bool rescan = mine && bestIdx < 0 && !complete;
            const int first = __ffs(pend) - 1;
            if (__shfl_sync(0xffffffffu, (int)rescan, first)) {
                // the first pending lane's truncated list ran out: exact rescan with the mask, by the whole warp
                const int qq = base + first;
                const int qfl = __shfl_sync(0xffffffffu, fl, first);
                const float qang = __shfl_sync(0xffffffffu, qa, first);
                const SbpQuery Q = sbp_query(A, qq, f);
                uint32_t out[kTopK];
                int32_t outi[kTopK];
                evals += sbp_scan_warp(A, Q, load_desc(A.W.desc, qq), blocked, out, outi);
                int bI = -1, e1 = 256;
                if (out[0] != kEmptyKey) { bI = outi[0]; e1 = (int)(out[0] >> kPosBits); }
                const bool acc = bI >= 0 && e1 <= A.th_dist;
                if (acc) {
                    int bin = 0;
                    if (A.check_orientation) bin = rot_bin(qang, A.F.keys[k0 + bI].angle);
                    if (lane == 0) {
                        if ((qfl & 4) || A.skip_any) blocked[bI >> 5] |= 1u << (bI & 31);
                        if (A.kp_match) atomicMax(&A.kp_match[k0 + bI], qq - q0);
                        A.q_bin[qq] = (int8_t)bin;
                        hist[bin] += 1;
                    }
                }
                if (lane == 0) {
                    if (A.mp_best_idx) A.mp_best_idx[qq] = bI;
                    if (A.mp_best_dist) A.mp_best_dist[qq] = e1;
                }
                nacc += acc;
                pend &= ~(1u << first);
                __syncwarp();
                continue;
            }
            const bool accept = mine && bestIdx >= 0 && d1 <= A.th_dist;
            const bool claims = accept && ((fl & 4) || A.skip_any);   // the key point now holds pMP: later candidates skip it
            bool stop = rescan;
            unsigned cl = __ballot_sync(0xffffffffu, claims);
            while (cl) {
                const int i = __ffs(cl) - 1;
                cl &= cl - 1;
                const int cb = __shfl_sync(0xffffffffu, bestIdx, i);
                if (mine && lane > i && cb == bestIdx) stop = true;
            }
            const unsigned sb = __ballot_sync(0xffffffffu, stop);
            const unsigned done = sb ? (pend & ((1u << (__ffs(sb) - 1)) - 1u)) : pend;
            if ((done >> lane) & 1u) {
                if (accept) {
                    int bin = 0;
                    if (A.check_orientation) bin = rot_bin(qa, bang);
                    if (claims) atomicOr(&blocked[bestIdx >> 5], 1u << (bestIdx & 31));
                    if (A.kp_match) atomicMax(&A.kp_match[k0 + bestIdx], q - q0);   // the last query in vector order keeps the key point
                    A.q_bin[q] = (int8_t)bin;
                    atomicAdd(&hist[bin], 1);
                }
                if (A.mp_best_idx) A.mp_best_idx[q] = bestIdx;
                if (A.mp_best_dist) A.mp_best_dist[q] = d1;
            }
            nacc += __popc(__ballot_sync(0xffffffffu, accept) & done);
            pend &= ~done;
            __syncwarp();
        }
    }
    __syncwarp();
    int removed = 0;
    if (A.check_orientation) {
        int ind1, ind2, ind3;
        three_maxima(hist, kHisto, ind1, ind2, ind3);
        for (int q = q0 + lane; q < q1; q += 32) {
            const int bin = A.q_bin[q];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) {
                if (A.kp_match && A.mp_best_idx) A.kp_match[k0 + A.mp_best_idx[q]] = -2;
                ++removed;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
    }
    if (lane == 0) {
        if (A.nmatches) A.nmatches[f] = nacc - removed;
        if (evals) atomicAdd(A.evals, (unsigned long long)evals);
    }
}

// Phase B of SearchForInitialization (ORBmatcher.cc:493-632), one warp per frame pair, queries (key points of frame 1) in
// index order.  The dynamic state is not a taken mask but vMatchedDistance: a candidate is skipped when it already holds a
// match of the same or a smaller distance (:546) — for the best AND the second best; a better match steals the key point
// from its earlier owner (:573-583).  rotHist keeps the stolen entries (:592-603), so the bin counts do too.
// Shared: mdist[n2] (u16, 0xFFFF = INT_MAX) | owner[n2] (i32) | hist[32].
__global__ void __launch_bounds__(32) k_init_select(const __grid_constant__ SbpArgs A, int32_t* __restrict__ match12) {
    extern __shared__ uint32_t smem_i[];
    const int lane = threadIdx.x, f = blockIdx.x;
    const int k0 = A.F.kp_off[f], n = A.F.kp_off[f + 1] - k0;
    int* owner = reinterpret_cast<int*>(smem_i);
    int* hist = owner + A.blocked_words * 32;
    uint16_t* mdist = reinterpret_cast<uint16_t*>(hist + 32);
    for (int i = lane; i < n; i += 32) { owner[i] = -1; mdist[i] = 0xFFFFu; }
    hist[lane] = 0;
    const int q0 = A.W.q_off[f], q1 = A.W.q_off[f + 1];
    for (int q = q0 + lane; q < q1; q += 32) { match12[q] = -1; A.q_bin[q] = -1; }
    __syncwarp();
    int nmatches = 0, evals = 0;
    for (int base = q0; base < q1; base += 32) {
        const int q = base + lane;
        uint32_t k[kTopK] = {kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey};
        int ci[kTopK] = {-1, -1, -1, -1};
        if (q < q1) {
            const uint4 v = *reinterpret_cast<const uint4*>(A.topk_key + (long long)q * kTopK);
            const int4 vi = *reinterpret_cast<const int4*>(A.topk_idx + (long long)q * kTopK);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
            ci[0] = vi.x; ci[1] = vi.y; ci[2] = vi.z; ci[3] = vi.w;
        }
        const int cnt_chunk = min(32, q1 - base);
        for (int l = 0; l < cnt_chunk; ++l) {
            uint32_t kk[kTopK];
            int cc[kTopK];
#pragma unroll
            for (int j = 0; j < kTopK; ++j) {
                kk[j] = __shfl_sync(0xffffffffu, k[j], l);
                cc[j] = __shfl_sync(0xffffffffu, ci[j], l);
            }
            const int qq = base + l;
            int cnt = 0, bestIdx = -1, d1 = 0x7fffffff, d2 = 0x7fffffff;
            bool complete = false;
#pragma unroll
            for (int j = 0; j < kTopK; ++j) {
                if (cnt == 2 || complete) break;
                if (kk[j] == kEmptyKey) { complete = true; break; }
                const int idx = cc[j], dist = (int)(kk[j] >> kPosBits);
                if ((int)mdist[idx] <= dist) continue;
                if (cnt == 0) { bestIdx = idx; d1 = dist; cnt = 1; }
                else { d2 = dist; cnt = 2; }
            }
            if (cnt == 2) complete = true;
            if (!complete) {   // truncated list ran dry: exact rescan under the current vMatchedDistance
                const SbpQuery Q = sbp_query(A, qq, f);
                uint32_t out[kTopK];
                int32_t outi[kTopK];
                evals += sbp_scan_warp(A, Q, load_desc(A.W.desc, qq), nullptr, out, outi, mdist);
                bestIdx = -1; d1 = 0x7fffffff; d2 = 0x7fffffff;
                if (out[0] != kEmptyKey) { bestIdx = outi[0]; d1 = (int)(out[0] >> kPosBits); }
                if (out[1] != kEmptyKey) d2 = (int)(out[1] >> kPosBits);
            }
            // bestDist <= TH_LOW && bestDist < (float)bestDist2 * mfNNratio (:561-563)
            const bool accept = bestIdx >= 0 && d1 <= A.th_dist && (float)d1 < __fmul_rn((float)d2, A.nnratio);
            if (accept) {
                const int prev = owner[bestIdx];
                int bin = 0;
                if (A.check_orientation) bin = rot_bin(A.W.angle[qq], A.F.keys[k0 + bestIdx].angle);
                if (lane == 0) {
                    if (prev >= 0) match12[q0 + prev] = -1;            // the earlier owner loses the key point (:573-577)
                    match12[qq] = bestIdx;
                    owner[bestIdx] = qq - q0;
                    mdist[bestIdx] = (uint16_t)d1;
                    if (A.check_orientation) { A.q_bin[qq] = (int8_t)bin; hist[bin] += 1; }
                }
                nmatches += prev >= 0 ? 0 : 1;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    int removed = 0;
    if (A.check_orientation) {
        int ind1, ind2, ind3;
        three_maxima(hist, kHisto, ind1, ind2, ind3);
        for (int q = q0 + lane; q < q1; q += 32) {
            const int bin = A.q_bin[q];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3 && match12[q] >= 0) {   // stolen entries are already -1 (:617-621)
                match12[q] = -1;
                ++removed;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
    }
    if (lane == 0) {
        if (A.nmatches) A.nmatches[f] = nmatches - removed;
        if (evals) atomicAdd(A.evals, (unsigned long long)evals);
    }
}

}  // namespace og

// ============================================================================================================
// Host side: handles, uploads, launch sequences, C ABI
// ============================================================================================================
namespace {

#define OGM_CUDA(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return og_fail(ORBGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
    } while (0)

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t grab(size_t bytes, void** out) {
        bytes = std::max<size_t>(bytes, 256);
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            cap = 0;
            const size_t want = bytes + bytes / 4;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) return e;
            cap = want;
        }
        *out = p;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct orbgpu_frame_set_dev {
    int device = 0;
    int n_frames = 0, nkp = 0, nnodes = 0, nfeat = 0, max_kp = 0, max_nodes = 0;
    bool has_fv = false, has_grid = false;
    std::vector<int32_t> h_kp_off;      // [n_frames+1]
    std::vector<int32_t> h_node_cnt;    // [n_frames] nodes per frame
    std::vector<int32_t> h_ent_cnt;     // [n_frames] fv_feat entries per frame
    og::FrameSetView v{};
    std::vector<void*> owned;
    // blocks borrowed from the matcher's arena (orbgpu_frame_set_from_extraction): returned to it on release, not freed
    orbgpu_matcher* arena_owner = nullptr;
    std::vector<std::pair<void*, size_t>> borrowed;
};

struct orbgpu_mappoint_set_dev {
    int device = 0;
    int n_frames = 0, nmp = 0;
    og::MapPointView v{};
    std::vector<void*> owned;
};

struct orbgpu_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_stage = nullptr;
    bool stage_busy = false;
    int dense_q = 512, dense_c = 256, tile_rq = 4;
    int sm_count = 148;
    int last_launches = 0;
    unsigned long long* d_evals = nullptr;
    int* d_nitems = nullptr;
    // grow-only scratch
    Scratch s_ctrl, s_topk, s_topk2, s_bin, s_items, s_grid_start, s_grid_items, s_res1, s_res2, s_out[5], s_in[2];
    Scratch s_tmp[32];            // device copies of the caller's arrays in the host-pointer variants (grow-only, reused)
    int tmp_next = 0;
    uint8_t* h_stage = nullptr;   // pinned staging for the per-call control arrays
    size_t h_stage_cap = 0;
    // free blocks of released extraction-built frame sets: a pipeline that builds a frame set per batch reuses them instead
    // of paying cudaMalloc / cudaFree (which synchronises the device) every step
    std::vector<std::pair<void*, size_t>> arena_free;
    cudaError_t arena_take(size_t bytes, void** out, size_t* cap) {
        bytes = std::max<size_t>(bytes, 256);
        int best = -1;
        for (int i = 0; i < (int)arena_free.size(); ++i)
            if (arena_free[i].second >= bytes && (best < 0 || arena_free[i].second < arena_free[best].second)) best = i;
        if (best >= 0 && arena_free[best].second <= 4 * bytes + (1 << 20)) {
            *out = arena_free[best].first;
            *cap = arena_free[best].second;
            arena_free.erase(arena_free.begin() + best);
            return cudaSuccess;
        }
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(out, want);
        *cap = want;
        return e;
    }
};

namespace {

// `pool` = the matcher whose scratch slots back the copy (temporary sets of the host-pointer variants), or null for an
// allocation owned by the set (uploaded handles).
template <class T>
int upload_array(const T* host, size_t n, std::vector<void*>& owned, const T** out, cudaStream_t st, orbgpu_matcher* pool = nullptr) {
    *out = nullptr;
    if (!host || n == 0) return ORBGPU_OK;
    void* d = nullptr;
    if (pool && pool->tmp_next < 32) {
        OGM_CUDA(pool->s_tmp[pool->tmp_next++].grab(n * sizeof(T), &d));
    } else {
        OGM_CUDA(cudaMalloc(&d, n * sizeof(T)));
        owned.push_back(d);
    }
    OGM_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, st));
    *out = (const T*)d;
    return ORBGPU_OK;
}

int stage_reserve(orbgpu_matcher* m, size_t bytes) {
    if (m->stage_busy) {
        OGM_CUDA(cudaEventSynchronize(m->ev_stage));
        m->stage_busy = false;
    }
    if (bytes > m->h_stage_cap) {
        if (m->h_stage) cudaFreeHost(m->h_stage);
        m->h_stage = nullptr;
        m->h_stage_cap = 0;
        OGM_CUDA(cudaMallocHost((void**)&m->h_stage, bytes + bytes / 2 + 4096));
        m->h_stage_cap = bytes + bytes / 2 + 4096;
    }
    return ORBGPU_OK;
}

// Packs several small host arrays into the pinned staging buffer and uploads them with one copy.
struct CtrlPack {
    orbgpu_matcher* m;
    struct Item { const void* src; size_t bytes; size_t off; };
    std::vector<Item> items;
    size_t total = 0;
    explicit CtrlPack(orbgpu_matcher* mm) : m(mm) {}
    size_t add(const void* src, size_t bytes) {
        const size_t off = total;
        items.push_back({src, bytes, off});
        total += (bytes + 255) & ~size_t(255);
        return off;
    }
    int upload(uint8_t** d_base) {
        int rc = stage_reserve(m, total);
        if (rc) return rc;
        for (auto& it : items) memcpy(m->h_stage + it.off, it.src, it.bytes);
        void* d = nullptr;
        OGM_CUDA(m->s_ctrl.grab(total, &d));
        OGM_CUDA(cudaMemcpyAsync(d, m->h_stage, total, cudaMemcpyHostToDevice, m->stream));
        OGM_CUDA(cudaEventRecord(m->ev_stage, m->stream));
        m->stage_busy = true;
        *d_base = (uint8_t*)d;
        return ORBGPU_OK;
    }
};

int check_matcher(orbgpu_matcher* m) {
    if (!m) return og_fail(ORBGPU_ERR_ARG, "null matcher");
    OGM_CUDA(cudaSetDevice(m->device));
    return ORBGPU_OK;
}

int build_frame_set(orbgpu_matcher* m, const orbgpu_frame_set* h, orbgpu_frame_set_dev* fs, orbgpu_matcher* pool = nullptr) {
    if (!h || h->n_frames < 0 || !h->kp_off || !h->desc) return og_fail(ORBGPU_ERR_ARG, "frame set: null kp_off/desc");
    fs->device = m->device;
    fs->n_frames = h->n_frames;
    fs->h_kp_off.assign(h->kp_off, h->kp_off + h->n_frames + 1);
    fs->nkp = h->kp_off[h->n_frames];
    for (int f = 0; f < h->n_frames; ++f) {
        const int n = h->kp_off[f + 1] - h->kp_off[f];
        if (n < 0) return og_fail(ORBGPU_ERR_ARG, "frame set: kp_off not ascending");
        fs->max_kp = std::max(fs->max_kp, n);
    }
    if (fs->max_kp > (int)og::kPosMask) return og_fail(ORBGPU_ERR_ARG, "frame set: more than 2^20-1 keypoints in one frame");
    cudaStream_t st = m->stream;
    int rc;
    if ((rc = upload_array(h->kp_off, (size_t)h->n_frames + 1, fs->owned, &fs->v.kp_off, st, pool))) return rc;
    if ((rc = upload_array((const og::KeyPoint*)h->keys_un, (size_t)fs->nkp, fs->owned, &fs->v.keys, st, pool))) return rc;
    if ((rc = upload_array(h->desc, (size_t)fs->nkp * 32, fs->owned, &fs->v.desc, st, pool))) return rc;
    if ((rc = upload_array(h->u_right, (size_t)fs->nkp, fs->owned, &fs->v.u_right, st, pool))) return rc;
    if ((rc = upload_array(h->kp_flags, (size_t)fs->nkp, fs->owned, &fs->v.flags, st, pool))) return rc;
    if (h->grid) {
        fs->has_grid = true;
        if ((rc = upload_array(h->grid, (size_t)h->n_frames * 4, fs->owned, &fs->v.grid, st, pool))) return rc;
    }
    if (h->fv_node_off) {
        if (!h->fv_node_id || !h->fv_feat_off || !h->fv_feat) return og_fail(ORBGPU_ERR_ARG, "frame set: incomplete FeatureVector arrays");
        fs->has_fv = true;
        fs->nnodes = h->fv_node_off[h->n_frames];
        fs->nfeat = h->fv_feat_off[fs->nnodes];
        fs->h_node_cnt.resize(h->n_frames);
        fs->h_ent_cnt.resize(h->n_frames);
        for (int f = 0; f < h->n_frames; ++f) {
            const int a0 = h->fv_node_off[f], a1 = h->fv_node_off[f + 1];
            fs->h_node_cnt[f] = a1 - a0;
            fs->h_ent_cnt[f] = h->fv_feat_off[a1] - h->fv_feat_off[a0];
            fs->max_nodes = std::max(fs->max_nodes, a1 - a0);
            if (a1 < a0 || fs->h_ent_cnt[f] < 0) return og_fail(ORBGPU_ERR_ARG, "frame set: FeatureVector offsets not ascending");
        }
        if ((rc = upload_array(h->fv_node_off, (size_t)h->n_frames + 1, fs->owned, &fs->v.node_off, st, pool))) return rc;
        if ((rc = upload_array(h->fv_node_id, (size_t)fs->nnodes, fs->owned, &fs->v.node_id, st, pool))) return rc;
        if ((rc = upload_array(h->fv_feat_off, (size_t)fs->nnodes + 1, fs->owned, &fs->v.feat_off, st, pool))) return rc;
        if ((rc = upload_array(h->fv_feat, (size_t)fs->nfeat, fs->owned, &fs->v.feat, st, pool))) return rc;
    }
    return ORBGPU_OK;
}

void free_owned(std::vector<void*>& owned) {
    for (void* p : owned) cudaFree(p);
    owned.clear();
}

struct PairPlan {
    std::vector<long long> entry_off;
    long long total_entries = 0;
    long long items_cap = 0;
    int max_nb = 0, max_nodes1 = 0;
};

int plan_pairs(const orbgpu_frame_set_dev* s1, const orbgpu_frame_set_dev* s2, int n_pairs, const int32_t* idx1, const int32_t* idx2,
               int qt, PairPlan& P) {
    if (!s1 || !s2 || !s1->has_fv || !s2->has_fv) return og_fail(ORBGPU_ERR_ARG, "node scans need frame sets with FeatureVector arrays");
    if (n_pairs < 0 || (n_pairs && (!idx1 || !idx2))) return og_fail(ORBGPU_ERR_ARG, "bad pair arrays");
    P.entry_off.resize((size_t)n_pairs + 1);
    long long acc = 0, items = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const int fa = idx1[p], fb = idx2[p];
        if (fa < 0 || fa >= s1->n_frames || fb < 0 || fb >= s2->n_frames) return og_fail(ORBGPU_ERR_ARG, "pair index out of range");
        P.entry_off[p] = acc;
        acc += s1->h_ent_cnt[fa];
        items += s1->h_node_cnt[fa] + s1->h_ent_cnt[fa] / qt;
        P.max_nb = std::max(P.max_nb, s2->h_kp_off[fb + 1] - s2->h_kp_off[fb]);
        P.max_nodes1 = std::max(P.max_nodes1, s1->h_node_cnt[fa]);
    }
    P.entry_off[n_pairs] = acc;
    P.total_entries = acc;
    P.items_cap = items + 1;
    return ORBGPU_OK;
}

int bow_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* s1, const orbgpu_frame_set_dev* s2, int n_pairs, const int32_t* idx1,
            const int32_t* idx2, float nnratio, int check_orientation, int th_low, int th_inclusive, int require_mp2,
            const int64_t* match_off, int32_t* d_match12, int32_t* d_match_dist, int32_t* d_nmatches) {
    using namespace og;
    m->last_launches = 0;
    if (n_pairs == 0) return ORBGPU_OK;
    if (!match_off || !d_match12) return og_fail(ORBGPU_ERR_ARG, "null match_off/match12");
    const int rq = m->tile_rq;
    PairPlan P;
    int rc = plan_pairs(s1, s2, n_pairs, idx1, idx2, kTileThreads * rq, P);
    if (rc) return rc;
    if (P.items_cap > 0x7fffffffLL) return og_fail(ORBGPU_ERR_CAPACITY, "too many node pairs in one call");
    CtrlPack C(m);
    const size_t o1 = C.add(idx1, (size_t)n_pairs * 4), o2 = C.add(idx2, (size_t)n_pairs * 4);
    const size_t oe = C.add(P.entry_off.data(), ((size_t)n_pairs + 1) * 8), om = C.add(match_off, (size_t)n_pairs * 8);
    uint8_t* dc = nullptr;
    if ((rc = C.upload(&dc))) return rc;

    BowArgs A;
    A.S1 = s1->v; A.S2 = s2->v;
    A.n_pairs = n_pairs;
    A.idx1 = (const int32_t*)(dc + o1); A.idx2 = (const int32_t*)(dc + o2);
    A.entry_off = (const long long*)(dc + oe); A.match_off = (const long long*)(dc + om);
    A.require_mp2 = require_mp2;
    A.dense_q = m->dense_q; A.dense_c = m->dense_c;
    // A best above th_low is rejected whatever follows, and a second best s > th_low / nnratio passes the ratio test
    // (best < nnratio * s) for every admissible best, exactly like "no second best" (256): such candidates need not be
    // listed.  +2 keeps the bound clear of float rounding in nnratio * s.
    {
        const double bound = nnratio > 0.f ? std::floor((double)th_low / (double)nnratio) + 2.0 : 256.0;
        A.dmax = (int)std::min(256.0, std::max((double)th_low, bound));
    }
    void* ptr = nullptr;
    OGM_CUDA(m->s_topk.grab((size_t)std::max<long long>(P.total_entries, 1) * kTopK * 4, &ptr));
    A.topk = (uint32_t*)ptr;
    OGM_CUDA(m->s_items.grab((size_t)P.items_cap * sizeof(int4), &ptr));
    A.items = (int4*)ptr;
    A.items_cap = (int)P.items_cap;
    A.n_items = m->d_nitems;
    A.n_sparse = m->dense_q > 0 ? m->d_nitems + 1 : nullptr;
    A.evals = m->d_evals;
    SelectArgs Z;
    Z.nnratio = nnratio; Z.check_orientation = check_orientation; Z.th_low = th_low; Z.th_inclusive = th_inclusive;
    OGM_CUDA(m->s_bin.grab((size_t)std::max<long long>(P.total_entries, 1), &ptr));
    Z.entry_bin = (int8_t*)ptr;
    Z.match12 = d_match12; Z.match_dist = d_match_dist; Z.nmatches = d_nmatches;
    Z.taken_words = (P.max_nb + 31) / 32 + 1;
    const size_t smem = ((size_t)Z.taken_words + 32) * 4 * kSelectWarps;
    if (smem > 200 * 1024) return og_fail(ORBGPU_ERR_CAPACITY, "frame too large for the shared-memory taken mask");

    cudaStream_t st = m->stream;
    OGM_CUDA(cudaEventRecord(m->ev0, st));
    OGM_CUDA(cudaMemsetAsync(m->d_evals, 0, 8, st));
    OGM_CUDA(cudaMemsetAsync(m->d_nitems, 0, 8, st));
    if (P.total_entries > 0) {
        if (m->dense_q > 0) {
            dim3 pg(n_pairs, (P.max_nodes1 + 127) / 128);
            if (rq == 8) k_bow_plan<8><<<pg, 128, 0, st>>>(A); else k_bow_plan<4><<<pg, 128, 0, st>>>(A);
            const int grid = (int)std::min<long long>(P.items_cap, (long long)m->sm_count * 4);
            if (rq == 8) k_bow_topk_tile<8><<<grid, kTileThreads, 0, st>>>(A); else k_bow_topk_tile<4><<<grid, kTileThreads, 0, st>>>(A);
            m->last_launches += 2;
        }
        const long long blocks = std::min<long long>((P.total_entries + kWarpsPerBlock - 1) / kWarpsPerBlock, (long long)m->sm_count * 64);
        k_bow_topk_warp<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(A, P.total_entries);
        m->last_launches += 1;
    }
    if (smem > 48 * 1024) OGM_CUDA(cudaFuncSetAttribute(k_bow_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bow_select<<<(n_pairs + kSelectWarps - 1) / kSelectWarps, kSelectWarps * 32, smem, st>>>(A, Z);
    m->last_launches += 1;
    OGM_CUDA(cudaEventRecord(m->ev1, st));
    OGM_CUDA(cudaGetLastError());
    return ORBGPU_OK;
}

int tri_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* s1, const orbgpu_frame_set_dev* s2, int n_pairs, const int32_t* idx1,
            const int32_t* idx2, const float* f12, const float* epipole, const float* scale_factors, const float* level_sigma2,
            int n_levels, int only_stereo, int check_orientation, const int64_t* match_off, int32_t* d_match12,
            int32_t* d_match_dist, int32_t* d_nmatches) {
    using namespace og;
    m->last_launches = 0;
    if (n_pairs == 0) return ORBGPU_OK;
    if (!match_off || !d_match12 || !f12 || !epipole || !scale_factors || !level_sigma2 || n_levels < 1)
        return og_fail(ORBGPU_ERR_ARG, "null argument");
    PairPlan P;
    int rc = plan_pairs(s1, s2, n_pairs, idx1, idx2, 1 << 30, P);
    if (rc) return rc;
    CtrlPack C(m);
    const size_t o1 = C.add(idx1, (size_t)n_pairs * 4), o2 = C.add(idx2, (size_t)n_pairs * 4);
    const size_t oe = C.add(P.entry_off.data(), ((size_t)n_pairs + 1) * 8), om = C.add(match_off, (size_t)n_pairs * 8);
    const size_t of = C.add(f12, (size_t)n_pairs * 36), op = C.add(epipole, (size_t)n_pairs * 8);
    const size_t os = C.add(scale_factors, (size_t)n_levels * 4), og2 = C.add(level_sigma2, (size_t)n_levels * 4);
    uint8_t* dc = nullptr;
    if ((rc = C.upload(&dc))) return rc;
    TriArgs A;
    A.S1 = s1->v; A.S2 = s2->v;
    A.n_pairs = n_pairs;
    A.idx1 = (const int32_t*)(dc + o1); A.idx2 = (const int32_t*)(dc + o2);
    A.entry_off = (const long long*)(dc + oe); A.match_off = (const long long*)(dc + om);
    A.f12 = (const float*)(dc + of); A.epipole = (const float*)(dc + op);
    A.scale = (const float*)(dc + os); A.sigma2 = (const float*)(dc + og2);
    A.only_stereo = only_stereo; A.check_orientation = check_orientation;
    void* ptr = nullptr;
    OGM_CUDA(m->s_res1.grab((size_t)std::max<long long>(P.total_entries, 1) * 4, &ptr));
    A.res_idx2 = (int32_t*)ptr;
    OGM_CUDA(m->s_res2.grab((size_t)std::max<long long>(P.total_entries, 1) * 4, &ptr));
    A.res_dist = (int32_t*)ptr;
    A.match12 = d_match12; A.match_dist = d_match_dist; A.nmatches = d_nmatches;
    A.evals = m->d_evals;
    cudaStream_t st = m->stream;
    OGM_CUDA(cudaEventRecord(m->ev0, st));
    OGM_CUDA(cudaMemsetAsync(m->d_evals, 0, 8, st));
    if (P.total_entries > 0) {
        const long long blocks = (P.total_entries + kWarpsPerBlock - 1) / kWarpsPerBlock;
        if (blocks > 0x7fffffffLL) return og_fail(ORBGPU_ERR_CAPACITY, "too many queries in one call");
        k_tri_scan<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(A, P.total_entries);
        m->last_launches += 1;
    }
    k_tri_finalize<<<n_pairs, 32, 0, st>>>(A);
    m->last_launches += 1;
    OGM_CUDA(cudaEventRecord(m->ev1, st));
    OGM_CUDA(cudaGetLastError());
    return ORBGPU_OK;
}

int sbp_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* fs, const orbgpu_mappoint_set_dev* mp, const float* scale_factors,
            int n_levels, float th, float nnratio, int32_t* d_kp_match, int32_t* d_best_idx, int32_t* d_best_dist,
            int32_t* d_second_dist, int32_t* d_nmatches) {
    using namespace og;
    m->last_launches = 0;
    if (!fs || !mp || !scale_factors || n_levels < 1) return og_fail(ORBGPU_ERR_ARG, "null argument");
    if (!fs->has_grid) return og_fail(ORBGPU_ERR_ARG, "projection search needs the frame set's grid parameters");
    if (fs->n_frames != mp->n_frames) return og_fail(ORBGPU_ERR_ARG, "frame set and map-point set disagree on n_frames");
    if (fs->n_frames == 0) return ORBGPU_OK;
    CtrlPack C(m);
    const size_t os = C.add(scale_factors, (size_t)n_levels * 4);
    uint8_t* dc = nullptr;
    int rc = C.upload(&dc);
    if (rc) return rc;
    SbpArgs A;
    memset(&A, 0, sizeof(A));
    A.F = fs->v; A.M = mp->v;
    A.n_frames = fs->n_frames;
    A.scale = (const float*)(dc + os);
    A.th = th; A.nnratio = nnratio;
    void* ptr = nullptr;
    OGM_CUDA(m->s_grid_start.grab((size_t)fs->n_frames * (kGridCells + 1) * 4, &ptr));
    A.cell_start = (int32_t*)ptr;
    OGM_CUDA(m->s_grid_items.grab((size_t)std::max(fs->nkp, 1) * 4, &ptr));
    A.cell_items = (int32_t*)ptr;
    OGM_CUDA(m->s_topk.grab((size_t)std::max(mp->nmp, 1) * kTopK * 4, &ptr));
    A.topk_key = (uint32_t*)ptr;
    OGM_CUDA(m->s_topk2.grab((size_t)std::max(mp->nmp, 1) * kTopK * 4, &ptr));
    A.topk_idx = (int32_t*)ptr;
    A.kp_match = d_kp_match; A.mp_best_idx = d_best_idx; A.mp_best_dist = d_best_dist; A.mp_second_dist = d_second_dist;
    A.nmatches = d_nmatches;
    A.blocked_words = (fs->max_kp + 31) / 32 + 1;
    A.evals = m->d_evals;
    const size_t smem = (size_t)A.blocked_words * 4;
    if (smem > 200 * 1024) return og_fail(ORBGPU_ERR_CAPACITY, "frame too large for the shared-memory blocked mask");
    cudaStream_t st = m->stream;
    OGM_CUDA(cudaEventRecord(m->ev0, st));
    OGM_CUDA(cudaMemsetAsync(m->d_evals, 0, 8, st));
    k_grid_build<<<fs->n_frames, kGridThreads, 0, st>>>(A);
    m->last_launches += 1;
    if (mp->nmp > 0) {
        // ORBGPU_SBP_WARP=1: a whole warp per map point in phase A (development)
        static const int sbp_warp_env = []() { const char* e = getenv("ORBGPU_SBP_WARP"); return e ? atoi(e) : 0; }();
        if (sbp_warp_env) {
            k_sbp_topk<<<(mp->nmp + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(A, mp->nmp);
        } else {
            if (A.th <= 2.0f) {
                const int per_cta = kWarpsPerBlock * 32 / 4;
                k_sbp_topk_g<4><<<(mp->nmp + per_cta - 1) / per_cta, kWarpsPerBlock * 32, 0, st>>>(A, mp->nmp);
            } else {
                const int per_cta = kWarpsPerBlock * 32 / 8;
                k_sbp_topk_g<8><<<(mp->nmp + per_cta - 1) / per_cta, kWarpsPerBlock * 32, 0, st>>>(A, mp->nmp);
            }
        }
        m->last_launches += 1;
    }
    if (smem > 48 * 1024) OGM_CUDA(cudaFuncSetAttribute(k_sbp_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sbp_select<<<fs->n_frames, 32, smem, st>>>(A);
    m->last_launches += 1;
    OGM_CUDA(cudaEventRecord(m->ev1, st));
    OGM_CUDA(cudaGetLastError());
    return ORBGPU_OK;
}

// Generic windowed search (host pointers): the query arrays are uploaded into pooled scratch, the frame set likewise.
int win_host(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_window_query_set* qs, int th_dist, int skip_any,
             int check_orientation, int32_t* kp_match, int32_t* q_best_idx, int32_t* q_best_dist, int32_t* nmatches, bool best_only = false,
             const float* inv_sigma2 = nullptr, int n_levels = 0, bool init_mode = false, float nnratio = 0.f, int32_t* match12 = nullptr) {
    using namespace og;
    m->last_launches = 0;
    if (!frames || !qs || !qs->q_off) return og_fail(ORBGPU_ERR_ARG, "null argument");
    if (!frames->grid) return og_fail(ORBGPU_ERR_ARG, "windowed search needs the frame set's grid parameters");
    const int nf = frames->n_frames;
    if (nf == 0) return ORBGPU_OK;
    const int nq = qs->q_off[nf];
    if (nq && (!qs->u || !qs->v || !qs->radius || !qs->min_level || !qs->max_level || !qs->flags || !qs->desc || (check_orientation && !qs->angle)))
        return og_fail(ORBGPU_ERR_ARG, "window query set: null array");
    orbgpu_frame_set_dev F;
    std::vector<void*> owned;
    m->tmp_next = 0;
    int rc = build_frame_set(m, frames, &F, m);
    SbpArgs A;
    memset(&A, 0, sizeof(A));
    cudaStream_t st = m->stream;
    if (!rc) rc = upload_array(qs->q_off, (size_t)nf + 1, owned, &A.W.q_off, st, m);
    if (!rc) rc = upload_array(qs->u, (size_t)nq, owned, &A.W.u, st, m);
    if (!rc) rc = upload_array(qs->v, (size_t)nq, owned, &A.W.v, st, m);
    if (!rc) rc = upload_array(qs->radius, (size_t)nq, owned, &A.W.radius, st, m);
    if (!rc) rc = upload_array(qs->ur, (size_t)nq, owned, &A.W.ur, st, m);
    if (!rc) rc = upload_array(qs->angle, (size_t)nq, owned, &A.W.angle, st, m);
    if (!rc) rc = upload_array(qs->min_level, (size_t)nq, owned, &A.W.min_level, st, m);
    if (!rc) rc = upload_array(qs->max_level, (size_t)nq, owned, &A.W.max_level, st, m);
    if (!rc) rc = upload_array(qs->flags, (size_t)nq, owned, &A.W.flags, st, m);
    if (!rc) rc = upload_array(qs->desc, (size_t)nq * 32, owned, &A.W.desc, st, m);
    if (!rc && inv_sigma2) rc = upload_array(inv_sigma2, (size_t)n_levels, owned, &A.inv_sigma2, st, m);
    A.no_xr_window = (best_only || init_mode) ? 1 : 0;
    A.nnratio = nnratio;
    void* d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void* ptr = nullptr;
    cudaError_t ce = cudaSuccess;
    if (!rc) {
        const size_t sz[5] = {(size_t)std::max(F.nkp, 1) * 4, (size_t)std::max(nq, 1) * 4, (size_t)std::max(nq, 1) * 4, (size_t)std::max(nq, 1), (size_t)nf * 4};
        for (int i = 0; i < 5 && ce == cudaSuccess; ++i) ce = m->s_out[i].grab(sz[i], &d[i]);
        if (ce == cudaSuccess) ce = m->s_grid_start.grab((size_t)nf * (kGridCells + 1) * 4, &ptr);
        A.cell_start = (int32_t*)ptr;
        if (ce == cudaSuccess) ce = m->s_grid_items.grab((size_t)std::max(F.nkp, 1) * 4, &ptr);
        A.cell_items = (int32_t*)ptr;
        if (ce == cudaSuccess) ce = m->s_topk.grab((size_t)std::max(nq, 1) * kTopK * 4, &ptr);
        A.topk_key = (uint32_t*)ptr;
        if (ce == cudaSuccess) ce = m->s_topk2.grab((size_t)std::max(nq, 1) * kTopK * 4, &ptr);
        A.topk_idx = (int32_t*)ptr;
        if (ce != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("scratch: ") + cudaGetErrorString(ce));
    }
    if (!rc) {
        A.F = F.v;
        if ((best_only && !skip_any) || init_mode) A.F.flags = nullptr;   // Fuse / SearchForInitialization look at every candidate
        A.n_frames = nf;
        A.generic = 1;
        A.skip_any = skip_any;
        A.th_dist = th_dist;
        A.check_orientation = check_orientation;
        A.kp_match = (int32_t*)d[0];
        A.mp_best_idx = (int32_t*)d[1];
        A.mp_best_dist = (int32_t*)d[2];
        A.q_bin = (int8_t*)d[3];
        A.nmatches = (int32_t*)d[4];
        A.blocked_words = (F.max_kp + 31) / 32 + 1;
        A.evals = m->d_evals;
        const size_t smem = ((size_t)A.blocked_words + 32) * 4;
        if (smem > 200 * 1024) rc = og_fail(ORBGPU_ERR_CAPACITY, "frame too large for the shared-memory blocked mask");
        if (!rc) {
            cudaEventRecord(m->ev0, st);
            cudaMemsetAsync(m->d_evals, 0, 8, st);
            k_grid_build<<<nf, kGridThreads, 0, st>>>(A);
            if (nq > 0) {
                static const int win_warp_env = []() { const char* e = getenv("ORBGPU_SBP_WARP"); return e ? atoi(e) : 0; }();
                const int per_cta = kWarpsPerBlock * 32 / 8;
                if (win_warp_env) k_sbp_topk<<<(nq + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(A, nq);
                else k_sbp_topk_g<8><<<(nq + per_cta - 1) / per_cta, kWarpsPerBlock * 32, 0, st>>>(A, nq);
            }
            if (best_only) {
                if (nq > 0) k_win_best<<<(nq + 255) / 256, 256, 0, st>>>(A, nq);
            } else if (init_mode) {
                // per key point of frame 2: owner (4 B) + matched distance (2 B); A.mp_best_idx doubles as the match12 output
                const size_t smem_i = (size_t)A.blocked_words * 32 * 4 + 32 * 4 + (size_t)A.blocked_words * 32 * 2;
                if (smem_i > 200 * 1024) rc = og_fail(ORBGPU_ERR_CAPACITY, "frame too large for the shared-memory match state");
                else {
                    if (smem_i > 48 * 1024) cudaFuncSetAttribute(k_init_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i);
                    k_init_select<<<nf, 32, smem_i, st>>>(A, A.mp_best_idx);
                }
            } else {
                if (smem > 48 * 1024) cudaFuncSetAttribute(k_win_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k_win_select<<<nf, 32, smem, st>>>(A);
            }
            cudaEventRecord(m->ev1, st);
            m->last_launches = 3;
            ce = cudaGetLastError();
            if (ce != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("windowed search: ") + cudaGetErrorString(ce));
        }
    }
    auto get = [&](void* host, const void* dev, size_t bytes) {
        if (!rc && host && bytes) {
            cudaError_t e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("download: ") + cudaGetErrorString(e));
        }
    };
    get(kp_match, d[0], (size_t)F.nkp * 4);
    get(init_mode ? match12 : q_best_idx, d[1], (size_t)nq * 4);
    get(q_best_dist, d[2], (size_t)nq * 4);
    get(nmatches, d[4], (size_t)nf * 4);
    cudaError_t se = cudaStreamSynchronize(st);
    if (!rc && se != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("windowed search: ") + cudaGetErrorString(se));
    free_owned(F.owned);
    free_owned(owned);
    return rc;
}

// D2H helper of the host-pointer variants
int fetch(orbgpu_matcher* m, void* host, const void* dev, size_t bytes) {
    if (!host || bytes == 0) return ORBGPU_OK;
    OGM_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, m->stream));
    return ORBGPU_OK;
}

}  // namespace

extern "C" {

int orbgpu_matcher_create(orbgpu_matcher** out, int device) {
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return og_fail(ORBGPU_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return og_fail(ORBGPU_ERR_ARG, "device index out of range");
    OGM_CUDA(cudaSetDevice(device));
    orbgpu_matcher* m = new orbgpu_matcher();
    m->device = device;
    cudaError_t ce = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreate(&m->ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&m->ev1);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_stage, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&m->d_evals, 8);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&m->d_nitems, 8);   // [0] tiled work items, [1] node pairs left to the warp kernel
    if (ce == cudaSuccess) ce = cudaMemset(m->d_evals, 0, 8);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (ce != cudaSuccess) {
        std::string msg = std::string("matcher setup failed: ") + cudaGetErrorString(ce);
        orbgpu_matcher_destroy(m);
        return og_fail(ORBGPU_ERR_CUDA, msg);
    }
    *out = m;
    return ORBGPU_OK;
}

int orbgpu_matcher_destroy(orbgpu_matcher* m) {
    if (!m) return ORBGPU_OK;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    Scratch* all[] = {&m->s_ctrl, &m->s_topk, &m->s_topk2, &m->s_bin, &m->s_items, &m->s_grid_start, &m->s_grid_items, &m->s_res1,
                      &m->s_res2, &m->s_out[0], &m->s_out[1], &m->s_out[2], &m->s_out[3], &m->s_out[4], &m->s_in[0], &m->s_in[1]};
    for (Scratch* s : all) s->release();
    for (Scratch& s : m->s_tmp) s.release();
    for (auto& b : m->arena_free) cudaFree(b.first);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->d_evals) cudaFree(m->d_evals);
    if (m->d_nitems) cudaFree(m->d_nitems);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->ev_stage) cudaEventDestroy(m->ev_stage);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return ORBGPU_OK;
}

int orbgpu_matcher_sync(orbgpu_matcher* m) {
    int rc = check_matcher(m);
    if (rc) return rc;
    OGM_CUDA(cudaStreamSynchronize(m->stream));
    return ORBGPU_OK;
}

int orbgpu_matcher_stream(orbgpu_matcher* m, void** stream_out) {
    if (!m || !stream_out) return og_fail(ORBGPU_ERR_ARG, "null argument");
    *stream_out = (void*)m->stream;
    return ORBGPU_OK;
}

int orbgpu_matcher_last_launches(const orbgpu_matcher* m) { return m ? m->last_launches : 0; }

int orbgpu_matcher_last_stats(orbgpu_matcher* m, float* kernel_ms, int64_t* distance_evals) {
    int rc = check_matcher(m);
    if (rc) return rc;
    OGM_CUDA(cudaStreamSynchronize(m->stream));
    if (kernel_ms) {
        *kernel_ms = 0.f;
        if (m->last_launches > 0) OGM_CUDA(cudaEventElapsedTime(kernel_ms, m->ev0, m->ev1));
    }
    if (distance_evals) {
        unsigned long long v = 0;
        OGM_CUDA(cudaMemcpy(&v, m->d_evals, 8, cudaMemcpyDeviceToHost));
        *distance_evals = (int64_t)v;
    }
    return ORBGPU_OK;
}

int orbgpu_matcher_configure(orbgpu_matcher* m, int min_queries, int min_candidates, int queries_per_thread) {
    if (!m) return og_fail(ORBGPU_ERR_ARG, "null matcher");
    if (queries_per_thread != 0 && queries_per_thread != 4 && queries_per_thread != 8)
        return og_fail(ORBGPU_ERR_ARG, "queries_per_thread must be 4 or 8 (0 keeps the current value)");
    m->dense_q = min_queries;
    m->dense_c = std::max(min_candidates, 1);
    if (queries_per_thread) m->tile_rq = queries_per_thread;
    return ORBGPU_OK;
}

int orbgpu_frame_set_upload(orbgpu_matcher* m, const orbgpu_frame_set* host, orbgpu_frame_set_dev** out) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    orbgpu_frame_set_dev* fs = new orbgpu_frame_set_dev();
    rc = build_frame_set(m, host, fs);
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(m->stream);
        if (e != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
    }
    if (rc) { free_owned(fs->owned); delete fs; return rc; }
    *out = fs;
    return ORBGPU_OK;
}

int orbgpu_frame_set_release(orbgpu_frame_set_dev* fs) {
    if (!fs) return ORBGPU_OK;
    cudaSetDevice(fs->device);
    free_owned(fs->owned);
    if (fs->arena_owner) {
        // searches that read the set were enqueued on the owner's stream; later users of the blocks are ordered after them there
        for (auto& b : fs->borrowed) fs->arena_owner->arena_free.push_back(b);
    } else {
        for (auto& b : fs->borrowed) cudaFree(b.first);
    }
    delete fs;
    return ORBGPU_OK;
}

static int build_mappoint_set(orbgpu_matcher* m, const orbgpu_mappoint_set* h, int n_frames, orbgpu_mappoint_set_dev* mp, orbgpu_matcher* pool) {
    if (!h || n_frames < 0 || !h->mp_off) return og_fail(ORBGPU_ERR_ARG, "null argument");
    const int n = h->mp_off[n_frames];
    if (n && (!h->proj_x || !h->proj_y || !h->view_cos || !h->level || !h->flags || !h->desc))
        return og_fail(ORBGPU_ERR_ARG, "map-point set: null array");
    mp->device = m->device; mp->n_frames = n_frames; mp->nmp = n;
    cudaStream_t st = m->stream;
    int rc = upload_array(h->mp_off, (size_t)n_frames + 1, mp->owned, &mp->v.mp_off, st, pool);
    if (!rc) rc = upload_array(h->proj_x, (size_t)n, mp->owned, &mp->v.proj_x, st, pool);
    if (!rc) rc = upload_array(h->proj_y, (size_t)n, mp->owned, &mp->v.proj_y, st, pool);
    if (!rc) rc = upload_array(h->proj_xr, (size_t)n, mp->owned, &mp->v.proj_xr, st, pool);
    if (!rc) rc = upload_array(h->view_cos, (size_t)n, mp->owned, &mp->v.view_cos, st, pool);
    if (!rc) rc = upload_array(h->level, (size_t)n, mp->owned, &mp->v.level, st, pool);
    if (!rc) rc = upload_array(h->flags, (size_t)n, mp->owned, &mp->v.flags, st, pool);
    if (!rc) rc = upload_array(h->desc, (size_t)n * 32, mp->owned, &mp->v.desc, st, pool);
    return rc;
}

int orbgpu_mappoint_set_upload(orbgpu_matcher* m, const orbgpu_mappoint_set* h, int n_frames, orbgpu_mappoint_set_dev** out) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    orbgpu_mappoint_set_dev* mp = new orbgpu_mappoint_set_dev();
    rc = build_mappoint_set(m, h, n_frames, mp, nullptr);
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(m->stream);
        if (e != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
    }
    if (rc) { free_owned(mp->owned); delete mp; return rc; }
    *out = mp;
    return ORBGPU_OK;
}

int orbgpu_mappoint_set_release(orbgpu_mappoint_set_dev* mp) {
    if (!mp) return ORBGPU_OK;
    cudaSetDevice(mp->device);
    free_owned(mp->owned);
    delete mp;
    return ORBGPU_OK;
}

int orbgpu_hamming_pairs(orbgpu_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist_out) {
    OG_NVTX("orbgpu_hamming_pairs");
    int rc = check_matcher(m);
    if (rc) return rc;
    m->last_launches = 0;
    if (n < 0 || (n && (!a || !b || !dist_out))) return og_fail(ORBGPU_ERR_ARG, "bad argument");
    if (n == 0) return ORBGPU_OK;
    void *da = nullptr, *db = nullptr, *dd = nullptr;
    OGM_CUDA(m->s_in[0].grab((size_t)n * 32, &da));
    OGM_CUDA(m->s_in[1].grab((size_t)n * 32, &db));
    OGM_CUDA(m->s_out[0].grab((size_t)n * 4, &dd));
    cudaStream_t st = m->stream;
    OGM_CUDA(cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    OGM_CUDA(cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    OGM_CUDA(cudaEventRecord(m->ev0, st));
    og::k_hamming_pairs<<<(n + 255) / 256, 256, 0, st>>>((const uint8_t*)da, (const uint8_t*)db, n, (int32_t*)dd);
    OGM_CUDA(cudaEventRecord(m->ev1, st));
    OGM_CUDA(cudaGetLastError());
    m->last_launches = 1;
    OGM_CUDA(cudaMemcpyAsync(dist_out, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    OGM_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

int orbgpu_search_by_bow_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* set1, const orbgpu_frame_set_dev* set2, int n_pairs,
                             const int32_t* idx1, const int32_t* idx2, float nnratio, int check_orientation, int th_low,
                             int th_inclusive, int require_mp2, const int64_t* match_off, int32_t* match12_dev,
                             int32_t* match_dist_dev, int32_t* nmatches_dev) {
    OG_NVTX("orbgpu_search_by_bow_dev");
    int rc = check_matcher(m);
    if (rc) return rc;
    return bow_dev(m, set1, set2, n_pairs, idx1, idx2, nnratio, check_orientation, th_low, th_inclusive, require_mp2, match_off,
                   match12_dev, match_dist_dev, nmatches_dev);
}

int orbgpu_search_for_triangulation_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* set1, const orbgpu_frame_set_dev* set2,
                                        int n_pairs, const int32_t* idx1, const int32_t* idx2, const float* f12,
                                        const float* epipole, const float* scale_factors, const float* level_sigma2,
                                        int n_levels, int only_stereo, int check_orientation, const int64_t* match_off,
                                        int32_t* match12_dev, int32_t* match_dist_dev, int32_t* nmatches_dev) {
    OG_NVTX("orbgpu_search_for_triangulation_dev");
    int rc = check_matcher(m);
    if (rc) return rc;
    return tri_dev(m, set1, set2, n_pairs, idx1, idx2, f12, epipole, scale_factors, level_sigma2, n_levels, only_stereo,
                   check_orientation, match_off, match12_dev, match_dist_dev, nmatches_dev);
}

int orbgpu_search_by_projection_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* frames, const orbgpu_mappoint_set_dev* mps,
                                    const float* scale_factors, int n_levels, float th, float nnratio, int32_t* kp_match_dev,
                                    int32_t* mp_best_idx_dev, int32_t* mp_best_dist_dev, int32_t* mp_second_dist_dev,
                                    int32_t* nmatches_dev) {
    OG_NVTX("orbgpu_search_by_projection_dev");
    int rc = check_matcher(m);
    if (rc) return rc;
    return sbp_dev(m, frames, mps, scale_factors, n_levels, th, nnratio, kp_match_dev, mp_best_idx_dev, mp_best_dist_dev,
                   mp_second_dist_dev, nmatches_dev);
}

// ---- host-pointer variants: upload, run, download ------------------------------------------------------------
static int node_scan_host(orbgpu_matcher* m, const orbgpu_frame_set* set1, const orbgpu_frame_set* set2, int n_pairs,
                          const int64_t* match_off, const int32_t* idx1, int32_t* match12, int32_t* match_dist, int32_t* nmatches,
                          int (*run)(orbgpu_matcher*, const orbgpu_frame_set_dev*, const orbgpu_frame_set_dev*, int32_t*, int32_t*,
                                     int32_t*, void*),
                          void* ctx) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (n_pairs < 0 || !set1 || !set2) return og_fail(ORBGPU_ERR_ARG, "bad argument");
    if (n_pairs == 0) return ORBGPU_OK;
    if (!match_off || !match12 || !idx1) return og_fail(ORBGPU_ERR_ARG, "null match_off/match12/idx1");
    orbgpu_frame_set_dev A, B;
    m->tmp_next = 0;
    rc = build_frame_set(m, set1, &A, m);
    const bool same = set1 == set2;
    if (!rc && !same) rc = build_frame_set(m, set2, &B, m);
    // output extent: pair p writes match12[match_off[p] .. + keypoints of frame idx1[p])
    long long out_n = 0;
    if (!rc)
        for (int p = 0; p < n_pairs; ++p) {
            if (idx1[p] < 0 || idx1[p] >= A.n_frames) { rc = og_fail(ORBGPU_ERR_ARG, "pair index out of range"); break; }
            out_n = std::max<long long>(out_n, match_off[p] + (A.h_kp_off[idx1[p] + 1] - A.h_kp_off[idx1[p]]));
        }
    void *d12 = nullptr, *dd = nullptr, *dn = nullptr;
    cudaError_t ce = cudaSuccess;
    if (!rc) {
        ce = m->s_out[0].grab((size_t)std::max<long long>(out_n, 1) * 4, &d12);
        if (ce == cudaSuccess) ce = m->s_out[1].grab((size_t)std::max<long long>(out_n, 1) * 4, &dd);
        if (ce == cudaSuccess) ce = m->s_out[2].grab((size_t)n_pairs * 4, &dn);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d12, 0xff, (size_t)std::max<long long>(out_n, 1) * 4, m->stream);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(dd, 0xff, (size_t)std::max<long long>(out_n, 1) * 4, m->stream);
        if (ce != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("scratch: ") + cudaGetErrorString(ce));
    }
    if (!rc) rc = run(m, &A, same ? &A : &B, (int32_t*)d12, match_dist ? (int32_t*)dd : nullptr, (int32_t*)dn, ctx);
    if (!rc) rc = fetch(m, match12, d12, (size_t)out_n * 4);
    if (!rc && match_dist) rc = fetch(m, match_dist, dd, (size_t)out_n * 4);
    if (!rc && nmatches) rc = fetch(m, nmatches, dn, (size_t)n_pairs * 4);
    cudaError_t se = cudaStreamSynchronize(m->stream);
    if (!rc && se != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("search: ") + cudaGetErrorString(se));
    free_owned(A.owned);
    free_owned(B.owned);
    return rc;
}

struct BowCtx { int n_pairs; const int32_t *idx1, *idx2; float nnratio; int check_ori, th_low, th_incl, req2; const int64_t* match_off; };
struct TriCtx { int n_pairs; const int32_t *idx1, *idx2; const float *f12, *epipole, *scale, *sigma2; int n_levels, only_stereo, check_ori; const int64_t* match_off; };

int orbgpu_search_by_bow(orbgpu_matcher* m, const orbgpu_frame_set* set1, const orbgpu_frame_set* set2, int n_pairs,
                         const int32_t* idx1, const int32_t* idx2, float nnratio, int check_orientation, int th_low,
                         int th_inclusive, int require_mp2, const int64_t* match_off, int32_t* match12, int32_t* match_dist,
                         int32_t* nmatches) {
    OG_NVTX("orbgpu_search_by_bow");
    BowCtx c{n_pairs, idx1, idx2, nnratio, check_orientation, th_low, th_inclusive, require_mp2, match_off};
    return node_scan_host(m, set1, set2, n_pairs, match_off, idx1, match12, match_dist, nmatches,
                          [](orbgpu_matcher* mm, const orbgpu_frame_set_dev* a, const orbgpu_frame_set_dev* b, int32_t* d12, int32_t* dd,
                             int32_t* dn, void* vc) {
                              BowCtx* c = (BowCtx*)vc;
                              return bow_dev(mm, a, b, c->n_pairs, c->idx1, c->idx2, c->nnratio, c->check_ori, c->th_low, c->th_incl,
                                             c->req2, c->match_off, d12, dd, dn);
                          },
                          &c);
}

int orbgpu_search_for_triangulation(orbgpu_matcher* m, const orbgpu_frame_set* set1, const orbgpu_frame_set* set2, int n_pairs,
                                    const int32_t* idx1, const int32_t* idx2, const float* f12, const float* epipole,
                                    const float* scale_factors, const float* level_sigma2, int n_levels, int only_stereo,
                                    int check_orientation, const int64_t* match_off, int32_t* match12, int32_t* match_dist,
                                    int32_t* nmatches) {
    OG_NVTX("orbgpu_search_for_triangulation");
    TriCtx c{n_pairs, idx1, idx2, f12, epipole, scale_factors, level_sigma2, n_levels, only_stereo, check_orientation, match_off};
    return node_scan_host(m, set1, set2, n_pairs, match_off, idx1, match12, match_dist, nmatches,
                          [](orbgpu_matcher* mm, const orbgpu_frame_set_dev* a, const orbgpu_frame_set_dev* b, int32_t* d12, int32_t* dd,
                             int32_t* dn, void* vc) {
                              TriCtx* c = (TriCtx*)vc;
                              return tri_dev(mm, a, b, c->n_pairs, c->idx1, c->idx2, c->f12, c->epipole, c->scale, c->sigma2, c->n_levels,
                                             c->only_stereo, c->check_ori, c->match_off, d12, dd, dn);
                          },
                          &c);
}

int orbgpu_search_by_projection(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_mappoint_set* mps,
                                const float* scale_factors, int n_levels, float th, float nnratio, int32_t* kp_match,
                                int32_t* mp_best_idx, int32_t* mp_best_dist, int32_t* mp_second_dist, int32_t* nmatches) {
    OG_NVTX("orbgpu_search_by_projection");
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!frames || !mps) return og_fail(ORBGPU_ERR_ARG, "null argument");
    orbgpu_frame_set_dev F;
    orbgpu_mappoint_set_dev Mset;
    orbgpu_mappoint_set_dev* M = &Mset;
    m->tmp_next = 0;
    rc = build_frame_set(m, frames, &F, m);
    if (!rc) rc = build_mappoint_set(m, mps, frames->n_frames, M, m);
    void* d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!rc) {
        const size_t sz[5] = {(size_t)std::max(F.nkp, 1) * 4, (size_t)std::max(M->nmp, 1) * 4, (size_t)std::max(M->nmp, 1) * 4,
                              (size_t)std::max(M->nmp, 1) * 4, (size_t)std::max(F.n_frames, 1) * 4};
        for (int i = 0; i < 5 && !rc; ++i) {
            cudaError_t ce = m->s_out[i].grab(sz[i], &d[i]);
            if (ce != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("scratch: ") + cudaGetErrorString(ce));
        }
    }
    if (!rc)
        rc = sbp_dev(m, &F, M, scale_factors, n_levels, th, nnratio, kp_match ? (int32_t*)d[0] : nullptr, mp_best_idx ? (int32_t*)d[1] : nullptr,
                     mp_best_dist ? (int32_t*)d[2] : nullptr, mp_second_dist ? (int32_t*)d[3] : nullptr, (int32_t*)d[4]);
    if (!rc && kp_match) rc = fetch(m, kp_match, d[0], (size_t)F.nkp * 4);
    if (!rc && mp_best_idx) rc = fetch(m, mp_best_idx, d[1], (size_t)M->nmp * 4);
    if (!rc && mp_best_dist) rc = fetch(m, mp_best_dist, d[2], (size_t)M->nmp * 4);
    if (!rc && mp_second_dist) rc = fetch(m, mp_second_dist, d[3], (size_t)M->nmp * 4);
    if (!rc && nmatches) rc = fetch(m, nmatches, d[4], (size_t)F.n_frames * 4);
    cudaError_t se = cudaStreamSynchronize(m->stream);
    if (!rc && se != cudaSuccess) rc = og_fail(ORBGPU_ERR_CUDA, std::string("search: ") + cudaGetErrorString(se));
    free_owned(F.owned);
    free_owned(Mset.owned);
    return rc;
}

int orbgpu_search_windowed(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_window_query_set* queries, int th_dist,
                           int skip_any_mappoint, int check_orientation, int32_t* kp_match, int32_t* q_best_idx, int32_t* q_best_dist,
                           int32_t* nmatches) {
    OG_NVTX("orbgpu_search_windowed");
    int rc = check_matcher(m);
    if (rc) return rc;
    return win_host(m, frames, queries, th_dist, skip_any_mappoint, check_orientation, kp_match, q_best_idx, q_best_dist, nmatches);
}

}  // extern "C"

// ---- frame set straight from an extractor's last batch (+ vocabulary transform), nothing leaves the device -------------
int og_extractor_last_results(orbgpu_extractor* ex, int* device, int* batch, int* stride, const og::KeyPoint** kp, const uint8_t** desc,
                              const int32_t** counts, cudaStream_t* stream);   // og_capi.cu

namespace og {
// one block per frame: strided extractor output -> the concatenated layout of a frame set
__global__ void __launch_bounds__(256) k_compact_extraction(const KeyPoint* __restrict__ kp, const uint8_t* __restrict__ desc,
                                                            const float* __restrict__ u_right, int stride, int ur_stride,
                                                            const int32_t* __restrict__ kp_off, KeyPoint* __restrict__ kp_out,
                                                            uint8_t* __restrict__ desc_out, float* __restrict__ ur_out) {
    const int f = blockIdx.x, base = kp_off[f], n = kp_off[f + 1] - base;
    const uint32_t* ks = reinterpret_cast<const uint32_t*>(kp + (size_t)f * stride);
    uint32_t* kd = reinterpret_cast<uint32_t*>(kp_out + base);
    for (int i = threadIdx.x; i < n * 7; i += 256) kd[i] = ks[i];
    const uint4* ds = reinterpret_cast<const uint4*>(desc + (size_t)f * stride * 32);
    uint4* dd = reinterpret_cast<uint4*>(desc_out + (size_t)base * 32);
    for (int i = threadIdx.x; i < n * 2; i += 256) dd[i] = ds[i];
    if (u_right)
        for (int i = threadIdx.x; i < n; i += 256) ur_out[base + i] = u_right[(size_t)f * ur_stride + i];
}
__global__ void k_gather_i32(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int n, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
}  // namespace og

extern "C" int orbgpu_frame_set_from_extraction(orbgpu_matcher* m, orbgpu_extractor* ex, orbgpu_vocabulary* voc, int levelsup, int kp_flag,
                                                const float* u_right_dev, int u_right_stride, const float* grid,
                                                orbgpu_frame_set_dev** out) {
    OG_NVTX("orbgpu_frame_set_from_extraction");
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    int dev = 0, batch = 0, stride = 0;
    const og::KeyPoint* kp = nullptr;
    const uint8_t* desc = nullptr;
    const int32_t* counts = nullptr;
    cudaStream_t ex_stream = nullptr;
    if ((rc = og_extractor_last_results(ex, &dev, &batch, &stride, &kp, &desc, &counts, &ex_stream))) return rc;
    if (dev != m->device) return og_fail(ORBGPU_ERR_ARG, "frame_set_from_extraction: extractor and matcher live on different devices");
    if (u_right_dev && u_right_stride < stride) return og_fail(ORBGPU_ERR_ARG, "frame_set_from_extraction: u_right stride too small");
    cudaStream_t st = m->stream;
    // the per-frame key-point counts come to the host: the searches plan their launches from them
    std::vector<int32_t> cnt(batch);
    OGM_CUDA(cudaMemcpyAsync(cnt.data(), counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, ex_stream));
    OGM_CUDA(cudaStreamSynchronize(ex_stream));
    orbgpu_frame_set_dev* fs = new orbgpu_frame_set_dev();
    fs->device = m->device;
    fs->n_frames = batch;
    fs->h_kp_off.resize(batch + 1);
    fs->h_kp_off[0] = 0;
    for (int f = 0; f < batch; ++f) {
        fs->h_kp_off[f + 1] = fs->h_kp_off[f] + cnt[f];
        fs->max_kp = std::max(fs->max_kp, cnt[f]);
    }
    fs->nkp = fs->h_kp_off[batch];
    fs->arena_owner = m;
    auto bail = [&](int code) { orbgpu_frame_set_release(fs); return code; };
    auto alloc = [&](size_t bytes, void** p) -> cudaError_t {
        size_t cap = 0;
        cudaError_t e = m->arena_take(bytes, p, &cap);
        if (e == cudaSuccess) fs->borrowed.push_back(std::make_pair(*p, cap));
        return e;
    };
#define OGF_CUDA(expr)                                                                                             \
    do {                                                                                                           \
        cudaError_t e_ = (expr);                                                                                   \
        if (e_ != cudaSuccess) return bail(og_fail(ORBGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_))); \
    } while (0)
    const size_t nkp = (size_t)fs->nkp;
    void *d_off, *d_keys, *d_desc, *d_flags, *d_ur = nullptr, *d_grid = nullptr;
    OGF_CUDA(alloc((size_t)(batch + 1) * 4, &d_off));
    OGF_CUDA(alloc(nkp * sizeof(og::KeyPoint), &d_keys));
    OGF_CUDA(alloc(nkp * 32, &d_desc));
    OGF_CUDA(alloc(nkp, &d_flags));
    if (u_right_dev) OGF_CUDA(alloc(nkp * 4, &d_ur));
    OGF_CUDA(cudaMemcpyAsync(d_off, fs->h_kp_off.data(), (size_t)(batch + 1) * 4, cudaMemcpyHostToDevice, st));
    OGF_CUDA(cudaMemsetAsync(d_flags, kp_flag & 0xff, std::max<size_t>(nkp, 1), st));
    og::k_compact_extraction<<<batch, 256, 0, st>>>(kp, desc, u_right_dev, stride, u_right_stride, (const int32_t*)d_off, (og::KeyPoint*)d_keys,
                                                   (uint8_t*)d_desc, (float*)d_ur);
    fs->v.kp_off = (const int32_t*)d_off;
    fs->v.keys = (const og::KeyPoint*)d_keys;
    fs->v.desc = (const uint8_t*)d_desc;
    fs->v.flags = (const uint8_t*)d_flags;
    fs->v.u_right = (const float*)d_ur;
    if (grid) {
        std::vector<float> g((size_t)batch * 4);
        for (int f = 0; f < batch; ++f) std::copy(grid, grid + 4, g.begin() + (size_t)f * 4);
        OGF_CUDA(alloc(g.size() * 4, &d_grid));
        OGF_CUDA(cudaMemcpyAsync(d_grid, g.data(), g.size() * 4, cudaMemcpyHostToDevice, st));
        OGF_CUDA(cudaStreamSynchronize(st));   // g is a local
        fs->v.grid = (const float*)d_grid;
        fs->has_grid = true;
    }
    if (voc) {
        void *d_node_off, *d_node_id, *d_feat_off, *d_feat, *d_ent;
        OGF_CUDA(alloc((size_t)(batch + 1) * 4, &d_node_off));
        OGF_CUDA(alloc(nkp * 4, &d_node_id));
        OGF_CUDA(alloc((nkp + 1) * 4, &d_feat_off));
        OGF_CUDA(alloc(nkp * 4, &d_feat));
        OGF_CUDA(alloc((size_t)(batch + 1) * 4, &d_ent));
        void* vs = nullptr;
        if (orbgpu_vocabulary_stream(voc, &vs) != ORBGPU_OK) return bail(ORBGPU_ERR_ARG);
        cudaEvent_t ev;
        OGF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t ce = cudaEventRecord(ev, st);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent((cudaStream_t)vs, ev, 0);
        int vrc = ORBGPU_OK;
        if (ce == cudaSuccess)
            vrc = orbgpu_bow_transform_dev(voc, batch, (const int32_t*)d_off, fs->nkp, fs->max_kp, (const uint8_t*)d_desc, levelsup, nullptr, nullptr,
                                           nullptr, (int32_t*)d_node_off, (int32_t*)d_node_id, (int32_t*)d_feat_off, (int32_t*)d_feat, nullptr, nullptr);
        if (ce == cudaSuccess && vrc == ORBGPU_OK) ce = cudaEventRecord(ev, (cudaStream_t)vs);
        if (ce == cudaSuccess && vrc == ORBGPU_OK) ce = cudaStreamWaitEvent(st, ev, 0);
        cudaEventDestroy(ev);
        if (vrc != ORBGPU_OK) return bail(vrc);
        OGF_CUDA(ce);
        // per-frame node / entry counts back to the host (the BoW searches size their work lists from them)
        og::k_gather_i32<<<(batch + 1 + 255) / 256, 256, 0, st>>>((const int32_t*)d_feat_off, (const int32_t*)d_node_off, batch + 1, (int32_t*)d_ent);
        std::vector<int32_t> node_off(batch + 1), ent_off(batch + 1);
        OGF_CUDA(cudaMemcpyAsync(node_off.data(), d_node_off, (size_t)(batch + 1) * 4, cudaMemcpyDeviceToHost, st));
        OGF_CUDA(cudaMemcpyAsync(ent_off.data(), d_ent, (size_t)(batch + 1) * 4, cudaMemcpyDeviceToHost, st));
        OGF_CUDA(cudaStreamSynchronize(st));
        fs->has_fv = true;
        fs->nnodes = node_off[batch];
        fs->nfeat = ent_off[batch];
        fs->h_node_cnt.resize(batch);
        fs->h_ent_cnt.resize(batch);
        for (int f = 0; f < batch; ++f) {
            fs->h_node_cnt[f] = node_off[f + 1] - node_off[f];
            fs->h_ent_cnt[f] = ent_off[f + 1] - ent_off[f];
            fs->max_nodes = std::max(fs->max_nodes, fs->h_node_cnt[f]);
        }
        fs->v.node_off = (const int32_t*)d_node_off;
        fs->v.node_id = (const int32_t*)d_node_id;
        fs->v.feat_off = (const int32_t*)d_feat_off;
        fs->v.feat = (const int32_t*)d_feat;
    } else {
        OGF_CUDA(cudaStreamSynchronize(st));
    }
    OGF_CUDA(cudaGetLastError());
#undef OGF_CUDA
    *out = fs;
    return ORBGPU_OK;
}

extern "C" int orbgpu_frame_set_dev_info(const orbgpu_frame_set_dev* fs, int* n_frames, int32_t* kp_off, int* n_nodes, int* n_feat) {
    if (!fs) return og_fail(ORBGPU_ERR_ARG, "null frame set");
    if (n_frames) *n_frames = fs->n_frames;
    if (kp_off) std::copy(fs->h_kp_off.begin(), fs->h_kp_off.end(), kp_off);
    if (n_nodes) *n_nodes = fs->nnodes;
    if (n_feat) *n_feat = fs->nfeat;
    return ORBGPU_OK;
}

/* Copies the arrays of a device-resident frame set back to the host (any pointer may be NULL); sizes as reported by
 * orbgpu_frame_set_dev_info. */
extern "C" int orbgpu_frame_set_download(orbgpu_matcher* m, const orbgpu_frame_set_dev* fs, orbgpu_keypoint* keys, uint8_t* desc, int32_t* fv_node_off,
                                         int32_t* fv_node_id, int32_t* fv_feat_off, int32_t* fv_feat) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!fs) return og_fail(ORBGPU_ERR_ARG, "null frame set");
    cudaStream_t st = m->stream;
    if (keys && fs->nkp) OGM_CUDA(cudaMemcpyAsync(keys, fs->v.keys, (size_t)fs->nkp * sizeof(og::KeyPoint), cudaMemcpyDeviceToHost, st));
    if (desc && fs->nkp) OGM_CUDA(cudaMemcpyAsync(desc, fs->v.desc, (size_t)fs->nkp * 32, cudaMemcpyDeviceToHost, st));
    if (fs->has_fv) {
        if (fv_node_off) OGM_CUDA(cudaMemcpyAsync(fv_node_off, fs->v.node_off, (size_t)(fs->n_frames + 1) * 4, cudaMemcpyDeviceToHost, st));
        if (fv_node_id && fs->nnodes) OGM_CUDA(cudaMemcpyAsync(fv_node_id, fs->v.node_id, (size_t)fs->nnodes * 4, cudaMemcpyDeviceToHost, st));
        if (fv_feat_off) OGM_CUDA(cudaMemcpyAsync(fv_feat_off, fs->v.feat_off, (size_t)(fs->nnodes + 1) * 4, cudaMemcpyDeviceToHost, st));
        if (fv_feat && fs->nfeat) OGM_CUDA(cudaMemcpyAsync(fv_feat, fs->v.feat, (size_t)fs->nfeat * 4, cudaMemcpyDeviceToHost, st));
    }
    OGM_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

// ---- Frame::isInFrustum + MapPoint::PredictScale (Frame.cc:274-342, MapPoint.cc:421-436) ----------------------------------
namespace og {
struct FrustumArgs {
    const float* cam;        // [n_frames][24]
    const int32_t* mp_off;   // device copy
    const float *world_pos, *normal, *min_d, *max_d, *max_distance;
    const uint8_t* flags_in;   // optional: bits 1.. are kept, bit 0 becomes the visibility (orbgpu_mappoint_set flags)
    uint8_t* in_view;
    float *proj_x, *proj_y, *proj_xr, *view_cos;
    int32_t* level;
    float log_sf, cos_limit;
    int n_levels;
};
// grid (chunks of 256 points, frames).  Plain IEEE single / double operations in the reference's order (the library is built
// with --fmad=false); cv::gemm's 3x3 * 3x1 + 3x1: float dot product left to right, addend joined in double.
__global__ void __launch_bounds__(256) k_frustum(FrustumArgs A) {
    const int f = blockIdx.y;
    const int q0 = A.mp_off[f], n = A.mp_off[f + 1] - q0;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int q = q0 + i;
    const float* c = A.cam + 24 * f;
    const float Px = A.world_pos[3 * q], Py = A.world_pos[3 * q + 1], Pz = A.world_pos[3 * q + 2];
    uint8_t ok = 0;
    float u = 0.f, v = 0.f, ur = 0.f, vc = 0.f;
    int lvl = 0;
    do {
        float Pc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float t = __ldg(c + 3 * r) * Px + __ldg(c + 3 * r + 1) * Py + __ldg(c + 3 * r + 2) * Pz;
            Pc[r] = (float)((double)t + (double)__ldg(c + 9 + r));
        }
        if (Pc[2] < 0.0f) break;
        const float invz = 1.0f / Pc[2];
        const float uu = __ldg(c + 15) * Pc[0] * invz + __ldg(c + 17);
        const float vv = __ldg(c + 16) * Pc[1] * invz + __ldg(c + 18);
        if (uu < __ldg(c + 20) || uu > __ldg(c + 21)) break;
        if (vv < __ldg(c + 22) || vv > __ldg(c + 23)) break;
        const float PO0 = Px - __ldg(c + 12), PO1 = Py - __ldg(c + 13), PO2 = Pz - __ldg(c + 14);
        const float dist = (float)sqrt((double)PO0 * (double)PO0 + (double)PO1 * (double)PO1 + (double)PO2 * (double)PO2);
        if (dist < A.min_d[q] || dist > A.max_d[q]) break;
        const double dot = (double)PO0 * (double)A.normal[3 * q] + (double)PO1 * (double)A.normal[3 * q + 1] + (double)PO2 * (double)A.normal[3 * q + 2];
        const float viewCos = (float)(dot / (double)dist);
        if (viewCos < A.cos_limit) break;
        const float ratio = A.max_distance[q] / dist;
        const float qf = ceilf((float)log((double)ratio) / A.log_sf);
        lvl = qf >= (float)A.n_levels ? A.n_levels - 1 : (qf < 0.f ? 0 : (int)qf);
        if (!(qf == qf)) lvl = 0;   // NaN: the reference's float -> int conversion is undefined here
        ok = 1; u = uu; v = vv; ur = uu - __ldg(c + 19) * invz; vc = viewCos;
    } while (0);
    A.in_view[q] = A.flags_in ? (uint8_t)((A.flags_in[q] & 0xFEu) | ok) : ok;
    A.proj_x[q] = u; A.proj_y[q] = v; A.proj_xr[q] = ur; A.view_cos[q] = vc;
    A.level[q] = lvl;
}
}  // namespace og

extern "C" int orbgpu_is_in_frustum_dev(orbgpu_matcher* m, int n_frames, const float* cam, float log_scale_factor, int n_levels,
                                        float viewing_cos_limit, const int32_t* mp_off, const float* world_pos, const float* normal,
                                        const float* min_d, const float* max_d, const float* max_distance, uint8_t* in_view, float* proj_x,
                                        float* proj_y, float* proj_xr, int32_t* level, float* view_cos) {
    OG_NVTX("orbgpu_is_in_frustum_dev");
    int rc = check_matcher(m);
    if (rc) return rc;
    m->last_launches = 0;
    if (n_frames < 0 || (n_frames && (!cam || !mp_off))) return og_fail(ORBGPU_ERR_ARG, "is_in_frustum: null cam / mp_off");
    if (n_frames == 0) return ORBGPU_OK;
    int max_n = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (mp_off[f + 1] < mp_off[f]) return og_fail(ORBGPU_ERR_ARG, "is_in_frustum: mp_off must be non-decreasing");
        max_n = std::max(max_n, mp_off[f + 1] - mp_off[f]);
    }
    if (max_n == 0) return ORBGPU_OK;
    if (!world_pos || !normal || !min_d || !max_d || !max_distance || !in_view || !proj_x || !proj_y || !proj_xr || !level || !view_cos)
        return og_fail(ORBGPU_ERR_ARG, "is_in_frustum: null array");
    if (!(log_scale_factor > 0.f) || n_levels < 1) return og_fail(ORBGPU_ERR_ARG, "is_in_frustum: bad scale factor / level count");
    void *d_cam, *d_off;
    OGM_CUDA(m->s_grid_start.grab((size_t)n_frames * 24 * 4, &d_cam));
    OGM_CUDA(m->s_grid_items.grab((size_t)(n_frames + 1) * 4, &d_off));
    OGM_CUDA(cudaMemcpyAsync(d_cam, cam, (size_t)n_frames * 24 * 4, cudaMemcpyHostToDevice, m->stream));
    OGM_CUDA(cudaMemcpyAsync(d_off, mp_off, (size_t)(n_frames + 1) * 4, cudaMemcpyHostToDevice, m->stream));
    og::FrustumArgs A = {(const float*)d_cam, (const int32_t*)d_off, world_pos, normal, min_d, max_d, max_distance, nullptr, in_view, proj_x, proj_y,
                         proj_xr, view_cos, level, log_scale_factor, viewing_cos_limit, n_levels};
    og::k_frustum<<<dim3((max_n + 255) / 256, n_frames), 256, 0, m->stream>>>(A);
    m->last_launches = 1;
    OGM_CUDA(cudaGetLastError());
    return ORBGPU_OK;
}

extern "C" int orbgpu_is_in_frustum(orbgpu_matcher* m, int n_frames, const float* cam, float log_scale_factor, int n_levels,
                                    float viewing_cos_limit, const int32_t* mp_off, const float* world_pos, const float* normal,
                                    const float* min_d, const float* max_d, const float* max_distance, uint8_t* in_view, float* proj_x,
                                    float* proj_y, float* proj_xr, int32_t* level, float* view_cos) {
    OG_NVTX("orbgpu_is_in_frustum");
    int rc = check_matcher(m);
    if (rc) return rc;
    if (n_frames < 0 || (n_frames && !mp_off)) return og_fail(ORBGPU_ERR_ARG, "is_in_frustum: null mp_off");
    const size_t n = n_frames ? (size_t)mp_off[n_frames] : 0;
    if (n == 0) return ORBGPU_OK;
    if (!world_pos || !normal || !min_d || !max_d || !max_distance || !in_view || !proj_x || !proj_y || !proj_xr || !level || !view_cos)
        return og_fail(ORBGPU_ERR_ARG, "is_in_frustum: null array");
    cudaStream_t st = m->stream;
    m->tmp_next = 0;
    const void* src[5] = {world_pos, normal, min_d, max_d, max_distance};
    const size_t sb[5] = {n * 12, n * 12, n * 4, n * 4, n * 4};
    void* din[5];
    for (int i = 0; i < 5; ++i) {
        OGM_CUDA(m->s_tmp[m->tmp_next++].grab(sb[i], &din[i]));
        OGM_CUDA(cudaMemcpyAsync(din[i], src[i], sb[i], cudaMemcpyHostToDevice, st));
    }
    void* dout[6];
    const size_t ob[6] = {n, n * 4, n * 4, n * 4, n * 4, n * 4};
    for (int i = 0; i < 6; ++i) OGM_CUDA(m->s_tmp[m->tmp_next++].grab(ob[i], &dout[i]));
    rc = orbgpu_is_in_frustum_dev(m, n_frames, cam, log_scale_factor, n_levels, viewing_cos_limit, mp_off, (const float*)din[0], (const float*)din[1],
                                  (const float*)din[2], (const float*)din[3], (const float*)din[4], (uint8_t*)dout[0], (float*)dout[1], (float*)dout[2],
                                  (float*)dout[3], (int32_t*)dout[4], (float*)dout[5]);
    if (rc) return rc;
    void* host[6] = {in_view, proj_x, proj_y, proj_xr, level, view_cos};
    for (int i = 0; i < 6; ++i) OGM_CUDA(cudaMemcpyAsync(host[i], dout[i], ob[i], cudaMemcpyDeviceToHost, st));
    OGM_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-316) -------------------------------------------------------
namespace og {
// One block per map point.  Shared: N descriptors (uint4 pairs) | N x N distance matrix (u16).
__global__ void __launch_bounds__(128) k_distinctive(const int32_t* __restrict__ obs_off, const uint8_t* __restrict__ desc, int n_max,
                                                     int32_t* __restrict__ best_idx, int32_t* __restrict__ best_median) {
    extern __shared__ __align__(16) unsigned char dd_smem[];
    uint4* sd = reinterpret_cast<uint4*>(dd_smem);
    uint16_t* dm = reinterpret_cast<uint16_t*>(dd_smem + (size_t)n_max * 32);
    __shared__ uint32_t wbest[4];
    const int p = blockIdx.x, t = threadIdx.x;
    const int o = obs_off[p], N = obs_off[p + 1] - o;
    if (N <= 0) {
        if (t == 0) { best_idx[p] = -1; if (best_median) best_median[p] = 0x7fffffff; }
        return;
    }
    const uint4* g = reinterpret_cast<const uint4*>(desc + (size_t)o * 32);
    for (int i = t; i < 2 * N; i += 128) sd[i] = __ldg(g + i);
    __syncthreads();
    for (int e = t; e < N * N; e += 128) {
        const int i = e / N, j = e - i * N;
        Desc a, b;
        a.lo = sd[2 * i]; a.hi = sd[2 * i + 1];
        b.lo = sd[2 * j]; b.hi = sd[2 * j + 1];
        dm[e] = (uint16_t)hamming256(a, b);
    }
    __syncthreads();
    const int k = (int)(0.5 * (N - 1));   // vDists[0.5*(N-1)] (:301)
    uint32_t best = 0xffffffffu;
    for (int r = t; r < N; r += 128) {
        const uint16_t* row = dm + (size_t)r * N;
        int lo = 0, hi = 256;             // smallest v with #{d <= v} >= k + 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j = 0; j < N; ++j) c += row[j] <= mid;
            if (c >= k + 1) hi = mid; else lo = mid + 1;
        }
        best = min(best, ((uint32_t)lo << 16) | (uint32_t)r);
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((t & 31) == 0) wbest[t >> 5] = best;
    __syncthreads();
    if (t == 0) {
        best = min(min(wbest[0], wbest[1]), min(wbest[2], wbest[3]));
        best_idx[p] = (int32_t)(best & 0xffffu);
        if (best_median) best_median[p] = (int32_t)(best >> 16);
    }
}
}  // namespace og

extern "C" int orbgpu_distinctive_descriptors(orbgpu_matcher* m, int n_points, const int32_t* obs_off, const uint8_t* desc, int32_t* best_idx,
                                              int32_t* best_median) {
    OG_NVTX("orbgpu_distinctive_descriptors");
    int rc = check_matcher(m);
    if (rc) return rc;
    m->last_launches = 0;
    if (n_points < 0 || (n_points && (!obs_off || !best_idx))) return og_fail(ORBGPU_ERR_ARG, "distinctive_descriptors: null argument");
    if (n_points == 0) return ORBGPU_OK;
    int n_max = 0;
    for (int p = 0; p < n_points; ++p) {
        if (obs_off[p + 1] < obs_off[p]) return og_fail(ORBGPU_ERR_ARG, "distinctive_descriptors: obs_off must be non-decreasing");
        n_max = std::max(n_max, obs_off[p + 1] - obs_off[p]);
    }
    if (n_max > 256) return og_fail(ORBGPU_ERR_CAPACITY, "distinctive_descriptors: more than 256 observations of one map point");
    const size_t total = (size_t)obs_off[n_points];
    if (total && !desc) return og_fail(ORBGPU_ERR_ARG, "distinctive_descriptors: null descriptors");
    cudaStream_t st = m->stream;
    m->tmp_next = 0;
    void *d_off, *d_desc, *d_idx, *d_med;
    OGM_CUDA(m->s_tmp[m->tmp_next++].grab((size_t)(n_points + 1) * 4, &d_off));
    OGM_CUDA(m->s_tmp[m->tmp_next++].grab(std::max<size_t>(total, 1) * 32, &d_desc));
    OGM_CUDA(m->s_tmp[m->tmp_next++].grab((size_t)n_points * 4, &d_idx));
    OGM_CUDA(m->s_tmp[m->tmp_next++].grab((size_t)n_points * 4, &d_med));
    OGM_CUDA(cudaMemcpyAsync(d_off, obs_off, (size_t)(n_points + 1) * 4, cudaMemcpyHostToDevice, st));
    if (total) OGM_CUDA(cudaMemcpyAsync(d_desc, desc, total * 32, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)n_max * 32 + (size_t)n_max * n_max * 2;
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        OGM_CUDA(cudaFuncSetAttribute(og::k_distinctive, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 + 256 * 256 * 2));
        attr_set = true;
    }
    og::k_distinctive<<<n_points, 128, smem, st>>>((const int32_t*)d_off, (const uint8_t*)d_desc, n_max, (int32_t*)d_idx, (int32_t*)d_med);
    m->last_launches = 1;
    OGM_CUDA(cudaGetLastError());
    OGM_CUDA(cudaMemcpyAsync(best_idx, d_idx, (size_t)n_points * 4, cudaMemcpyDeviceToHost, st));
    if (best_median) OGM_CUDA(cudaMemcpyAsync(best_median, d_med, (size_t)n_points * 4, cudaMemcpyDeviceToHost, st));
    OGM_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

// Best-only windowed search (the search loops of Fuse x2 and SearchBySim3): see include/orbgpu.h
extern "C" int orbgpu_search_window_best(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_window_query_set* queries,
                                         const float* inv_level_sigma2, int n_levels, int skip_flagged, int32_t* q_best_idx,
                                         int32_t* q_best_dist) {
    OG_NVTX("orbgpu_search_window_best");
    int rc = check_matcher(m);
    if (rc) return rc;
    if (inv_level_sigma2 && n_levels < 1) return og_fail(ORBGPU_ERR_ARG, "search_window_best: n_levels");
    if (inv_level_sigma2 && queries && !queries->ur) return og_fail(ORBGPU_ERR_ARG, "search_window_best: the chi-square gate needs the queries' ur");
    return win_host(m, frames, queries, 256, skip_flagged, 0, nullptr, q_best_idx, q_best_dist, nullptr, true, inv_level_sigma2, n_levels);
}

// ORBmatcher::SearchForInitialization (ORBmatcher.cc:493-632): see include/orbgpu.h
extern "C" int orbgpu_search_for_initialization(orbgpu_matcher* m, const orbgpu_frame_set* frames2, const orbgpu_window_query_set* queries1,
                                                float nnratio, int check_orientation, int32_t* match12, int32_t* nmatches) {
    OG_NVTX("orbgpu_search_for_initialization");
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!match12) return og_fail(ORBGPU_ERR_ARG, "search_for_initialization: null match12");
    return win_host(m, frames2, queries1, ORBGPU_TH_LOW, 0, check_orientation, nullptr, nullptr, nullptr, nmatches, false, nullptr, 0, true, nnratio,
                    match12);
}

// Tracking::SearchLocalPoints' projection step (Tracking.cc:1150-1200) straight into a device-resident map-point set
extern "C" int orbgpu_mappoint_set_project(orbgpu_matcher* m, int n_frames, const float* cam, float log_scale_factor, int n_levels,
                                           float viewing_cos_limit, const int32_t* mp_off, const float* world_pos, const float* normal,
                                           const float* min_d, const float* max_d, const float* max_distance, const uint8_t* flags,
                                           const uint8_t* desc, orbgpu_mappoint_set_dev** out) {
    OG_NVTX("orbgpu_mappoint_set_project");
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    if (n_frames < 1 || !cam || !mp_off) return og_fail(ORBGPU_ERR_ARG, "mappoint_set_project: null cam / mp_off");
    const size_t n = (size_t)mp_off[n_frames];
    int max_n = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (mp_off[f + 1] < mp_off[f]) return og_fail(ORBGPU_ERR_ARG, "mappoint_set_project: mp_off must be non-decreasing");
        max_n = std::max(max_n, mp_off[f + 1] - mp_off[f]);
    }
    if (n && (!world_pos || !normal || !min_d || !max_d || !max_distance || !flags || !desc))
        return og_fail(ORBGPU_ERR_ARG, "mappoint_set_project: null array");
    if (!(log_scale_factor > 0.f) || n_levels < 1) return og_fail(ORBGPU_ERR_ARG, "mappoint_set_project: bad scale factor / level count");
    cudaStream_t st = m->stream;
    orbgpu_mappoint_set_dev* mp = new orbgpu_mappoint_set_dev();
    mp->device = m->device; mp->n_frames = n_frames; mp->nmp = (int)n;
    auto bail = [&](int code) { free_owned(mp->owned); delete mp; return code; };
    rc = upload_array(mp_off, (size_t)n_frames + 1, mp->owned, &mp->v.mp_off, st);
    if (!rc) rc = upload_array(desc, n * 32, mp->owned, &mp->v.desc, st);
    if (rc) return bail(rc);
    void* o[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const size_t ob[6] = {n * 4, n * 4, n * 4, n * 4, n * 4, n};
    for (int i = 0; i < 6; ++i) {
        cudaError_t e = cudaMalloc(&o[i], std::max<size_t>(ob[i], 256));
        if (e != cudaSuccess) return bail(og_fail(ORBGPU_ERR_CUDA, std::string("mappoint_set_project: ") + cudaGetErrorString(e)));
        mp->owned.push_back(o[i]);
    }
    mp->v.proj_x = (const float*)o[0]; mp->v.proj_y = (const float*)o[1]; mp->v.proj_xr = (const float*)o[2];
    mp->v.view_cos = (const float*)o[3]; mp->v.level = (const int32_t*)o[4]; mp->v.flags = (const uint8_t*)o[5];
    if (n) {
        m->tmp_next = 0;
        const void* src[6] = {world_pos, normal, min_d, max_d, max_distance, flags};
        const size_t sb[6] = {n * 12, n * 12, n * 4, n * 4, n * 4, n};
        void* din[6];
        for (int i = 0; i < 6; ++i) {
            cudaError_t e = m->s_tmp[m->tmp_next++].grab(sb[i], &din[i]);
            if (e == cudaSuccess) e = cudaMemcpyAsync(din[i], src[i], sb[i], cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return bail(og_fail(ORBGPU_ERR_CUDA, std::string("mappoint_set_project: ") + cudaGetErrorString(e)));
        }
        void *d_cam, *d_off;
        cudaError_t e = m->s_grid_start.grab((size_t)n_frames * 24 * 4, &d_cam);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cam, cam, (size_t)n_frames * 24 * 4, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return bail(og_fail(ORBGPU_ERR_CUDA, std::string("mappoint_set_project: ") + cudaGetErrorString(e)));
        d_off = (void*)mp->v.mp_off;
        og::FrustumArgs A = {(const float*)d_cam, (const int32_t*)d_off, (const float*)din[0], (const float*)din[1], (const float*)din[2],
                             (const float*)din[3], (const float*)din[4], (const uint8_t*)din[5], (uint8_t*)o[5], (float*)o[0], (float*)o[1],
                             (float*)o[2], (float*)o[3], (int32_t*)o[4], log_scale_factor, viewing_cos_limit, n_levels};
        og::k_frustum<<<dim3((max_n + 255) / 256, n_frames), 256, 0, st>>>(A);
        m->last_launches = 1;
    }
    cudaError_t e = cudaStreamSynchronize(st);   // the staging buffers are reused by the next call
    if (e != cudaSuccess) return bail(og_fail(ORBGPU_ERR_CUDA, std::string("mappoint_set_project: ") + cudaGetErrorString(e)));
    *out = mp;
    return ORBGPU_OK;
}

// og_octree.cuh — DistributeOctTree (ORBextractor.cc:539-763) as a block-cooperative state machine.
//
// The reference grows a std::list of quadtree nodes with push_front/erase and pointer-chasing; here the
// same list is an array in list order and every step of the algorithm is a data-parallel map over keys or
// nodes plus prefix sums:
//   * a node owns a contiguous segment of the key array; DivideNode (:481-537) is a stable 4-way partition of
//     that segment, computed for ALL dividing nodes at once from one block-wide scan of per-key class counts;
//   * "push_front children / erase parent" becomes: new position = T-1-creationIndex for children and
//     T + (rank among survivors) for the nodes that stay, T = number of children created;
//   * the careful phase (:676-737: sort by size, expand from the back, stop at N) is a rank computation plus a
//     prefix sum of "nodes gained" in processing order and a search for the first prefix that reaches N.
// Tie-break of the size sort: creation order (the oracle's documented patch of the pointer compare at :684).
//
// The code is written once and compiled twice: as device code run by one CTA (OG_FOR = block-stride loop,
// OG_SYNC = __syncthreads) and as host code (OG_FOR = plain loop) used ONLY by tests/host_model to check the
// list-order logic on CPU before a GPU is available.  The host build is not part of liborbgpu.so.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define OG_DEVICE_PASS 1
#define OG_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define OG_SYNC() __syncthreads()
#define OG_ONE if (threadIdx.x == 0)
#else
#define OG_DEVICE_PASS 0
#define OG_FOR(i, n) for (int i = 0; i < (n); ++i)
#define OG_SYNC() ((void)0)
#define OG_ONE
#endif

// Sections in which every thread walks its own contiguous chunk of the key array.  On the host the
// "threads" are emulated by a loop (OG_HOST_THREADS of them) so that chunk-boundary logic is exercised too.
#if OG_DEVICE_PASS
#define OG_THREADS_BEGIN(tt, TT) { const int TT = blockDim.x; const int tt = threadIdx.x; {
#define OG_THREADS_END }}
#define OG_NTHREADS() ((int)blockDim.x)
#else
#ifndef OG_HOST_THREADS
#define OG_HOST_THREADS 7
#endif
#define OG_THREADS_BEGIN(tt, TT) { const int TT = OG_HOST_THREADS; for (int tt = 0; tt < TT; ++tt) {
#define OG_THREADS_END }}
#define OG_NTHREADS() (OG_HOST_THREADS)
#endif
// per-thread chunk length; the device blocks are powers of two (128 / 256 / 512 threads): a shift instead of a division
#if OG_DEVICE_PASS
#define OG_DIV_THREADS(x, TT) ((x) >> (31 - __clz(TT)))
#else
#define OG_DIV_THREADS(x, TT) ((x) / (TT))
#endif
#define OG_CHUNK(M, tt, TT, lo, hi)                       \
    const int og_chunk_ = OG_DIV_THREADS((M) + (TT) - 1, TT); \
    const int lo = (tt) * og_chunk_ < (M) ? (tt) * og_chunk_ : (M); \
    const int hi = lo + og_chunk_ < (M) ? lo + og_chunk_ : (M);

#if defined(__CUDACC__)
#define OG_HD __host__ __device__ __forceinline__
#else
#define OG_HD inline
#endif

namespace og {

struct OtNode {
    int16_t x0, y0, x1, y1;  // UL.x, UL.y, UR.x (= BR.x), BR.y
    int32_t start, count;    // segment of the key array
};

// Per-node scratch for one division pass.
struct alignas(8) OtTmp {
    // header: one 8-byte load gives the key loops everything they need to classify a key of this node
    int16_t sx, sy;   // split lines
    uint8_t cand, processed;
    uint16_t pad;
    int32_t pre[4];   // class counts of keys BEFORE this node's segment (scan value at segment start)
    int32_t cnt[4];   // keys per child
    int32_t ord;      // index in processing order, -1 = not a candidate
    int32_t qbase;    // creation index of the first child
    int32_t newpos;   // list position in the next list (survivors) / unused
    int32_t pad2;
};

// sx | sy << 16 in .x, cand | processed << 8 in .y
struct OtHead { int sx, sy; bool cand, processed; };
OG_HD OtHead ot_head(const OtTmp* t) {
    OtHead h;
#if OG_DEVICE_PASS
    const uint2 v = *reinterpret_cast<const uint2*>(t);
    h.sx = (int)(int16_t)(v.x & 0xffffu);
    h.sy = (int)(int16_t)(v.x >> 16);
    h.cand = (v.y & 0xffu) != 0;
    h.processed = ((v.y >> 8) & 0xffu) != 0;
#else
    h.sx = t->sx; h.sy = t->sy; h.cand = t->cand != 0; h.processed = t->processed != 0;
#endif
    return h;
}
constexpr int kOtBatch = 2;   // keys a thread has in flight in the key loops (measured on B200: 2, 3 and 4 are equal, 6 and 8 slower)

// Block-shared scalars.
struct OtShared {
    int n;        // nodes in the list
    int nR;       // expandable nodes recorded by the last pass (nToExpand / vSizeAndPointerToNode.size())
    int T;        // children created by the current pass
    int nRnew;
    int ostar;    // last processed index in processing order
    int m;        // candidates in the current pass
    int scan_total;
    int warp_sums[32];
    int warp_sums4[32][4];
    int finish;
    int mode;
};

// Workspace (global memory, L1-cached; one per (frame, level)).
struct OtWork {
    uint32_t* kxy[2];    // [cap] packed keys: y << 16 | x   (coordinates relative to minBorder)
    uint8_t* kresp[2];   // [cap] FAST score
    uint16_t* knode[2];  // [cap] list index of the node that owns the key
    OtNode* nodes[2];    // [node_cap]
    OtTmp* tmp;          // [node_cap]
    int32_t* R[2];       // [node_cap] list positions of expandable nodes, creation order
    int32_t* ordv;       // [node_cap] scratch indexed by processing order
    int32_t* ordv2;      // [node_cap]
    int32_t* surv;       // [node_cap] survivor flags / ranks
    int32_t* thr;        // [block threads * 4] per-thread class totals for the key scan
    int cap, node_cap;
};

// ---- block-wide exclusive scan, in place, of a[0..n) ------------------------------------------------------
// Device: each thread owns a contiguous chunk; host: sequential.  total -> sh->scan_total.
OG_HD void block_exscan(int32_t* a, int n, OtShared* sh) {
#if OG_DEVICE_PASS
    const int T = blockDim.x, t = threadIdx.x;
    const int chunk = OG_DIV_THREADS(n + T - 1, T);
    const int lo = t * chunk < n ? t * chunk : n;
    const int hi = lo + chunk < n ? lo + chunk : n;
    int s = 0;
    for (int i = lo; i < hi; ++i) s += a[i];
    // inclusive scan of s across the block
    const int lane = t & 31, w = t >> 5;
    int v = s;
    for (int d = 1; d < 32; d <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    if (lane == 31) sh->warp_sums[w] = v;
    __syncthreads();
    if (w == 0) {
        int ws = lane < (T + 31) / 32 ? sh->warp_sums[lane] : 0;
        int x = ws;
        for (int d = 1; d < 32; d <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += u;
        }
        sh->warp_sums[lane] = x - ws;  // exclusive warp offsets
        if (lane == 31) sh->scan_total = x;
    }
    __syncthreads();
    int run = v - s + sh->warp_sums[w];
    for (int i = lo; i < hi; ++i) {
        int x = a[i];
        a[i] = run;
        run += x;
    }
    __syncthreads();
#else
    int run = 0;
    for (int i = 0; i < n; ++i) {
        int x = a[i];
        a[i] = run;
        run += x;
    }
    sh->scan_total = run;
#endif
}

OG_HD int ot_ceil_half(int d) { return (d + 1) >> 1; }  // ceil((float)d/2) for d >= 0 (:483-484)

// One division pass over the current list (cur) producing the next list (cur^1).
//   mode 0: full pass (:606-665) — every node with more than one key divides, in list order.
//   mode 1: careful sweep (:676-737) — the nodes in R[rcur] divide in order of decreasing size (ties: later
//           creation first) until the list reaches N nodes.
// Returns nothing; updates sh->n, sh->nR, and the buffers.  `cur` / `rcur` select the ping-pong halves.
OG_HD void ot_pass(const OtWork& W, OtShared* sh, int cur, int rcur, int M, int N, int mode) {
    OtNode* nodes = W.nodes[cur];
    OtNode* nnodes = W.nodes[cur ^ 1];
    OtTmp* tmp = W.tmp;
    const int n = sh->n;
    const int nR = sh->nR;

    // 1. candidates and split lines
    OG_FOR(i, n) {
        OtTmp& t = tmp[i];
        t.cand = (mode == 0) ? (nodes[i].count > 1) : 0;
        t.processed = 0;
        t.ord = -1;
        t.sx = (int16_t)(nodes[i].x0 + ot_ceil_half(nodes[i].x1 - nodes[i].x0));
        t.sy = (int16_t)(nodes[i].y0 + ot_ceil_half(nodes[i].y1 - nodes[i].y0));
    }
    OG_SYNC();
    if (mode == 1) {
        OG_FOR(j, nR) tmp[W.R[rcur][j]].cand = 1;
        OG_SYNC();
    }

    // 2. per-key child class + block scan of the four class counters (chunked: each thread keeps only its
    //    chunk totals; values at segment boundaries are written to the owning node)
    const uint32_t* kxy = W.kxy[cur];
    const uint16_t* knode = W.knode[cur];
    OG_THREADS_BEGIN(tt, TT)
        OG_CHUNK(M, tt, TT, lo, hi)
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        for (int p0 = lo; p0 < hi; p0 += kOtBatch) {
            int ii[kOtBatch];
            uint32_t kk[kOtBatch];
            OtHead hh[kOtBatch];
#pragma unroll
            for (int u = 0; u < kOtBatch; ++u) {
                const int p = p0 + u < hi ? p0 + u : hi - 1;
                ii[u] = knode[p];
                kk[u] = kxy[p];
            }
#pragma unroll
            for (int u = 0; u < kOtBatch; ++u) hh[u] = ot_head(&tmp[ii[u]]);
#pragma unroll
            for (int u = 0; u < kOtBatch; ++u) {
                if (p0 + u >= hi || !hh[u].cand) continue;
                const int x = (int)(kk[u] & 0xffffu), y = (int)(kk[u] >> 16);
                const int c = (x >= hh[u].sx ? 1 : 0) + (y >= hh[u].sy ? 2 : 0);
                c0 += (c == 0); c1 += (c == 1); c2 += (c == 2); c3 += (c == 3);
            }
        }
        W.thr[tt * 4 + 0] = c0; W.thr[tt * 4 + 1] = c1; W.thr[tt * 4 + 2] = c2; W.thr[tt * 4 + 3] = c3;
    OG_THREADS_END
    OG_SYNC();
    // exclusive scan over threads of each of the 4 counters
#if OG_DEVICE_PASS
    {
        const int tt = threadIdx.x, lane = tt & 31, w = tt >> 5, nwarp = (blockDim.x + 31) >> 5;
        int v[4], own[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { own[k] = W.thr[tt * 4 + k]; v[k] = own[k]; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int u = __shfl_up_sync(0xffffffffu, v[k], d);
                if (lane >= d) v[k] += u;
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sh->warp_sums4[w][k] = v[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int off = 0;
            for (int w2 = 0; w2 < nwarp; ++w2) off += (w2 < w) ? sh->warp_sums4[w2][k] : 0;
            W.thr[tt * 4 + k] = off + v[k] - own[k];
        }
    }
#else
    OG_FOR(k, 4) {
        const int TTn = OG_NTHREADS();
        int run = 0;
        for (int t2 = 0; t2 < TTn; ++t2) {
            int x = W.thr[t2 * 4 + k];
            W.thr[t2 * 4 + k] = run;
            run += x;
        }
    }
#endif
    OG_SYNC();
    OG_THREADS_BEGIN(tt, TT)
        OG_CHUNK(M, tt, TT, lo, hi)
        int c[4] = {W.thr[tt * 4 + 0], W.thr[tt * 4 + 1], W.thr[tt * 4 + 2], W.thr[tt * 4 + 3]};
        // a node's segment starts / ends where the owner of the neighbouring key differs (no look-up of the node itself)
        int prev = lo > 0 && lo < hi ? (int)knode[lo - 1] : -1;
        for (int p0 = lo; p0 < hi; p0 += kOtBatch) {
            int ii[kOtBatch], nx[kOtBatch];
            uint32_t kk[kOtBatch];
            OtHead hh[kOtBatch];
#pragma unroll
            for (int u = 0; u < kOtBatch; ++u) {
                const int p = p0 + u < hi ? p0 + u : hi - 1;
                ii[u] = knode[p];
                kk[u] = kxy[p];
                nx[u] = p + 1 < M ? (int)knode[p + 1] : -1;
            }
#pragma unroll
            for (int u = 0; u < kOtBatch; ++u) hh[u] = ot_head(&tmp[ii[u]]);
#pragma unroll
            for (int u = 0; u < kOtBatch; ++u) {
                const int p = p0 + u, i = ii[u];
                if (p >= hi) continue;
                const int nxt = nx[u];
                if (hh[u].cand) {
                    if (i != prev) { tmp[i].pre[0] = c[0]; tmp[i].pre[1] = c[1]; tmp[i].pre[2] = c[2]; tmp[i].pre[3] = c[3]; }
                    const int x = (int)(kk[u] & 0xffffu), y = (int)(kk[u] >> 16);
                    const int cl = (x >= hh[u].sx ? 1 : 0) + (y >= hh[u].sy ? 2 : 0);
                    c[cl]++;
                    if (i != nxt) {
                        // pre[] of this node is written by the thread that owns the segment start (possibly another
                        // thread), so the end values are stashed in cnt and pre is subtracted after the sync
                        tmp[i].cnt[0] = c[0]; tmp[i].cnt[1] = c[1]; tmp[i].cnt[2] = c[2]; tmp[i].cnt[3] = c[3];
                    }
                }
                prev = i;
            }
        }
    OG_THREADS_END
    OG_SYNC();
    OG_FOR(i, n) {
        if (tmp[i].cand) for (int k = 0; k < 4; ++k) tmp[i].cnt[k] -= tmp[i].pre[k];
    }
    OG_SYNC();

    // 3. processing order of the candidates
    int m;
    if (mode == 0) {
        OG_FOR(i, n) W.ordv[i] = tmp[i].cand;
        OG_SYNC();
        block_exscan(W.ordv, n, sh);
        m = sh->scan_total;
        OG_FOR(i, n) if (tmp[i].cand) tmp[i].ord = W.ordv[i];
        OG_SYNC();
    } else {
        m = nR;
        const int32_t* R = W.R[rcur];
        int32_t* sz = W.ordv2;   // sizes of the expandable nodes, gathered once: the ranking below reads them nR times each
        OG_FOR(j, nR) sz[j] = nodes[R[j]].count;
        OG_SYNC();
        OG_FOR(j, nR) {
            const int cj = sz[j];
            int r = 0;
            for (int j2 = 0; j2 < nR; ++j2) {
                const int c2 = sz[j2];
                r += (c2 > cj) || (c2 == cj && j2 > j);
            }
            tmp[R[j]].ord = r;
        }
        OG_SYNC();
    }

    // 4. nodes gained per candidate, in processing order; cut-off (careful phase only)
    OG_FOR(i, n) {
        if (tmp[i].cand) {
            const int nc = (tmp[i].cnt[0] > 0) + (tmp[i].cnt[1] > 0) + (tmp[i].cnt[2] > 0) + (tmp[i].cnt[3] > 0);
            W.ordv[tmp[i].ord] = nc - 1;
        }
    }
    OG_ONE { sh->ostar = m - 1; }
    OG_SYNC();
    if (mode == 1) {
        block_exscan(W.ordv, m, sh);  // exclusive prefix of gains
        OG_FOR(i, n) {
            if (tmp[i].cand) {
                const int nc = (tmp[i].cnt[0] > 0) + (tmp[i].cnt[1] > 0) + (tmp[i].cnt[2] > 0) + (tmp[i].cnt[3] > 0);
                const int o = tmp[i].ord;
                const int before = n + W.ordv[o];          // list size before dividing this node
                const int after = before + nc - 1;         // ... and after (:730 break test)
                if (after >= N && before < N) sh->ostar = o;
            }
        }
        OG_SYNC();
    }
    const int ostar = sh->ostar;

    // 5. creation indices (children) and R indices (expandable children) in processing order
    OG_FOR(i, n) {
        if (tmp[i].cand) {
            const int o = tmp[i].ord;
            const bool pr = o <= ostar;
            tmp[i].processed = pr;
            int nc = 0, ne = 0;
            for (int k = 0; k < 4; ++k) { nc += tmp[i].cnt[k] > 0; ne += tmp[i].cnt[k] > 1; }
            W.ordv[o] = pr ? nc : 0;
            W.ordv2[o] = pr ? ne : 0;
        }
    }
    OG_SYNC();
    block_exscan(W.ordv, m, sh);
    const int T = sh->scan_total;
    block_exscan(W.ordv2, m, sh);
    const int nRnew = sh->scan_total;
    OG_SYNC();

    // 6. survivors keep their relative order behind the T new children
    int32_t* surv = W.surv;
    OG_FOR(i, n) surv[i] = tmp[i].processed ? 0 : 1;
    OG_SYNC();
    block_exscan(surv, n, sh);
    const int nsurv = sh->scan_total;

    // 7. emit nodes
    int32_t* Rn = W.R[rcur ^ 1];
    OG_FOR(i, n) {
        const OtNode nd = nodes[i];
        OtTmp& t = tmp[i];
        if (!t.processed) {
            t.newpos = T + surv[i];
            nnodes[t.newpos] = nd;
        } else {
            int q = W.ordv[t.ord];
            int rq = W.ordv2[t.ord];
            int start = nd.start;
            for (int k = 0; k < 4; ++k) {
                const int c = t.cnt[k];
                if (c > 0) {
                    OtNode ch;
                    ch.x0 = (k & 1) ? t.sx : nd.x0;
                    ch.x1 = (k & 1) ? nd.x1 : t.sx;
                    ch.y0 = (k & 2) ? t.sy : nd.y0;
                    ch.y1 = (k & 2) ? nd.y1 : t.sy;
                    ch.start = start;
                    ch.count = c;
                    const int pos = T - 1 - q;
                    nnodes[pos] = ch;
                    if (c > 1) Rn[rq++] = pos;
                    ++q;
                }
                start += c;
            }
            // move records for the key pass below: a key of class cl with running class count c goes to pre[cl] + c and
            // belongs to list node cnt[cl] (pre / cnt have served their purpose)
            {
                const int q0 = W.ordv[t.ord];
                int off = 0, before = 0;
                for (int k = 0; k < 4; ++k) {
                    const int c = t.cnt[k];
                    t.pre[k] = nd.start + off - t.pre[k];
                    t.cnt[k] = T - 1 - (q0 + before);
                    off += c;
                    before += c > 0;
                }
            }
        }
    }
    OG_SYNC();

    // 8. move keys (stable partition inside every processed node; everything else stays in place)
    {
        uint32_t* kxy2 = W.kxy[cur ^ 1];
        uint8_t* kr2 = W.kresp[cur ^ 1];
        uint16_t* kn2 = W.knode[cur ^ 1];
        const uint8_t* kr = W.kresp[cur];
        OG_THREADS_BEGIN(tt, TT)
            OG_CHUNK(M, tt, TT, lo, hi)
            int c[4] = {W.thr[tt * 4 + 0], W.thr[tt * 4 + 1], W.thr[tt * 4 + 2], W.thr[tt * 4 + 3]};
            for (int p0 = lo; p0 < hi; p0 += kOtBatch) {
                int ii[kOtBatch], cls[kOtBatch], run[kOtBatch], np[kOtBatch], nn[kOtBatch];
                uint32_t kk[kOtBatch];
                uint8_t rr[kOtBatch];
                OtHead hh[kOtBatch];
#pragma unroll
                for (int u = 0; u < kOtBatch; ++u) {
                    const int p = p0 + u < hi ? p0 + u : hi - 1;
                    ii[u] = knode[p];
                    kk[u] = kxy[p];
                    rr[u] = kr[p];
                }
#pragma unroll
                for (int u = 0; u < kOtBatch; ++u) hh[u] = ot_head(&tmp[ii[u]]);
                // running class counts first (sequential over the batch), then the dependent loads of all keys together
#pragma unroll
                for (int u = 0; u < kOtBatch; ++u) {
                    cls[u] = -1;
                    run[u] = 0;
                    if (p0 + u < hi && hh[u].cand) {
                        const int x = (int)(kk[u] & 0xffffu), y = (int)(kk[u] >> 16);
                        const int cl = (x >= hh[u].sx ? 1 : 0) + (y >= hh[u].sy ? 2 : 0);
                        cls[u] = cl;
                        run[u] = c[cl];
                        c[cl]++;
                    }
                }
#pragma unroll
                for (int u = 0; u < kOtBatch; ++u) {
                    const OtTmp& t = tmp[ii[u]];
                    if (cls[u] >= 0 && hh[u].processed) {
                        np[u] = t.pre[cls[u]] + run[u];
                        nn[u] = t.cnt[cls[u]];
                    } else {
                        np[u] = p0 + u;
                        nn[u] = t.newpos;
                    }
                }
#pragma unroll
                for (int u = 0; u < kOtBatch; ++u) {
                    if (p0 + u >= hi) continue;
                    kxy2[np[u]] = kk[u];
                    kr2[np[u]] = rr[u];
                    kn2[np[u]] = (uint16_t)nn[u];
                }
            }
        OG_THREADS_END
    }
    OG_SYNC();
    OG_ONE {
        sh->n = T + nsurv;
        sh->nR = nRnew;
    }
    OG_SYNC();
}

// Whole DistributeOctTree.  The M candidate keys must sit in W.kxy[1] / W.kresp[1] in emission order (they are
// partitioned by root into half 0).  nIni / hX are the host-computed root count and root width (:543-545).
// Output: the selected key of every final node, in list order; returns the count.
OG_HD int ot_run(const OtWork& W, OtShared* sh, int M, int nIni, float hX, int height, int N, uint32_t* out_xy,
                 uint8_t* out_resp, int out_cap) {
    if (M == 0) return 0;
    // ---- roots (:547-585): bucket by (int)(x / hX), stable; empty roots are erased -------------------------
    {
        uint16_t* root = W.knode[1];
        OG_FOR(p, M) {
            const float x = (float)(W.kxy[1][p] & 0xffffu);
#if OG_DEVICE_PASS
            root[p] = (uint16_t)(int)__fdiv_rn(x, hX);
#else
            root[p] = (uint16_t)(int)(x / hX);
#endif
        }
        OG_SYNC();
        int base = 0, nn = 0;
        for (int r = 0; r < nIni; ++r) {
            OG_THREADS_BEGIN(tt, TT)
                OG_CHUNK(M, tt, TT, lo, hi)
                int cnt = 0;
                for (int p = lo; p < hi; ++p) cnt += (root[p] == r);
                W.thr[tt] = cnt;
            OG_THREADS_END
            OG_SYNC();
            block_exscan(W.thr, OG_NTHREADS(), sh);
            const int total = sh->scan_total;
            OG_THREADS_BEGIN(tt, TT)
                OG_CHUNK(M, tt, TT, lo, hi)
                int run = W.thr[tt];
                for (int p = lo; p < hi; ++p) {
                    if (root[p] == r) {
                        const int np = base + run++;
                        W.kxy[0][np] = W.kxy[1][p];
                        W.kresp[0][np] = W.kresp[1][p];
                        W.knode[0][np] = (uint16_t)nn;
                    }
                }
            OG_THREADS_END
            if (total > 0) {
                OG_ONE {
                    OtNode nd;
                    nd.x0 = (int16_t)(int)(hX * (float)r);
                    nd.x1 = (int16_t)(int)(hX * (float)(r + 1));
                    nd.y0 = 0;
                    nd.y1 = (int16_t)height;
                    nd.start = base;
                    nd.count = total;
                    W.nodes[0][nn] = nd;
                }
                ++nn;
            }
            base += total;
            OG_SYNC();
        }
        OG_ONE { sh->n = nn; sh->nR = 0; }
        OG_SYNC();
    }
    int cur = 0;
    int rcur = 0;
    bool finish = false;
    while (!finish) {
        const int prevSize = sh->n;
        ot_pass(W, sh, cur, rcur, M, N, 0);
        cur ^= 1; rcur ^= 1;
        const int n = sh->n, nExp = sh->nR;
        if (n >= N || n == prevSize) finish = true;
        else if (n + nExp * 3 > N) {
            while (!finish) {
                const int prev2 = sh->n;
                ot_pass(W, sh, cur, rcur, M, N, 1);
                cur ^= 1; rcur ^= 1;
                if (sh->n >= N || sh->n == prev2) finish = true;
            }
        }
    }

    // best key per node, first maximum wins (:742-760)
    const int n = sh->n;
    const OtNode* nodes = W.nodes[cur];
    OG_FOR(i, n) {
        if (i < out_cap) {
            const int s = nodes[i].start, c = nodes[i].count;
            int best = s;
            int br = W.kresp[cur][s];
            for (int p = s + 1; p < s + c; ++p) {
                const int r = W.kresp[cur][p];
                if (r > br) { br = r; best = p; }
            }
            out_xy[i] = W.kxy[cur][best];
            out_resp[i] = (uint8_t)br;
        }
    }
    OG_SYNC();
    return n < out_cap ? n : out_cap;
}

}  // namespace og

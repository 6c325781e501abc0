// og_octree2.cuh — DistributeOctTree (ORBextractor.cc:539-763) without division passes.
//
// The reference's full passes (:590-665) divide EVERY node that holds more than one key, so after d passes the list
// consists of the non-empty cells of a fixed quadtree: the cell of a key at depth d follows from its coordinates alone
// (root = (int)(x / hX), then d times "which side of the ceil-half split lines", DivideNode :481-537), and a node whose
// cell holds one key stays in the list as it is.  Hence
//   * list size after d passes          n_d    = number of non-empty depth-d cells,
//   * nToExpand of pass d (:617-653)    nExp_d = number of depth-d cells with more than one key,
// both read off ONE histogram of the keys over the depth-Dh cells (shared-memory atomics) reduced 4 -> 1 per depth.  The
// pass loop (:588-665) becomes a scan over d for the first depth D with n_D >= N or n_D == n_(D-1) (finish) or
// n_D + 3 nExp_D > N (careful phase).
//
// List order.  A pass walks the list front to back and pushes the children of every dividing node to the FRONT in the
// order n1..n4, undivided (single-key) nodes keep their place behind them.  So the list after pass d is
//   [children created by pass d, last created first] ++ [single-key nodes created by pass d-1, in their order] ++ ...
//   ++ [single-key roots in root order],
// and "last created first" unrolls into a lexicographic order of the cell paths (root, q_1 .. q_d) whose direction
// alternates per component: q_j descends when d - j is even and ascends otherwise, the root component goes with q_1 (the
// roots are pushed to the BACK, :556).  Enumerating a depth's cells in that order is an XOR of the path with 0x33..3 and a
// flip of the root index; list positions are a prefix sum of "this cell is in the list" over that enumeration.
//
// The careful phase (:666-737: candidates sorted by size, divided from the largest until the list holds N nodes) works on
// the same cell counts: no key ever moves.  Creation order of the candidates = reverse list order, so the reference's
// (size, pointer) sort — with the oracle's documented creation-order tie-break — is "size descending, then list position
// ascending" in its first round and "size descending, later pushed first" afterwards.
//
// Finally every key walks down its path to the first cell that is a list node and competes for it with
// (response << 24 | ~emission index): the maximum is the reference's "first maximum in vKeys order" (:742-760), because
// DivideNode keeps the keys of a node in emission order.
//
// Anything the histogram cannot hold (a decision deeper than Dh, 32767 or more keys in a careful-phase node, ...) returns -1 and the caller
// runs the division-pass state machine of og_octree.cuh instead (same results, it is the general path).
//
// Written once, compiled twice like og_octree.cuh (device: one CTA; host: tests/host_model only).
#pragma once
#include <string.h>

#include "og_octree.cuh"
#include "og_types.h"

namespace og {

constexpr int kOt2MaxDepth = 8;

struct Ot2Shared {
    int nd[kOt2MaxDepth + 2];     // non-empty cells per depth
    int nexp[kOt2MaxDepth + 2];   // cells with more than one key per depth
    int D, careful, fallback;
    int Pn;                       // candidates divided by the current careful round
};

// Deepest histogram level for n_ini roots: the largest Dh (<= kOt2MaxDepth) with n_ini * 4^Dh <= budget cells.
OG_HD int ot2_depth(int n_ini, int budget) {
    int d = 0;
    long long c = n_ini;
    while (d < kOt2MaxDepth && c * 4 <= budget) { c *= 4; ++d; }
    return d;
}
// 16-bit counters of all depths 0..Dh (deepest first, so every depth starts at an even index).
OG_HD int ot2_hist_cells(int n_ini, int Dh) {
    int tot = 0;
    for (int d = 0; d <= Dh; ++d) tot += n_ini << (2 * d);
    return tot;
}
// Bytes of the histogram and the small per-node arrays behind it (list x2, candidates x2, sizes, child counts, order, 3 scan
// arrays, best).  node_cap here = max(node capacity of the level, threads of the block): the scan arrays also carry
// per-thread totals.
// kcap = keys whose path code and response are cached between the two key passes (3 bytes each; a multiple of 16).
OG_HD size_t ot2_smem_bytes(int n_ini, int Dh, int node_cap, int kcap) {
    const size_t hist = ((size_t)ot2_hist_cells(n_ini, Dh) * 2 + 15) & ~size_t(15);
    const size_t nc = ((size_t)node_cap + 7) & ~size_t(3);
    return hist + nc * (4 + 4 + 2 + 2 + 2 + 2 + 2 + 4 + 4 + 4 + 4) + (size_t)kcap * 3;
}

struct Ot2Work {
    uint16_t* hist;                    // per depth: keys per cell (later: 0x8000 | list position for list nodes), deepest level first
    int n_ini, top;                    // top = 4^(Dh+1)
    OG_HD uint16_t* cnt(int d) const { return hist + n_ini * ((top - (4 << (2 * d))) / 3); }   // n_ini * (4^(d+1) + .. + 4^Dh) cells come first
    uint32_t* list[2];                 // list node = depth << 28 | cell
    uint16_t* C[2];                    // candidates of a careful round (list positions), creation order
    uint16_t* sz;                      // candidate sizes
    uint16_t* kids;                    // nc | ne << 8: non-empty / expandable children of a candidate
    uint16_t* ordj;                    // candidate index by processing rank
    int32_t* sa;                       // scan arrays
    int32_t* sb;
    int32_t* sc;
    uint32_t* best;                    // per list node: response << 24 | (0xffffff - emission index)
    uint16_t* kcode;                   // [kcap] first the owning cell of a key (Ot2CellKeys), then its depth-Dh cell
    uint8_t* kres;                     // [kcap] response
    int kcap;
};

OG_HD Ot2Work ot2_carve(uint8_t* mem, int n_ini, int Dh, int node_cap, int kcap) {
    Ot2Work Q;
    Q.hist = reinterpret_cast<uint16_t*>(mem);
    Q.n_ini = n_ini;
    Q.top = 4 << (2 * Dh);
    uint8_t* p = mem + (((size_t)ot2_hist_cells(n_ini, Dh) * 2 + 15) & ~size_t(15));
    const size_t nc = ((size_t)node_cap + 7) & ~size_t(3);   // + 4: the rank loop reads whole vectors
    Q.list[0] = (uint32_t*)p; p += nc * 4;
    Q.list[1] = (uint32_t*)p; p += nc * 4;
    Q.sa = (int32_t*)p; p += nc * 4;
    Q.sb = (int32_t*)p; p += nc * 4;
    Q.sc = (int32_t*)p; p += nc * 4;
    Q.best = (uint32_t*)p; p += nc * 4;
    Q.C[0] = (uint16_t*)p; p += nc * 2;
    Q.C[1] = (uint16_t*)p; p += nc * 2;
    Q.sz = (uint16_t*)p; p += nc * 2;
    Q.kids = (uint16_t*)p; p += nc * 2;
    Q.ordj = (uint16_t*)p; p += nc * 2;
    Q.kcode = (uint16_t*)p; p += (size_t)kcap * 2;
    Q.kres = p;
    Q.kcap = kcap;
    return Q;
}

// Walks a key down the quadtree (DivideNode's split lines, :481-537).
struct Ot2Path {
    int x, y, x0, x1, y0, y1, cell;
    OG_HD void start(uint32_t xy, float hX, int height) {
        x = (int)(xy & 0xffffu);
        y = (int)(xy >> 16);
#if OG_DEVICE_PASS
        cell = (int)__fdiv_rn((float)x, hX);   // vpIniNodes[kp.pt.x / hX] (:565)
#else
        cell = (int)((float)x / hX);
#endif
        x0 = (int)(hX * (float)cell);          // :550-551
        x1 = (int)(hX * (float)(cell + 1));
        y0 = 0;
        y1 = height;
    }
    OG_HD void step() {
        const int sx = x0 + ot_ceil_half(x1 - x0), sy = y0 + ot_ceil_half(y1 - y0);
        const bool rx = x >= sx, ry = y >= sy;
        cell = cell * 4 + (rx ? 1 : 0) + (ry ? 2 : 0);
        if (rx) x0 = sx; else x1 = sx;
        if (ry) y0 = sy; else y1 = sy;
    }
};

// Cell at position f of depth dd's enumeration in list order (see the header): flipped root, path XOR 0x33..3.
OG_HD int ot2_unflip(int f, int dd, int n_ini) {
    const int sh = 2 * dd;
    const int r = f >> sh, p = f & ((1 << sh) - 1);
    return (((dd & 1) ? n_ini - 1 - r : r) << sh) | (p ^ (0x33333333 & ((1 << sh) - 1)));
}

#if OG_DEVICE_PASS
// +1 on the 16-bit counter idx (two per 32-bit word; no carry: counts stay below 2^15), returns the counter's old value
__device__ __forceinline__ int ot2_atomic_inc_u16(uint16_t* base, int idx) {
    const int sh = 16 * (idx & 1);
    const unsigned int old = atomicAdd(reinterpret_cast<unsigned int*>(base) + (idx >> 1), 1u << sh);
    return (int)((old >> sh) & 0xffffu);
}
#define OT2_ATOMIC_ADD_U16(base, idx) ot2_atomic_inc_u16((base), (idx))
#define OT2_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define OT2_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define OT2_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define OT2_ATOMIC_ADD_U16(base, idx) ((base)[(idx)]++)
#define OT2_ATOMIC_MAX(p, v) (*(p) = *(p) > (v) ? *(p) : (v))
#define OT2_ATOMIC_MIN(p, v) (*(p) = *(p) < (v) ? *(p) : (v))
#define OT2_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

// Where the M candidate keys come from, by emission index.
struct Ot2CompactKeys {   // plain arrays in emission order (tests, orbgpu_octree)
    const uint32_t* kxy;
    const uint8_t* kresp;
    OG_HD void prefill(uint16_t*, int) const {}
    OG_HD void get(int i, int, uint32_t& xy, int& resp) const { xy = kxy[i]; resp = kresp[i]; }
};
struct Ot2CellKeys {      // the per-cell slots k_fast_seg fills: coff = exclusive scan of the per-cell counts (emission order)
    const int32_t* coff;
    const Cell* cells;
    const uint32_t* cxy;
    const uint8_t* crr;
    int n_cells;
    // owner[i] = cell of key i for the first `cap` keys (block-wide; the caller synchronises afterwards): spares the first key
    // pass the binary search
    OG_HD void prefill(uint16_t* owner, int cap) const {
        if (n_cells > 65535) return;
        OG_FOR(ci, n_cells) {
            const int lo = coff[ci], hi = ci + 1 < n_cells ? coff[ci + 1] : 0x7fffffff;
            for (int i = lo; i < hi && i < cap; ++i) owner[i] = (uint16_t)ci;   // hi of the last cell: cut by the caller's cap (<= M)
        }
    }
    OG_HD void get(int i, int owner, uint32_t& xy, int& resp) const {
        int lo = owner;
        if (lo < 0) {
            int hi = n_cells;   // last cell whose offset is <= i (empty cells share their successor's offset and are skipped)
            lo = 0;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (coff[mid] <= i) lo = mid; else hi = mid;
            }
        }
        const int slot = cells[lo].slot + (i - coff[lo]);
#if OG_DEVICE_PASS
        xy = __ldg(cxy + slot);
        resp = __ldg(crr + slot);
#else
        xy = cxy[slot];
        resp = crr[slot];
#endif
    }
};

OG_HD bool keys_cells_limit(const Ot2CompactKeys&) { return false; }
OG_HD bool keys_cells_limit(const Ot2CellKeys& k) { return k.n_cells > 65535; }

// keys: the M candidate keys by emission index.  mem: ot2_smem_bytes() of (shared) memory, 16-byte aligned.
// Returns the number of selected keys (written in list order), or -1 when the general path has to run.
template <class Keys>
OG_HD int ot_run_direct(const Keys& keys, int M, uint8_t* mem, int Dh, int node_cap, int kcap, OtShared* sh, Ot2Shared* s2,
                        int nIni, float hX, int height, int N, uint32_t* out_xy, uint8_t* out_resp, int out_cap) {
    if (M == 0) return 0;
    // 15-bit counters.  The deepest three levels are counted by atomics (exact: such a cell never holds more keys than
    // pixels, checked here); the shallower counts are sums that saturate at 0x7fff, and a careful round that meets a saturated
    // size (its sort needs the exact value) hands over.  24-bit emission index.
    const int dA = Dh > 2 ? Dh - 2 : 0;   // depths >= dA: one atomic per key and depth
    if (Dh < 1 || M >= (1 << 24) || node_cap > 32767) return -1;
    if ((long long)(((int)hX >> dA) + 2) * ((height >> dA) + 2) > 32767) return -1;
    if (((long long)nIni << (2 * Dh)) > 65536 || keys_cells_limit(keys)) kcap = 0;   // 16-bit path codes / owners
    const Ot2Work Q = ot2_carve(mem, nIni, Dh, node_cap, kcap);
    const int kc = kcap < M ? kcap : M;   // cached keys

    // ---- histogram of the keys over the cells of depths dA..Dh ---------------------------------------------------------
    {
        const int vecs = (ot2_hist_cells(nIni, Dh) * 2 + 15) / 16;
#if OG_DEVICE_PASS
        uint4* z = reinterpret_cast<uint4*>(Q.hist);
        OG_FOR(i, vecs) z[i] = make_uint4(0u, 0u, 0u, 0u);
#else
        memset(Q.hist, 0, (size_t)vecs * 16);
#endif
        OG_FOR(i, kOt2MaxDepth + 2) { s2->nd[i] = 0; s2->nexp[i] = 0; }
        OG_ONE { s2->D = 0; s2->careful = 0; s2->fallback = 0; }
        keys.prefill(Q.kcode, kc);
    }
    OG_SYNC();
    {
        // a cell's first key makes it non-empty, its second one expandable: n_d and nExp_d of the atomically counted depths
        int nz0 = 0, ne0 = 0, nz1 = 0, ne1 = 0, nz2 = 0, ne2 = 0;
        OG_FOR(i, M) {
            Ot2Path p;
            uint32_t xy;
            int resp;
            keys.get(i, i < kc ? (int)Q.kcode[i] : -1, xy, resp);
            p.start(xy, hX, height);
            for (int d = 0; d < dA; ++d) p.step();
            int old = OT2_ATOMIC_ADD_U16(Q.cnt(dA), p.cell);
            nz0 += old == 0; ne0 += old == 1;
            if (dA + 1 <= Dh) {
                p.step();
                old = OT2_ATOMIC_ADD_U16(Q.cnt(dA + 1), p.cell);
                nz1 += old == 0; ne1 += old == 1;
            }
            if (dA + 2 <= Dh) {
                p.step();
                old = OT2_ATOMIC_ADD_U16(Q.cnt(dA + 2), p.cell);
                nz2 += old == 0; ne2 += old == 1;
            }
            if (i < kc) {   // p.cell is the depth-Dh cell now (dA + 2 >= Dh)
                Q.kcode[i] = (uint16_t)p.cell;
                Q.kres[i] = (uint8_t)resp;
            }
        }
        if (nz0) OT2_ATOMIC_ADD(&s2->nd[dA], nz0);
        if (ne0) OT2_ATOMIC_ADD(&s2->nexp[dA], ne0);
        if (nz1) OT2_ATOMIC_ADD(&s2->nd[dA + 1], nz1);
        if (ne1) OT2_ATOMIC_ADD(&s2->nexp[dA + 1], ne1);
        if (nz2) OT2_ATOMIC_ADD(&s2->nd[dA + 2], nz2);
        if (ne2) OT2_ATOMIC_ADD(&s2->nexp[dA + 2], ne2);
    }
    OG_SYNC();
    // ---- counts of the shallower depths, their n_d and nExp_d -------------------------------------------------------------
    for (int d = dA - 1; d >= 0; --d) {
        const int cells = nIni << (2 * d);
        int pz = 0, pe = 0;
        OG_FOR(c, cells) {
            const uint16_t* k = Q.cnt(d + 1) + 4 * c;
#if OG_DEVICE_PASS
            const uint2 kk = *reinterpret_cast<const uint2*>(k);   // 8-byte aligned: every depth starts at a multiple of 4 cells
            const int sum = (int)((kk.x & 0xffffu) + (kk.x >> 16) + (kk.y & 0xffffu) + (kk.y >> 16));
#else
            const int sum = k[0] + k[1] + k[2] + k[3];
#endif
            Q.cnt(d)[c] = (uint16_t)(sum < 0x7fff ? sum : 0x7fff);
            pz += sum > 0;
            pe += sum > 1;
        }
        if (pz) OT2_ATOMIC_ADD(&s2->nd[d], pz);
        if (pe) OT2_ATOMIC_ADD(&s2->nexp[d], pe);
        OG_SYNC();
    }
    // ---- the pass loop (:588-665) as a scan over depths ----------------------------------------------------------------
    OG_ONE {
        int D = 0, careful = 0;
        for (int d = 1; d <= Dh && !D; ++d) {
            if (s2->nd[d] >= N || s2->nd[d] == s2->nd[d - 1]) D = d;
            else if (s2->nd[d] + 3 * s2->nexp[d] > N) { D = d; careful = 1; }
        }
        s2->D = D;
        s2->careful = careful;
        s2->fallback = D == 0;
    }
    OG_SYNC();
    if (s2->fallback) return -1;
    const int D = s2->D;

    // ---- the list after D passes: enumeration of depths D, D-1 .. 0 in list order, prefix sum of the member flags -------
    int E = 0;   // entries of the concatenated enumeration
    for (int dd = 0; dd <= D; ++dd) E += nIni << (2 * dd);
    // member(dd, c): depth-D cells born by pass D, single-key cells born earlier, single-key roots
    auto member = [&](int e, int& dd, int& c, bool& multi) -> bool {
        int base = 0;
        dd = D;
        while (e >= base + (nIni << (2 * dd))) { base += nIni << (2 * dd); --dd; }
        c = ot2_unflip(e - base, dd, nIni);
        const int k = Q.cnt(dd)[c];
        multi = false;
        if (dd == 0) return k == 1;
        const bool born = k > 0 && Q.cnt(dd - 1)[c >> 2] > 1;
        if (dd == D) { multi = born && k > 1; return born; }
        return born && k == 1;
    };
    int n = 0, nC = 0;
    {
        // chunked two-pass scan; per-thread totals travel through the block scan of `scr`
        int32_t* scr = Q.sc;   // per-thread totals, members << 16 | expandable members (the arrays hold >= OG_NTHREADS() entries)
        OG_THREADS_BEGIN(tt, TT)
            OG_CHUNK(E, tt, TT, lo, hi)
            int m = 0, x = 0;
            for (int e = lo; e < hi; ++e) {
                int dd, c; bool multi;
                m += member(e, dd, c, multi);
                x += multi;
            }
            scr[tt] = (m << 16) | x;
        OG_THREADS_END
        OG_SYNC();
        block_exscan(scr, OG_NTHREADS(), sh);
        n = sh->scan_total >> 16;
        nC = sh->scan_total & 0xffff;
        if (n > node_cap) return -1;   // cannot happen (n <= max(N + 2, 4 nIni)); keeps the arrays safe
        OG_THREADS_BEGIN(tt, TT)
            OG_CHUNK(E, tt, TT, lo, hi)
            int pos = scr[tt] >> 16, mi = scr[tt] & 0xffff;
            for (int e = lo; e < hi; ++e) {
                int dd, c; bool multi;
                if (member(e, dd, c, multi)) {
                    Q.list[0][pos] = ((uint32_t)dd << 28) | (uint32_t)c;
                    if (multi) Q.C[0][nC - 1 - mi++] = (uint16_t)pos;   // creation order = reverse list order
                    ++pos;
                }
            }
        OG_THREADS_END
        OG_SYNC();
    }

    // ---- careful rounds (:666-737) ------------------------------------------------------------------------------------------
    int cur = 0, dep = D;
    bool careful = s2->careful != 0;
    while (careful) {
        if (dep + 1 > Dh) return -1;
        const int prev = n;
        const uint32_t* L = Q.list[cur];
        uint32_t* L2 = Q.list[cur ^ 1];
        const uint16_t* Cc = Q.C[cur];
        uint16_t* C2 = Q.C[cur ^ 1];
        const uint32_t cmask = (1u << 28) - 1u;
        // a. size and children of every candidate
        OG_FOR(j, nC) {
            const int c = (int)(L[Cc[j]] & cmask);
            Q.sz[j] = Q.cnt(dep)[c];
            if (Q.sz[j] == 0x7fff) s2->fallback = 1;
            const uint16_t* k = Q.cnt(dep + 1) + 4 * c;
            const int ncn = (k[0] > 0) + (k[1] > 0) + (k[2] > 0) + (k[3] > 0), nen = (k[0] > 1) + (k[1] > 1) + (k[2] > 1) + (k[3] > 1);
            Q.kids[j] = (uint16_t)(ncn | (nen << 8));
        }
        OG_ONE { s2->Pn = nC; }
        OG_FOR(i, n) Q.sc[i] = 1;   // survivor flags
        OG_SYNC();
        if (s2->fallback) return -1;
        // b. processing rank: larger size first, equal sizes: later created first (std::sort ascending walked from the back) =
        //    descending order of size << 16 | creation index
        OG_FOR(j, nC + 4) Q.sa[j] = j < nC ? (((int)Q.sz[j] << 16) | j) : 0;
        OG_SYNC();
        OG_FOR(j, nC) {
            const int kj = Q.sa[j];
            int r = 0;
            for (int j2 = 0; j2 < nC; j2 += 4) {
#if OG_DEVICE_PASS
                const int4 k4 = *reinterpret_cast<const int4*>(Q.sa + j2);
                r += (k4.x > kj) + (k4.y > kj) + (k4.z > kj) + (k4.w > kj);
#else
                for (int u = 0; u < 4; ++u) r += Q.sa[j2 + u] > kj;
#endif
            }
            Q.ordj[r] = (uint16_t)j;
        }
        OG_SYNC();
        // c. list size after every division, in processing order; the round stops at the first one that reaches N (:730)
        OG_FOR(r, nC) Q.sa[r] = (Q.kids[Q.ordj[r]] & 0xff) - 1;
        OG_SYNC();
        block_exscan(Q.sa, nC, sh);
        OG_FOR(r, nC) {
            const int after = n + Q.sa[r] + (Q.kids[Q.ordj[r]] & 0xff) - 1;
            if (after >= N) OT2_ATOMIC_MIN(&s2->Pn, r + 1);
        }
        OG_SYNC();
        const int Pn = s2->Pn;
        // d. push indices of the children / of the expandable children, in processing order
        OG_FOR(r, nC) {
            const int kd = Q.kids[Q.ordj[r]];
            Q.sb[r] = r < Pn ? (((kd & 0xff) << 16) | (kd >> 8)) : 0;
            if (r < Pn) Q.sc[Cc[Q.ordj[r]]] = 0;   // the divided node leaves the list
        }
        OG_SYNC();
        block_exscan(Q.sb, nC, sh);
        const int T = sh->scan_total >> 16, nCnew = sh->scan_total & 0xffff;
        block_exscan(Q.sc, n, sh);
        const int nsurv = sh->scan_total;
        if (T + nsurv > node_cap) return -1;   // cannot happen, see above
        // e. the new list: children, last pushed first, then the undivided nodes in their order
        OG_FOR(i, n) {
            const bool alive = (i + 1 < n ? Q.sc[i + 1] : nsurv) != Q.sc[i];
            if (alive) L2[T + Q.sc[i]] = L[i];
        }
        OG_FOR(r, Pn) {
            const int c = (int)(L[Cc[Q.ordj[r]]] & cmask);
            int q = Q.sb[r] >> 16, rq = Q.sb[r] & 0xffff;
            const uint16_t* k = Q.cnt(dep + 1) + 4 * c;
            for (int kk = 0; kk < 4; ++kk) {
                if (k[kk] > 0) {
                    const int pos = T - 1 - q;
                    L2[pos] = ((uint32_t)(dep + 1) << 28) | (uint32_t)(4 * c + kk);
                    if (k[kk] > 1) C2[rq++] = (uint16_t)pos;
                    ++q;
                }
            }
        }
        OG_SYNC();
        n = T + nsurv;
        nC = nCnew;
        cur ^= 1;
        ++dep;
        if (n >= N || n == prev) careful = false;
    }

    // ---- best key of every list node (:742-760) ---------------------------------------------------------------------------------
    {
        const uint32_t* L = Q.list[cur];
        OG_FOR(i, n) {
            const uint32_t nd = L[i];
            Q.cnt(nd >> 28)[nd & ((1u << 28) - 1u)] = (uint16_t)(0x8000u | (uint32_t)i);
            Q.best[i] = 0u;
        }
        OG_SYNC();
        OG_FOR(i, M) {
            uint32_t v;
            int resp;
            if (i < kc) {
                // cached path code: the depth-d cell is a shift away
                const int code = Q.kcode[i];
                resp = Q.kres[i];
                int d = 0;
                v = Q.cnt(0)[code >> (2 * Dh)];
                while (!(v & 0x8000u) && d < Dh) {
                    ++d;
                    v = Q.cnt(d)[code >> (2 * (Dh - d))];
                }
            } else {
                Ot2Path p;
                uint32_t xy;
                keys.get(i, -1, xy, resp);
                p.start(xy, hX, height);
                int d = 0;
                v = Q.cnt(0)[p.cell];
                while (!(v & 0x8000u) && d < Dh) {
                    p.step();
                    ++d;
                    v = Q.cnt(d)[p.cell];
                }
            }
            if (v & 0x8000u) OT2_ATOMIC_MAX(&Q.best[v & 0x7fffu], ((uint32_t)resp << 24) | (0xffffffu - (uint32_t)i));
        }
        OG_SYNC();
        OG_FOR(i, n) {
            if (i < out_cap) {
                const uint32_t b = Q.best[i];
                uint32_t xy;
                int resp;
                keys.get((int)(0xffffffu - (b & 0xffffffu)), -1, xy, resp);
                out_xy[i] = xy;
                out_resp[i] = (uint8_t)resp;
            }
        }
        OG_SYNC();
    }
    return n < out_cap ? n : out_cap;
}

}  // namespace og

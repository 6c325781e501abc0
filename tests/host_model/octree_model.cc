// TEST INFRASTRUCTURE: host compile of the block-cooperative octree state machine (og_octree.cuh) with
// OG_FOR = plain loop.  Lets the CPU test-suite check the list-order logic of the CUDA kernel's algorithm
// against the oracle without a GPU.  Never linked into liborbgpu.so.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../orb_slam2_with_comment_b200/csrc/og_octree.cuh"
#include "../../orb_slam2_with_comment_b200/csrc/og_octree2.cuh"

extern "C" int ogm_octree(const uint32_t* xy, const uint8_t* resp, int M, int width, int height, int N,
                          uint32_t* out_xy, uint8_t* out_resp, int out_cap) {
    using namespace og;
    const int nIni = (int)roundf((float)width / (float)height);
    if (nIni < 1) return -1;
    const float hX = (float)width / (float)nIni;
    const int node_cap = std::max(N + 3, 4 * nIni) + 8;
    const int cap = M > 0 ? M : 1;
    std::vector<uint32_t> kxy0(cap), kxy1(cap);
    std::vector<uint8_t> kr0(cap), kr1(cap);
    std::vector<uint16_t> kn0(cap), kn1(cap);
    std::vector<OtNode> n0(node_cap), n1(node_cap);
    std::vector<OtTmp> tmp(node_cap);
    std::vector<int32_t> R0(node_cap), R1(node_cap), ordv(node_cap), ordv2(node_cap), surv(node_cap), thr(4 * 64);
    OtWork W;
    W.kxy[0] = kxy0.data(); W.kxy[1] = kxy1.data();
    W.kresp[0] = kr0.data(); W.kresp[1] = kr1.data();
    W.knode[0] = kn0.data(); W.knode[1] = kn1.data();
    W.nodes[0] = n0.data(); W.nodes[1] = n1.data();
    W.tmp = tmp.data();
    W.R[0] = R0.data(); W.R[1] = R1.data();
    W.ordv = ordv.data(); W.ordv2 = ordv2.data(); W.surv = surv.data(); W.thr = thr.data();
    W.cap = cap; W.node_cap = node_cap;
    if (M) { memcpy(kxy1.data(), xy, (size_t)M * 4); memcpy(kr1.data(), resp, M); }
    OtShared sh;
    memset(&sh, 0, sizeof(sh));
    return ot_run(W, &sh, M, nIni, hX, height, N, out_xy, out_resp, out_cap);
}

// The pass-free construction (og_octree2.cuh).  budget = cells of the deepest histogram level.  Returns -1 when the
// construction asks for the general path (ogm_octree).
extern "C" int ogm_octree_direct(const uint32_t* xy, const uint8_t* resp, int M, int width, int height, int N, int budget, int kcap,
                                 uint32_t* out_xy, uint8_t* out_resp, int out_cap) {
    using namespace og;
    const int nIni = (int)roundf((float)width / (float)height);
    if (nIni < 1) return -2;
    const float hX = (float)width / (float)nIni;
    const int node_cap = std::max(N + 3, 4 * nIni) + 8;
    const int Dh = ot2_depth(nIni, budget);
    const int small_cap = std::max(node_cap, OG_NTHREADS());
    std::vector<uint8_t> mem(ot2_smem_bytes(nIni, Dh, small_cap, kcap) + 16);
    uint8_t* base = mem.data() + ((16 - ((uintptr_t)mem.data() & 15)) & 15);
    OtShared sh;
    Ot2Shared s2;
    memset(&sh, 0, sizeof(sh));
    memset(&s2, 0, sizeof(s2));
    return ot_run_direct(Ot2CompactKeys{xy, resp}, M, base, Dh, small_cap, kcap, &sh, &s2, nIni, hX, height, N, out_xy, out_resp, out_cap);
}

// The same through the product's key source: the keys sit in per-cell slot runs (counts[c] keys in cell c, gaps between the
// runs), addressed by emission index through the cell offsets, with the first kcap path codes cached.
extern "C" int ogm_octree_direct_cells(const uint32_t* xy, const uint8_t* resp, int M, const int32_t* counts, int n_cells, int width,
                                       int height, int N, int budget, int kcap, uint32_t* out_xy, uint8_t* out_resp, int out_cap) {
    using namespace og;
    const int nIni = (int)roundf((float)width / (float)height);
    if (nIni < 1) return -2;
    const float hX = (float)width / (float)nIni;
    const int node_cap = std::max(N + 3, 4 * nIni) + 8;
    const int Dh = ot2_depth(nIni, budget);
    const int small_cap = std::max(node_cap, OG_NTHREADS());
    std::vector<Cell> cells(n_cells);
    std::vector<int32_t> coff(n_cells + 1, 0);
    int slot = 0, k = 0;
    std::vector<uint32_t> cxy;
    std::vector<uint8_t> crr;
    for (int c = 0; c < n_cells; ++c) {
        memset(&cells[c], 0, sizeof(Cell));
        cells[c].slot = slot;
        coff[c] = k;
        for (int j = 0; j < counts[c]; ++j, ++k) { cxy.push_back(xy[k]); crr.push_back(resp[k]); }
        for (int j = 0; j < 3; ++j) { cxy.push_back(0xdeadbeefu); crr.push_back(0); }   // unused slots of the cell
        slot += counts[c] + 3;
    }
    if (k != M) return -3;
    std::vector<uint8_t> mem(ot2_smem_bytes(nIni, Dh, small_cap, kcap) + 16);
    uint8_t* base = mem.data() + ((16 - ((uintptr_t)mem.data() & 15)) & 15);
    OtShared sh;
    Ot2Shared s2;
    memset(&sh, 0, sizeof(sh));
    memset(&s2, 0, sizeof(s2));
    const Ot2CellKeys keys{coff.data(), cells.data(), cxy.data(), crr.data(), n_cells};
    return ot_run_direct(keys, M, base, Dh, small_cap, kcap, &sh, &s2, nIni, hX, height, N, out_xy, out_resp, out_cap);
}

// og_capi.cu — the extern "C" layer of liborbgpu.so (include/orbgpu.h): extractor handle, workspace layout in
// HBM, launch sequence.  Host-side geometry follows the reference constructor and ComputePyramid /
// ComputeKeyPointsOctTree line by line (cited inline); no pixel is ever touched on the host.
#include "og_nvtx.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/orbgpu.h"
#include "og_extract.cu"
#include "og_stereo.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define OG_CUDA(expr)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(ORBGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));              \
    } while (0)

inline int cv_round_f(float v) { return (int)lrintf(v); }
inline int cv_round_d(double v) { return (int)lrint(v); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

int og_fail(int code, const std::string& msg) { return fail(code, msg); }  // shared with og_match.cu

struct orbgpu_extractor {
    int device = 0;
    int nfeatures = 0, nlevels = 0, ini_th = 0, min_th = 0;
    double scale_factor = 1.2;  // ORBextractor.h:98 keeps the float argument in a double member
    int max_w = 0, max_h = 0, max_batch = 0;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota, umax;
    int kp_cap = 0;

    cudaStream_t stream = nullptr;
    // host-pointer batch path: copies run on their own streams so that H2D of chunk k+1 and D2H of chunk k-1 overlap
    // the kernels of chunk k
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, stream2 = nullptr;   // stream2: odd chunks, so neighbouring chunks' kernels interleave
    std::vector<cudaEvent_t> ev_in, ev_out;
    cudaEvent_t ev_begin = nullptr;
    // the blur depends only on the pyramid: it runs on an auxiliary stream next to FAST + octree (the octree is latency
    // bound and leaves issue slots free); one auxiliary stream and event pair per compute stream
    static constexpr int kStreams = 8;                 // compute streams of the chunked host path: stream, stream2, extra[0..5]
    cudaStream_t s_extra[kStreams - 2] = {};
    cudaStream_t s_aux[kStreams] = {};
    cudaEvent_t ev_pyr[kStreams] = {}, ev_blur[kStreams] = {}, ev_fast[kStreams] = {};
    cudaStream_t compute_stream(int k) const { return k == 0 ? stream : (k == 1 ? stream2 : s_extra[k - 2]); }
    // geometry is rebuilt whenever the frame size changes (buffers are sized for max_w x max_h)
    int cur_w = 0, cur_h = 0;
    og::ExtractParams P;
    std::vector<og::Cell> cells;
    // device buffers (capacity fixed at creation)
    uint8_t *d_pyr = nullptr, *d_blur = nullptr, *d_images = nullptr, *d_ot = nullptr;
    uint8_t* d_color = nullptr;     // interleaved colour input of orbgpu_extract_batch_color (allocated on first use)
    size_t color_cap = 0;
    og::Cell* d_cells = nullptr;
    og::Segment* d_segs = nullptr;
    bool resize_tma[og::kMaxLevels] = {};   // per level: every 192 x 32 output tile reads at most a 256 x 42 source box (k_resize_tma)
    bool resize_mlp[og::kMaxLevels] = {};   // per level: every band of 8 output rows reads <= kResizeSpan source rows
    og::BlurTile* d_btiles = nullptr;
    uint32_t* d_ic_tab = nullptr;
    int n_btiles = 0;
    CUtensorMap* d_tmaps = nullptr;   // [5][kMaxLevels]: FAST tile boxes, blur input boxes, IC_Angle boxes (all over pyr), descriptor boxes over blur, resize source boxes over pyr
    int fast_smem = 0;
    bool frame_pending = false;   // the levels of the last call have no reflect-101 frame yet (written on demand)
    bool eager_frame = false;     // orbgpu_extractor_set_eager_frame: write the frame with every call, like the reference
    int oct_direct_smem = 0;      // shared memory of the pass-free octree (max over levels, for 256 / 512 threads: [0] / [1])
    int oct_direct_smem_lat = 0;
    int oct_kcap = 0, oct_kcap_lat = 0;   // keys whose path codes the pass-free octree caches in shared memory
    int last_octree_direct = 0;   // orbgpu_octree: 1 when the pass-free construction produced the last result
    og::Tap* d_taps = nullptr;
    int32_t *d_cell_count = nullptr, *d_sel_count = nullptr, *d_counts = nullptr;
    uint32_t *d_cand_xy = nullptr, *d_sel_xy = nullptr;
    uint8_t *d_cand_resp = nullptr, *d_sel_resp = nullptr;
    og::KeyPoint* d_kp = nullptr;
    uint8_t* d_desc = nullptr;
    size_t cap_pyr = 0, cap_cells = 0, cap_taps = 0, cap_cand = 0, cap_sel = 0, cap_ot = 0, cap_cellcount = 0;
    int last_batch = 0, last_launches = 0;
    // results of the last call as they sit on the device (consumed by orbgpu_stereo_matches)
    const og::KeyPoint* last_kp = nullptr;
    const uint8_t* last_desc = nullptr;
    const int32_t* last_counts = nullptr;
    int last_stride = 0;
    // stereo scratch (grow-only) and the event that orders the partner extractor's stream before ours
    int32_t *d_st_rows = nullptr, *d_st_items = nullptr, *d_st_sad = nullptr;
    float *d_st_u = nullptr, *d_st_d = nullptr;
    size_t st_cap_rows = 0, st_cap_items = 0, st_cap_kp = 0;
    cudaEvent_t ev_peer = nullptr;
    // optional per-stage timing (cudaEvents on the launching stream)
    bool profiling = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

struct Geometry {
    og::ExtractParams P;
    std::vector<og::Cell> cells;
    std::vector<og::Segment> segs;
    std::vector<og::BlurTile> btiles;
    std::vector<og::Tap> taps;            // all levels' x then y tables, concatenated
    std::vector<size_t> xt_off, yt_off, ytw_off;   // offsets into taps (ytw: decoded vertical taps, 16 bytes = two Tap slots each)
    size_t pyr_bytes_per_frame = 0;
};

// cv::resize coefficient table for one axis.  Horizontal: fraction zeroed when the first
// tap is clamped; vertical: weights kept, row indices clipped.
void make_taps(int sn, int dn, bool horizontal, std::vector<og::Tap>& out) {
    const double sc = (double)sn / dn;
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * sc - 0.5);
        int s = (int)floorf(f);
        f -= s;
        og::Tap t;
        if (horizontal) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= sn - 1) { s = sn - 1; f = 0.f; }
            t.s0 = (int16_t)s;
            t.s1 = (int16_t)std::min(s + 1, sn - 1);
        } else {
            t.s0 = (int16_t)std::min(std::max(s, 0), sn - 1);
            t.s1 = (int16_t)std::min(std::max(s + 1, 0), sn - 1);
        }
        t.w0 = (int16_t)cv_round_f((1.f - f) * 2048.f);
        t.w1 = (int16_t)cv_round_f(f * 2048.f);
        out.push_back(t);
    }
}

size_t octree_ws_bytes(int cap, int node_cap) {
    size_t b = 0;
    auto take = [&](size_t bytes) { b += (bytes + 15) & ~size_t(15); };
    take((size_t)cap * 4); take((size_t)cap * 4);
    take((size_t)cap * 2); take((size_t)cap * 2);
    take((size_t)cap); take((size_t)cap);
    take((size_t)node_cap * sizeof(og::OtNode)); take((size_t)node_cap * sizeof(og::OtNode));
    take((size_t)node_cap * sizeof(og::OtTmp));
    for (int i = 0; i < 5; ++i) take((size_t)node_cap * 4);
    take((size_t)og::kOctMaxThreads * 16);
    return b;
}

// The constructor's tables (ORBextractor.cc:413-469), pure host arithmetic in the reference's own types: the float
// scale argument lives in a double member (ORBextractor.h:98), every table entry is a float.
void compute_tables(int nfeatures, float scale_factor_f, int nlevels, std::vector<float>& scale, std::vector<float>& inv_scale,
                    std::vector<float>& sigma2, std::vector<float>& inv_sigma2, std::vector<int>& quota, std::vector<int>& umax) {
    const double scale_factor = scale_factor_f;
    scale.resize(nlevels); sigma2.resize(nlevels); inv_scale.resize(nlevels); inv_sigma2.resize(nlevels);
    scale[0] = 1.0f; sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        scale[i] = (float)(scale[i - 1] * scale_factor);
        sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < nlevels; i++) { inv_scale[i] = 1.0f / scale[i]; inv_sigma2[i] = 1.0f / sigma2[i]; }
    // :435-446
    quota.resize(nlevels);
    float factor = (float)(1.0f / scale_factor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        quota[l] = cv_round_f(nDesired);
        sum += quota[l];
        nDesired *= factor;
    }
    quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // :454-469
    umax.assign(16, 0);
    const int HP = og::kHalfPatch;
    int v, v0, vmax = (int)floor(HP * sqrt(2.f) / 2 + 1);
    int vmin = (int)ceil(HP * sqrt(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// u8 tensor [batch][rows][pitch] at `base`, box = box_w x box_h x 1 (zero fill outside the tensor)
std::string make_level_tmap(CUtensorMap* out, uint8_t* base, int pitch, int rows, long long frame_stride, int batch, int box_w, int box_h,
                            CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return "cuTensorMapEncodeTiled is not available from this driver";
    const cuuint64_t gdim[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t est[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return "";
}

// Everything that depends on the frame size.  Returns an error string or "".
std::string build_geometry(const orbgpu_extractor& ex, int w, int h, int batch_cap, Geometry& G) {
    og::ExtractParams& P = G.P;
    memset(&P, 0, sizeof(P));
    P.n_levels = ex.nlevels;
    P.ini_th = ex.ini_th;
    P.min_th = ex.min_th;
    for (int i = 0; i < 16; ++i) P.umax[i] = ex.umax[i];
    // cv::fastAtan2 constants, evaluated in float like OpenCV's static initialisers (Appendix A.4)
    const float sc = (float)(180.0 / 3.1415926535897932384626433832795);
    P.atan.p1 = 0.9997878412794807f * sc;
    P.atan.p3 = -0.3258083974640975f * sc;
    P.atan.p5 = 0.1555786518463281f * sc;
    P.atan.p7 = -0.04432655554792128f * sc;
    P.atan.eps = (float)DBL_EPSILON;
    P.factor_pi = (float)(3.1415926535897932384626433832795 / 180.f);  // ORBextractor.cc:107

    size_t pyr_off = 0;
    int cell_base = 0, cand_base = 0, sel_base = 0;
    size_t ot_off = 0;
    for (int l = 0; l < ex.nlevels; ++l) {
        og::Level& L = P.lv[l];
        L.w = cv_round_f((float)w * ex.inv_scale[l]);  // :1112
        L.h = cv_round_f((float)h * ex.inv_scale[l]);
        if (L.w < 62 || L.h < 62)  // maxBorder-minBorder = size-32 must hold one 30 px cell (:781-784)
            return "image too small: pyramid level " + std::to_string(l) + " is " + std::to_string(L.w) + "x" +
                   std::to_string(L.h) + " (every level needs a 30 px cell inside the 16 px border)";
        L.pitch = (int)align_up((size_t)L.w + 2 * og::kXPad, 128);
        L.rows = L.h + 2 * og::kEdge;
        L.frame_stride = (long long)L.pitch * L.rows;
        L.base = (long long)pyr_off;
        pyr_off += (size_t)L.frame_stride * batch_cap;
        if (l > 0) {
            while (G.taps.size() % 4) G.taps.push_back(og::Tap{0, 0, 0, 0});   // 32-byte aligned: k_resize4 loads 4 taps as 2 x uint4
            G.xt_off.push_back(G.taps.size());
            make_taps(P.lv[l - 1].w, L.w, true, G.taps);
            while (G.taps.size() % 4) G.taps.push_back(G.taps.back());          // padded outputs repeat the last column
            G.yt_off.push_back(G.taps.size());
            make_taps(P.lv[l - 1].h, L.h, false, G.taps);
            while (G.taps.size() % 2) G.taps.push_back(og::Tap{0, 0, 0, 0});   // 16-byte aligned
            G.ytw_off.push_back(G.taps.size());
            for (int y = 0; y < L.h; ++y) {
                const og::Tap t = G.taps[G.yt_off.back() + y];
                const uint32_t w[4] = {(uint32_t)t.s0 * 256u, (uint32_t)t.s1 * 256u, (uint32_t)(uint16_t)t.w0 << 16, (uint32_t)(uint16_t)t.w1 << 16};
                og::Tap two[2];
                memcpy(two, w, 16);
                G.taps.push_back(two[0]);
                G.taps.push_back(two[1]);
            }
        } else {
            G.xt_off.push_back(0);
            G.yt_off.push_back(0);
            G.ytw_off.push_back(0);
        }
        // detection grid, ORBextractor.cc:770-806
        const int minBX = og::kEdge - 3, minBY = minBX;
        const int maxBX = L.w - og::kEdge + 3, maxBY = L.h - og::kEdge + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        L.cell_base = cell_base;
        L.cand_base = cand_base;
        int slot = 0, ncell = 0;
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int tw = (int)maxX - (int)iniX - 6, th = (int)maxY - (int)iniY - 6;
                if (tw <= 0 || th <= 0) continue;  // cv::FAST tests nothing on a sub-image thinner than 7 px
                if (tw > og::kCellMax || th > og::kCellMax) return "internal: cell larger than kCellMax";
                og::Cell c;
                c.level = (int16_t)l;
                c.x0 = (int16_t)((int)iniX + 3);
                c.y0 = (int16_t)((int)iniY + 3);
                c.tw = (int16_t)tw;
                c.th = (int16_t)th;
                c.pad = 0;
                c.slot = slot;
                // 3x3 strict maxima cannot be 8-neighbours: at most one per 2x2 block
                slot += ((tw + 1) / 2) * ((th + 1) / 2);
                G.cells.push_back(c);
                ++ncell;
            }
        }
        L.n_cells = ncell;
        L.wcell = wCell;
        L.wcell_magic = (65536 + wCell - 1) / wCell;
        // segments: runs of consecutive cells of one cell row that fit one 256-byte tile (k_fast_seg)
        {
            int th_max = 0;
            for (int k = 0; k < ncell;) {
                const og::Cell& c0 = G.cells[cell_base + k];
                og::Segment sg;
                sg.level = (int16_t)l;
                sg.first_cell = cell_base + k;
                sg.x0 = c0.x0; sg.y0 = c0.y0; sg.th = c0.th;
                int tw = 0, n = 0;
                while (k + n < ncell) {
                    const og::Cell& c = G.cells[cell_base + k + n];
                    if (c.y0 != c0.y0 || c.x0 != c0.x0 + n * wCell || tw + c.tw > og::kSegMaxTw) break;
                    // every cell but the last of a row is exactly wCell wide, which the kernel's column table relies on
                    if (n > 0 && G.cells[cell_base + k + n - 1].tw != wCell) break;
                    tw += c.tw;
                    ++n;
                }
                sg.ncells = (int16_t)n;
                sg.tw = (int16_t)tw;
                const int gx = (og::kXPad + sg.x0 - 3) & ~15, ox = og::kXPad + sg.x0 - gx;   // as in k_fast_seg
                const uint32_t nw = (uint32_t)(((ox + tw - 1) >> 3) - (ox >> 3) + 1);   // 8-pixel steps covering the tested columns
                sg.nw_magic = (uint32_t)((0x100000000ull + nw - 1) / nw);
                G.segs.push_back(sg);
                th_max = std::max(th_max, (int)c0.th);
                k += n;
            }
            L.hbox = th_max + 6;
        }
        L.cand_cap = std::max(slot, 1);
        cell_base += ncell;
        cand_base += L.cand_cap;
        L.quota = ex.quota[l];
        L.det_w = maxBX - minBX;
        L.det_h = maxBY - minBY;
        L.n_ini = (int)roundf((float)L.det_w / (float)L.det_h);  // :543
        if (L.n_ini < 1)
            return "unsupported aspect ratio: level " + std::to_string(l) + " is taller than 2:1 (the reference divides by zero here)";
        L.hx = (float)L.det_w / (float)L.n_ini;  // :545
        L.sel_cap = std::max(L.quota + 3, 4 * L.n_ini);
        L.node_cap = L.sel_cap + 8;
        if (L.node_cap > 65535) return "nfeatures too large for the 16-bit node index";
        L.sel_base = sel_base;
        sel_base += L.sel_cap;
        L.scale = ex.scale[l];
        L.kp_size = (float)(int)(31 * ex.scale[l]);  // :837 (PATCH_SIZE*mvScaleFactor -> int)
        for (int ty = 0; ty < (L.h + og::kBlurH - 1) / og::kBlurH; ++ty)
            for (int tx = 0; tx < (L.w + og::kBlurW - 1) / og::kBlurW; ++tx) G.btiles.push_back({(int16_t)l, (int16_t)tx, (int16_t)ty, 0});
        L.ot_base = (long long)ot_off;
        ot_off += align_up(octree_ws_bytes(L.cand_cap, L.node_cap), 256);
    }
    P.total_cells = cell_base;
    P.total_cand_cap = cand_base;
    P.total_sel_cap = sel_base;
    P.ot_frame_bytes = (long long)ot_off;
    P.kp_cap = sel_base;
    G.pyr_bytes_per_frame = pyr_off / batch_cap;
    return "";
}

int ensure_geometry(orbgpu_extractor* ex, int w, int h) {
    if (ex->cur_w == w && ex->cur_h == h) return ORBGPU_OK;
    if (w > ex->max_w || h > ex->max_h)
        return fail(ORBGPU_ERR_ARG, "image " + std::to_string(w) + "x" + std::to_string(h) + " exceeds the extractor's max size");
    Geometry G;
    std::string err = build_geometry(*ex, w, h, ex->max_batch, G);
    if (!err.empty()) return fail(ORBGPU_ERR_ARG, err);
    const og::ExtractParams& P = G.P;
    const size_t B = (size_t)ex->max_batch;
    if ((size_t)P.lv[ex->nlevels - 1].base + (size_t)P.lv[ex->nlevels - 1].frame_stride * B > ex->cap_pyr ||
        G.cells.size() > ex->cap_cells || G.taps.size() > ex->cap_taps || (size_t)P.total_cand_cap > ex->cap_cand ||
        (size_t)P.total_sel_cap > ex->cap_sel || (size_t)P.ot_frame_bytes > ex->cap_ot)
        return fail(ORBGPU_ERR_CAPACITY, "internal: workspace sized at creation is too small for this frame size");
    OG_CUDA(cudaMemcpyAsync(ex->d_cells, G.cells.data(), G.cells.size() * sizeof(og::Cell), cudaMemcpyHostToDevice, ex->stream));
    OG_CUDA(cudaMemcpyAsync(ex->d_segs, G.segs.data(), G.segs.size() * sizeof(og::Segment), cudaMemcpyHostToDevice, ex->stream));
    if (G.btiles.size() > ex->cap_cells) return fail(ORBGPU_ERR_CAPACITY, "internal: blur tile table too small");
    OG_CUDA(cudaMemcpyAsync(ex->d_btiles, G.btiles.data(), G.btiles.size() * sizeof(og::BlurTile), cudaMemcpyHostToDevice, ex->stream));
    ex->n_btiles = (int)G.btiles.size();
    {
        std::vector<CUtensorMap> maps(5 * og::kMaxLevels);
        memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
        int smem = 0;
        for (int l = 0; l < ex->nlevels; ++l) {
            const og::Level& L = P.lv[l];
            std::string e = make_level_tmap(&maps[l], ex->d_pyr + L.base, L.pitch, L.rows, L.frame_stride, ex->max_batch, og::kSegPitch, L.hbox);
            if (e.empty())
                e = make_level_tmap(&maps[og::kMaxLevels + l], ex->d_pyr + L.base, L.pitch, L.rows, L.frame_stride, ex->max_batch, 256, og::kBlurBox);
            if (e.empty())
                e = make_level_tmap(&maps[2 * og::kMaxLevels + l], ex->d_pyr + L.base, L.pitch, L.rows, L.frame_stride, ex->max_batch, og::kIcBoxW,
                                    og::kIcRows);
            if (e.empty())
                e = make_level_tmap(&maps[3 * og::kMaxLevels + l], ex->d_blur + L.base, L.pitch, L.rows, L.frame_stride, ex->max_batch, og::kBlurBoxW,
                                    og::kPatchRows, OG_DESC_SWIZZLE ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
            if (e.empty())   // source boxes of the resize into level l + 1
                e = make_level_tmap(&maps[4 * og::kMaxLevels + l], ex->d_pyr + L.base, L.pitch, L.rows, L.frame_stride, ex->max_batch, 256, og::kRtBoxH);
            if (!e.empty()) return fail(ORBGPU_ERR_CUDA, e);
            smem = std::max(smem, og::fast_seg_smem_bytes(L.hbox, L.hbox - 6));
        }
        OG_CUDA(cudaMemcpyAsync(ex->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, ex->stream));
        OG_CUDA(cudaStreamSynchronize(ex->stream));   // `maps` is a local
        if (smem > 200 * 1024) return fail(ORBGPU_ERR_ARG, "internal: FAST tile does not fit shared memory");
        // the attribute is per function, not per handle: always allow the largest tile any geometry can ask for
        OG_CUDA(cudaFuncSetAttribute(og::k_octree<og::kOctLatThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        OG_CUDA(cudaFuncSetAttribute(og::k_octree<og::kOctThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, og::kOctSmem));
        OG_CUDA(cudaFuncSetAttribute(og::k_octree<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, og::kOctSmem));
        OG_CUDA(cudaFuncSetAttribute(og::k_fast_seg, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     og::fast_seg_smem_bytes(og::kCellMax + 6, og::kCellMax)));
        ex->fast_smem = smem;
        // shared memory of the pass-free octree: histogram + node arrays (max over levels) + cell offsets + a cache of path codes
        // for as many keys as fit into kOctDirectSmem (3 bytes each)
        int od = 0, odl = 0, cells_max = 0;
        for (int l = 0; l < ex->nlevels; ++l) {
            const og::Level& L = P.lv[l];
            const int Dh = og::ot2_depth(L.n_ini, og::kOt2Budget);
            od = std::max(od, (int)og::ot2_smem_bytes(L.n_ini, Dh, std::max(L.node_cap, 256), 0));
            odl = std::max(odl, (int)og::ot2_smem_bytes(L.n_ini, Dh, std::max(L.node_cap, og::kOctLatThreads), 0));
            cells_max = std::max(cells_max, L.n_cells);
        }
        const int coff_bytes = 4 * cells_max + 16;
        auto kcap_for = [&](int base, int budget) { return std::max(0, std::min(16384, (budget - base - coff_bytes) / 3)) & ~15; };
        ex->oct_kcap = kcap_for(od, og::kOctDirectSmem);
        ex->oct_kcap_lat = kcap_for(odl, 96 * 1024);
        od += 3 * ex->oct_kcap + coff_bytes;
        odl += 3 * ex->oct_kcap_lat + coff_bytes;
        ex->oct_direct_smem = od <= 100 * 1024 ? od : 0;        // larger than that (huge nfeatures): the general path alone
        ex->oct_direct_smem_lat = odl <= 100 * 1024 ? odl : 0;
    }
    if (!G.taps.empty())
        OG_CUDA(cudaMemcpyAsync(ex->d_taps, G.taps.data(), G.taps.size() * sizeof(og::Tap), cudaMemcpyHostToDevice, ex->stream));
    OG_CUDA(cudaStreamSynchronize(ex->stream));
    ex->P = P;
    for (int l = 1; l < ex->nlevels; ++l) {
        ex->P.lv[l].xt = ex->d_taps + G.xt_off[l];
        ex->P.lv[l].yt = ex->d_taps + G.yt_off[l];
        ex->P.lv[l].ytw = reinterpret_cast<const uint4*>(ex->d_taps + G.ytw_off[l]);
        const og::Tap* yt = G.taps.data() + G.yt_off[l];
        bool ok = true;
        for (int y0 = 0; y0 < P.lv[l].h && ok; y0 += og::kResizeRows) {
            const int y1 = std::min(y0 + og::kResizeRows, P.lv[l].h) - 1;
            ok = yt[y1].s1 - yt[y0].s0 + 1 <= og::kResizeSpan;
        }
        ex->resize_mlp[l] = ok;
        // k_resize_tma: the source bytes of every 192 x 32 output tile must lie inside one 256 x kRtBoxH box whose first column is
        // 16-byte aligned in the level buffer
        const og::Tap* xt = G.taps.data() + G.xt_off[l];
        bool fits = true;
        for (int X0 = 0; X0 < P.lv[l].w && fits; X0 += og::kRtW) {
            const int X1 = std::min(X0 + og::kRtW, P.lv[l].w) - 1, gx = (og::kXPad + xt[X0].s0) & ~15;
            fits = og::kXPad + xt[X1].s1 - gx < 256;   // (words read past a box row's end belong to the next row and are never selected)
        }
        for (int Y0 = 0; Y0 < P.lv[l].h && fits; Y0 += og::kRtH) {
            const int Y1 = std::min(Y0 + og::kRtH, P.lv[l].h) - 1;
            fits = yt[Y1].s1 - yt[Y0].s0 + 1 <= og::kRtBoxH;
        }
        ex->resize_tma[l] = fits;
    }
    ex->P.pyr = ex->d_pyr;
    ex->P.blur = ex->d_blur;
    ex->P.cells = ex->d_cells;
    ex->P.ic_tab = ex->d_ic_tab;
    ex->P.segs = ex->d_segs;
    ex->P.n_segs = (int)G.segs.size();
    ex->P.cell_count = ex->d_cell_count;
    ex->P.cand_xy = ex->d_cand_xy;
    ex->P.cand_resp = ex->d_cand_resp;
    ex->P.ot_ws = ex->d_ot;
    ex->P.sel_xy = ex->d_sel_xy;
    ex->P.sel_resp = ex->d_sel_resp;
    ex->P.sel_count = ex->d_sel_count;
    ex->cells = G.cells;
    ex->cur_w = w;
    ex->cur_h = h;
    return ORBGPU_OK;
}

// The launch sequence of one batch (device pointers in, device pointers out), all on ex->stream.
int launch_extract(orbgpu_extractor* ex, const uint8_t* d_images, int batch, size_t row_stride, size_t frame_stride,
                   og::KeyPoint* d_kp, uint8_t* d_desc, int kp_capacity, int32_t* d_counts, int frame0 = 0, cudaStream_t on = nullptr) {
    og::ExtractParams P = ex->P;
    P.batch = batch;
    P.frame0 = frame0;
    P.kp_cap = kp_capacity;
    cudaStream_t st = on ? on : ex->stream;
    int launches = 0;
    // stage boundaries: CUDA events when per-stage profiling is on (orbgpu_extractor_stage_ms), NVTX ranges on the host timeline always
    static const char* const kStageName[5] = {"pyramid", "fast_cells", "octree", "blur", "orient_desc"};
    auto mark = [&](int i) {
        if (ex->profiling) cudaEventRecord(ex->ev[i], st);
        if (i > 0) nvtxRangePop();
        if (i < 5) nvtxRangePushA(kStageName[i]);
    };
    mark(0);
    {
        const og::Level& L = P.lv[0];
        const int nvec = (L.w + 15) / 16, n_items = nvec * L.h;
        const uint32_t magic = (uint32_t)(((1ull << 32) + nvec - 1) / nvec);   // exact floor(id / nvec) for id < 2^32 / nvec
        og::k_level0<<<dim3((n_items + 255) / 256, batch), 256, 0, st>>>(P, d_images, (long long)row_stride, (long long)frame_stride, nvec, magic,
                                                                      n_items);
        ++launches;
    }
    bool generic = false;
    (void)generic;
    for (int l = 1; l < P.n_levels; ++l) {
        const og::Level& L = P.lv[l];
        // the vectorised kernel needs the 4 outputs of a thread inside 12 source bytes: scale <= 2
        if ((double)P.lv[l - 1].w / L.w <= 2.0) {
            const int nwx = (L.w + 3) / 4, bands = (L.h + og::kResizeRows - 1) / og::kResizeRows, n_items = nwx * bands;
            const uint32_t magic = (uint32_t)((0x100000000ull + nwx - 1) / nwx);
            // ORBGPU_RESIZE_TMA=0: the register-staged kernel for every level (development; measured within 1 % of each other)
            static const int tma_env = []() { const char* e = getenv("ORBGPU_RESIZE_TMA"); return e ? atoi(e) : 1; }();
            if (tma_env && ex->resize_tma[l]) {
                const int ntx = (L.w + og::kRtW - 1) / og::kRtW, nty = (L.h + og::kRtH - 1) / og::kRtH;
                og::k_resize_tma<<<dim3(ntx * nty, batch), og::kRtThreads, 0, st>>>(P, l, ntx, ex->d_tmaps);
            } else if (ex->resize_mlp[l])
                og::k_resize4_pp<<<dim3((n_items + og::kResizeThreads - 1) / og::kResizeThreads, batch), og::kResizeThreads, 0, st>>>(P, l, nwx, magic, n_items);
            else
                og::k_resize4<<<dim3((n_items + og::kResizeThreads - 1) / og::kResizeThreads, batch), og::kResizeThreads, 0, st>>>(P, l, nwx, magic, n_items);
        } else {
            dim3 grid((L.pitch / 4 + 127) / 128, L.rows, batch);
            og::k_resize<<<grid, 128, 0, st>>>(P, l);
            generic = true;
        }
        ++launches;
    }
    // per-stage profiling keeps everything on one stream so that the stage times add up
    int aux = 0;
    for (int k = 1; k < orbgpu_extractor::kStreams; ++k)
        if (st == ex->compute_stream(k)) aux = k;
    // The 19-px reflect-101 frame around every level (:1122-1128) is memory behind mvImagePyramid that nothing on this path reads:
    // FAST cells start 16 px inside a level, the orientation and descriptor windows stay inside it, the blur mirrors its own
    // 3-px halo.  It is therefore materialised ON DEMAND (orbgpu_extractor_read_level(bordered), the shell's mvImagePyramid
    // download) and costs the extraction nothing.  ORBGPU_EAGER_FRAME=1 writes it with every call (on the auxiliary stream
    // beside FAST, or in line when profiling) as before.
    static const int eager_frame_env = []() { const char* e = getenv("ORBGPU_EAGER_FRAME"); return e ? atoi(e) : 0; }();
    const bool eager_frame = eager_frame_env != 0 || ex->eager_frame;
    const bool async_frame = eager_frame && !ex->profiling;
    auto borders = [&](cudaStream_t s) {
        // the generic resize writes its own frame; the border kernels then only repeat it (and do level 0)
        const int sides0 = 2 * P.lv[0].h;
        og::k_border_sides<<<dim3((sides0 + og::kBorderThreads - 1) / og::kBorderThreads, P.n_levels, batch), og::kBorderThreads, 0, s>>>(P);
        og::k_border_caps<<<dim3(2 * og::kEdge, P.n_levels, batch), og::kBorderThreads, 0, s>>>(P);
        launches += 2;
    };
    if (eager_frame && !async_frame) borders(st);
    if (!eager_frame) ex->frame_pending = true;
    mark(1);
    // measured on B200: pays off while a launch cannot fill the GPU (53.5k vs 50.0k frames/s at 64 frames), costs 8 % at 1024
    const bool overlap_blur = !ex->profiling && batch <= 128;
    // larger batches: the blur (issue bound) runs beside the octree instead of FAST
    const bool blur_with_octree = !ex->profiling && !overlap_blur;
    if (async_frame || overlap_blur) {
        OG_CUDA(cudaEventRecord(ex->ev_pyr[aux], st));
        OG_CUDA(cudaStreamWaitEvent(ex->s_aux[aux], ex->ev_pyr[aux], 0));
        if (async_frame) borders(ex->s_aux[aux]);
    }
    if (overlap_blur) {
        og::k_blur_tma<<<dim3(ex->n_btiles, batch), og::kBlurThreads, 0, ex->s_aux[aux]>>>(P, ex->d_btiles, ex->d_tmaps);
        OG_CUDA(cudaEventRecord(ex->ev_blur[aux], ex->s_aux[aux]));
        ++launches;
    }
    og::k_fast_seg<<<dim3(P.n_segs, batch), og::kSegThreads, ex->fast_smem, st>>>(P, ex->d_tmaps);
    ++launches;
    mark(2);
    if (blur_with_octree) {
        OG_CUDA(cudaEventRecord(ex->ev_fast[aux], st));
        OG_CUDA(cudaStreamWaitEvent(ex->s_aux[aux], ex->ev_fast[aux], 0));
        og::k_blur_tma<<<dim3(ex->n_btiles, batch), og::kBlurThreads, 0, ex->s_aux[aux]>>>(P, ex->d_btiles, ex->d_tmaps);
        OG_CUDA(cudaEventRecord(ex->ev_blur[aux], ex->s_aux[aux]));
        ++launches;
    }
    {
        // tuning hooks (development): ORBGPU_OCT_SMEM = bytes of shared-memory workspace per CTA for large batches (0 = HBM only),
        // ORBGPU_OCT_THREADS = 128 | 256 | 512
        static const int oct_smem_env = []() { const char* e = getenv("ORBGPU_OCT_SMEM"); return e ? atoi(e) : -1; }();
        static const int oct_thr_env = []() { const char* e = getenv("ORBGPU_OCT_THREADS"); return e ? atoi(e) : 0; }();
        static const int direct_env = []() { const char* e = getenv("ORBGPU_OCT_DIRECT"); return e ? atoi(e) : 1; }();
        if (batch <= og::kOctSmemMaxBatch) {
            const int db = direct_env ? ex->oct_direct_smem_lat : 0;
            const int budget = std::min(og::kOctSmem, 226 * 1024 - db) & ~15;
            og::k_octree<og::kOctLatThreads><<<dim3(P.n_levels, batch), og::kOctLatThreads, db + budget, st>>>(P, budget, db, ex->oct_kcap_lat);
        } else {
            const int sm = oct_smem_env >= 0 ? oct_smem_env : 0;
            const int thr = oct_thr_env ? oct_thr_env : 256;   // measured at 1024 frames (division passes): 128 threads 1.62 ms, 256 1.56 ms, 512 2.19 ms
            const int db = direct_env ? (thr == 512 ? ex->oct_direct_smem_lat : ex->oct_direct_smem) : 0;
            const int kc = thr == 512 ? ex->oct_kcap_lat : ex->oct_kcap;
            // ORBGPU_OCT_PAD: unused extra shared memory per CTA (caps the resident octree CTAs so that blur CTAs fit beside them)
            static const int pad = []() { const char* e = getenv("ORBGPU_OCT_PAD"); return e ? atoi(e) : 0; }();
            if (thr == 512) og::k_octree<512><<<dim3(P.n_levels, batch), 512, db + sm + pad, st>>>(P, sm, db, kc);
            else if (thr == 256) og::k_octree<256><<<dim3(P.n_levels, batch), 256, db + sm + pad, st>>>(P, sm, db, kc);
            else og::k_octree<og::kOctThreads><<<dim3(P.n_levels, batch), og::kOctThreads, db + sm + pad, st>>>(P, sm, db, kc);
        }
    }
    ++launches;
    mark(3);
    if (overlap_blur || blur_with_octree) {
        OG_CUDA(cudaStreamWaitEvent(st, ex->ev_blur[aux], 0));
    } else {
        og::k_blur_tma<<<dim3(ex->n_btiles, batch), og::kBlurThreads, 0, st>>>(P, ex->d_btiles, ex->d_tmaps);
        ++launches;
    }
    mark(4);
    og::k_orient_desc<<<dim3((ex->kp_cap + og::kDescWarps * og::kDescPerWarp - 1) / (og::kDescWarps * og::kDescPerWarp), batch), og::kDescWarps * 32, 0, st>>>(
        P, ex->d_tmaps, d_kp, d_desc, d_counts);
    ++launches;
    mark(5);
    OG_CUDA(cudaGetLastError());
    ex->last_batch = batch;
    ex->last_launches = launches;
    ex->last_kp = d_kp;
    ex->last_desc = d_desc;
    ex->last_counts = d_counts;
    ex->last_stride = kp_capacity;
    return ORBGPU_OK;
}

int check_call(orbgpu_extractor* ex, int batch, int width, int height, size_t row_stride, int kp_capacity) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    if (batch < 1 || batch > ex->max_batch)
        return fail(ORBGPU_ERR_ARG, "batch " + std::to_string(batch) + " outside [1, max_batch=" + std::to_string(ex->max_batch) + "]");
    if (row_stride < (size_t)width) return fail(ORBGPU_ERR_ARG, "row_stride smaller than width");
    if (kp_capacity < ex->kp_cap)
        return fail(ORBGPU_ERR_CAPACITY, "kp_capacity " + std::to_string(kp_capacity) + " < orbgpu_extractor_max_keypoints() = " + std::to_string(ex->kp_cap));
    OG_CUDA(cudaSetDevice(ex->device));
    return ensure_geometry(ex, width, height);
}

}  // namespace

extern "C" {

const char* orbgpu_last_error(void) { return g_err.c_str(); }
int orbgpu_abi_version(void) { return 2; }

// process-wide default device of the C++ shells (ORBextractor::SetDevice): one selector for extraction, matching and vocabulary
static std::atomic<int> g_default_device{0};
int orbgpu_set_default_device(int device) {
    int n = 0;
    OG_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(ORBGPU_ERR_ARG, "device index out of range");
    g_default_device.store(device);
    return ORBGPU_OK;
}
int orbgpu_default_device(void) { return g_default_device.load(); }

int orbgpu_device_count(int* count) {
    if (!count) return fail(ORBGPU_ERR_ARG, "null count");
    *count = 0;
    OG_CUDA(cudaGetDeviceCount(count));
    return ORBGPU_OK;
}

int orbgpu_extractor_create(orbgpu_extractor** out, int device, int nfeatures, float scale_factor, int nlevels,
                            int ini_th_fast, int min_th_fast, int max_width, int max_height, int max_batch) {
    if (!out) return fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    if (nfeatures < 1 || nlevels < 1 || nlevels > og::kMaxLevels || !(scale_factor > 1.f) || max_batch < 1 ||
        ini_th_fast < 1 || min_th_fast < 1 || min_th_fast > ini_th_fast || ini_th_fast > 254)
        return fail(ORBGPU_ERR_ARG, "bad extractor parameters (need nfeatures>=1, 1<=nlevels<=12, scaleFactor>1, 1<=minTh<=iniTh<=254)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ORBGPU_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ORBGPU_ERR_ARG, "device index out of range");
    OG_CUDA(cudaSetDevice(device));

    orbgpu_extractor* ex = new orbgpu_extractor();
    ex->device = device;
    ex->nfeatures = nfeatures;
    ex->nlevels = nlevels;
    ex->ini_th = ini_th_fast;
    ex->min_th = min_th_fast;
    ex->scale_factor = scale_factor;
    ex->max_w = max_width;
    ex->max_h = max_height;
    ex->max_batch = max_batch;
    compute_tables(nfeatures, scale_factor, nlevels, ex->scale, ex->inv_scale, ex->sigma2, ex->inv_sigma2, ex->quota, ex->umax);

    Geometry G;
    std::string err = build_geometry(*ex, max_width, max_height, max_batch, G);
    if (!err.empty()) { delete ex; return fail(ORBGPU_ERR_ARG, err); }
    ex->kp_cap = G.P.total_sel_cap;
    const size_t B = (size_t)max_batch;
    // capacities with slack: a smaller frame can need marginally more cells per pixel than the maximum one
    ex->cap_pyr = (size_t)G.pyr_bytes_per_frame * B + 4096;
    ex->cap_cells = G.cells.size() * 2 + 64;
    ex->cap_taps = G.taps.size() + 64;
    ex->cap_cand = (size_t)G.P.total_cand_cap * 5 / 4 + 1024;
    ex->cap_sel = (size_t)G.P.total_sel_cap + 64;
    ex->cap_ot = (size_t)G.P.ot_frame_bytes * 5 / 4 + 65536;
    ex->cap_cellcount = ex->cap_cells;
    cudaError_t ce = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (ce == cudaSuccess) ce = cudaMalloc(p, std::max<size_t>(bytes, 16)); };
    alloc((void**)&ex->d_pyr, ex->cap_pyr);
    alloc((void**)&ex->d_blur, ex->cap_pyr);
    alloc((void**)&ex->d_images, (size_t)max_width * max_height * B);
    alloc((void**)&ex->d_cells, ex->cap_cells * sizeof(og::Cell));
    alloc((void**)&ex->d_segs, ex->cap_cells * sizeof(og::Segment));
    alloc((void**)&ex->d_btiles, ex->cap_cells * sizeof(og::BlurTile));
    alloc((void**)&ex->d_tmaps, 5 * og::kMaxLevels * sizeof(CUtensorMap));
    alloc((void**)&ex->d_taps, ex->cap_taps * sizeof(og::Tap));
    alloc((void**)&ex->d_cell_count, ex->cap_cellcount * B * 4);
    alloc((void**)&ex->d_cand_xy, ex->cap_cand * B * 4);
    alloc((void**)&ex->d_cand_resp, ex->cap_cand * B);
    alloc((void**)&ex->d_ot, ex->cap_ot * B);
    alloc((void**)&ex->d_sel_xy, ex->cap_sel * B * 4);
    alloc((void**)&ex->d_sel_resp, ex->cap_sel * B);
    alloc((void**)&ex->d_sel_count, (size_t)og::kMaxLevels * B * 4);
    alloc((void**)&ex->d_counts, B * 4);
    alloc((void**)&ex->d_kp, (size_t)ex->kp_cap * B * sizeof(og::KeyPoint));
    alloc((void**)&ex->d_desc, (size_t)ex->kp_cap * B * 32);
    alloc((void**)&ex->d_ic_tab, 4 * og::kIcRows * og::kIcWords * 2 * sizeof(uint32_t));
    if (ce == cudaSuccess) {
        // IC_Angle weights: word j of disc row v (alignment A) covers columns u = -15 - A + 4j + k, k = 0..3; a column inside the
        // disc (|u| <= umax[|v|], :85-99) weighs u for m_10 and v for m_01, everything else 0
        std::vector<uint32_t> tab((size_t)4 * og::kIcRows * og::kIcWords * 2, 0u);
        for (int A = 0; A < 4; ++A)
            for (int r = 0; r < og::kIcRows; ++r)
                for (int j = 0; j < og::kIcWords; ++j) {
                    const int v = r - og::kHalfPatch;
                    uint32_t wu = 0, wv = 0;
                    for (int k = 0; k < 4; ++k) {
                        const int u = -og::kHalfPatch - A + 4 * j + k;
                        if (std::abs(u) <= ex->umax[std::abs(v)]) {
                            wu |= (uint32_t)(uint8_t)(int8_t)u << (8 * k);
                            wv |= (uint32_t)(uint8_t)(int8_t)v << (8 * k);
                        }
                    }
                    const size_t o = (((size_t)A * og::kIcRows + r) * og::kIcWords + j) * 2;
                    tab[o] = wu;
                    tab[o + 1] = wv;
                }
        ce = cudaMemcpy(ex->d_ic_tab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    }
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ex->s_h2d, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ex->s_d2h, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ex->stream2, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ex->ev_begin, cudaEventDisableTiming);
    for (int k = 0; k < orbgpu_extractor::kStreams - 2; ++k)
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ex->s_extra[k], cudaStreamNonBlocking);
    for (int k = 0; k < orbgpu_extractor::kStreams; ++k) {
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ex->s_aux[k], cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ex->ev_pyr[k], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ex->ev_blur[k], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ex->ev_fast[k], cudaEventDisableTiming);
    }
    if (ce != cudaSuccess) {
        std::string m = std::string("workspace allocation failed: ") + cudaGetErrorString(ce);
        orbgpu_extractor_destroy(ex);
        return fail(ORBGPU_ERR_CUDA, m);
    }
    *out = ex;
    return ORBGPU_OK;
}

int orbgpu_extractor_destroy(orbgpu_extractor* ex) {
    if (!ex) return ORBGPU_OK;
    cudaSetDevice(ex->device);
    if (ex->stream) { cudaStreamSynchronize(ex->stream); cudaStreamDestroy(ex->stream); }
    if (ex->s_h2d) { cudaStreamSynchronize(ex->s_h2d); cudaStreamDestroy(ex->s_h2d); }
    if (ex->s_d2h) { cudaStreamSynchronize(ex->s_d2h); cudaStreamDestroy(ex->s_d2h); }
    if (ex->stream2) { cudaStreamSynchronize(ex->stream2); cudaStreamDestroy(ex->stream2); }
    for (cudaEvent_t e : ex->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ex->ev_out) cudaEventDestroy(e);
    if (ex->ev_begin) cudaEventDestroy(ex->ev_begin);
    for (int k = 0; k < orbgpu_extractor::kStreams - 2; ++k)
        if (ex->s_extra[k]) { cudaStreamSynchronize(ex->s_extra[k]); cudaStreamDestroy(ex->s_extra[k]); }
    for (int k = 0; k < orbgpu_extractor::kStreams; ++k) {
        if (ex->s_aux[k]) { cudaStreamSynchronize(ex->s_aux[k]); cudaStreamDestroy(ex->s_aux[k]); }
        if (ex->ev_pyr[k]) cudaEventDestroy(ex->ev_pyr[k]);
        if (ex->ev_blur[k]) cudaEventDestroy(ex->ev_blur[k]);
        if (ex->ev_fast[k]) cudaEventDestroy(ex->ev_fast[k]);
    }
    for (int i = 0; i < 6; ++i) if (ex->ev[i]) cudaEventDestroy(ex->ev[i]);
    if (ex->ev_peer) cudaEventDestroy(ex->ev_peer);
    void* ptrs[] = {ex->d_color, ex->d_st_rows, ex->d_st_items, ex->d_st_sad, ex->d_st_u, ex->d_st_d, ex->d_ic_tab, ex->d_btiles, ex->d_segs, ex->d_tmaps, ex->d_pyr, ex->d_blur, ex->d_images, ex->d_cells, ex->d_taps, ex->d_cell_count, ex->d_cand_xy,
                    ex->d_cand_resp, ex->d_ot, ex->d_sel_xy, ex->d_sel_resp, ex->d_sel_count, ex->d_counts, ex->d_kp, ex->d_desc};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete ex;
    return ORBGPU_OK;
}

int orbgpu_extractor_tables(const orbgpu_extractor* ex, float* scales, int32_t* features_per_level, int32_t* umax) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    const int n = ex->nlevels;
    for (int i = 0; i < n; ++i) {
        if (scales) {
            scales[i] = ex->scale[i]; scales[n + i] = ex->inv_scale[i];
            scales[2 * n + i] = ex->sigma2[i]; scales[3 * n + i] = ex->inv_sigma2[i];
        }
        if (features_per_level) features_per_level[i] = ex->quota[i];
    }
    if (umax) for (int i = 0; i < 16; ++i) umax[i] = ex->umax[i];
    return ORBGPU_OK;
}

int orbgpu_extractor_max_keypoints(const orbgpu_extractor* ex) { return ex ? ex->kp_cap : 0; }

int orbgpu_extractor_static_tables(int nfeatures, float scale_factor, int nlevels, float* scales, int32_t* features_per_level, int32_t* umax) {
    if (nfeatures < 1 || nlevels < 1 || nlevels > og::kMaxLevels || !(scale_factor > 1.f)) return fail(ORBGPU_ERR_ARG, "bad extractor parameters");
    std::vector<float> sc, isc, s2, is2;
    std::vector<int> q, um;
    compute_tables(nfeatures, scale_factor, nlevels, sc, isc, s2, is2, q, um);
    for (int i = 0; i < nlevels; ++i) {
        if (scales) { scales[i] = sc[i]; scales[nlevels + i] = isc[i]; scales[2 * nlevels + i] = s2[i]; scales[3 * nlevels + i] = is2[i]; }
        if (features_per_level) features_per_level[i] = q[i];
    }
    if (umax) for (int i = 0; i < 16; ++i) umax[i] = um[i];
    return ORBGPU_OK;
}

int orbgpu_extract_batch_dev(orbgpu_extractor* ex, const uint8_t* images_dev, int batch, int width, int height,
                             size_t row_stride, size_t frame_stride, orbgpu_keypoint* kp_out_dev, uint8_t* desc_out_dev,
                             int kp_capacity, int32_t* counts_dev) {
    OG_NVTX("orbgpu_extract_batch_dev");
    int rc = check_call(ex, batch, width, height, row_stride, kp_capacity);
    if (rc) return rc;
    if (!images_dev || !kp_out_dev || !desc_out_dev || !counts_dev) return fail(ORBGPU_ERR_ARG, "null device pointer");
    // Optional (ORBGPU_DEV_SPLIT=k): the batch as k sub-batches over the compute streams, so that the latency-bound octree of one
    // sub-batch could run beside the issue-bound FAST / blur / descriptor kernels of its neighbours; the side streams fork from
    // and join ex->stream.  Measured on B200 at 1024 KITTI frames: 10.24 ms unsplit, 10.93 / 10.81 / 10.82 ms for k = 2 / 4 / 8
    // (profiles/r2_dev_split_sweep.txt) — co-resident octree CTAs take the registers FAST needs and gain nothing themselves, so
    // the default stays one in-order sequence.
    static const int split_env = []() { const char* e = getenv("ORBGPU_DEV_SPLIT"); return e ? atoi(e) : -1; }();
    static const int nstr_env = []() { const char* e = getenv("ORBGPU_DEV_STREAMS"); return e ? atoi(e) : 0; }();
    const int min_sub = 128;
    int nsub = split_env >= 0 ? split_env : 1;
    if (ex->profiling || nsub <= 1 || batch < 2 * min_sub) nsub = 1;
    nsub = std::min(nsub, batch / min_sub);
    if (nsub <= 1)
        return launch_extract(ex, images_dev, batch, row_stride, frame_stride, (og::KeyPoint*)kp_out_dev, desc_out_dev, kp_capacity, counts_dev);
    const int n_streams = std::min(nstr_env > 0 ? nstr_env : 4, std::min(nsub, (int)orbgpu_extractor::kStreams));
    while ((int)ex->ev_out.size() < n_streams) {
        cudaEvent_t a, b;
        OG_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        ex->ev_in.push_back(a);
        OG_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ex->ev_out.push_back(b);
    }
    OG_CUDA(cudaEventRecord(ex->ev_begin, ex->stream));
    for (int k = 1; k < n_streams; ++k) OG_CUDA(cudaStreamWaitEvent(ex->compute_stream(k), ex->ev_begin, 0));
    int launches = 0;
    for (int c = 0; c < nsub; ++c) {
        const int f0 = (int)((long long)batch * c / nsub), f1 = (int)((long long)batch * (c + 1) / nsub);
        rc = launch_extract(ex, images_dev, f1 - f0, row_stride, frame_stride, (og::KeyPoint*)kp_out_dev, desc_out_dev, kp_capacity, counts_dev, f0,
                            ex->compute_stream(c % n_streams));
        if (rc) return rc;
        launches += ex->last_launches;
    }
    for (int k = 1; k < n_streams; ++k) {
        OG_CUDA(cudaEventRecord(ex->ev_out[k], ex->compute_stream(k)));
        OG_CUDA(cudaStreamWaitEvent(ex->stream, ex->ev_out[k], 0));
    }
    ex->last_batch = batch;
    ex->last_launches = launches;
    return ORBGPU_OK;
}

int orbgpu_extract_batch(orbgpu_extractor* ex, const uint8_t* images, int batch, int width, int height, size_t row_stride,
                         size_t frame_stride, orbgpu_keypoint* kp_out, uint8_t* desc_out, int kp_capacity, int32_t* counts) {
    OG_NVTX("orbgpu_extract_batch");
    int rc = check_call(ex, batch, width, height, row_stride, kp_capacity);
    if (rc) return rc;
    if (!images || !kp_out || !desc_out || !counts) return fail(ORBGPU_ERR_ARG, "null pointer");
    cudaStream_t st = ex->stream;
    // Chunked pipeline: H2D (s_h2d) -> kernels (round robin over n_streams compute streams) -> D2H (s_d2h), chained by events
    // per chunk.  Measured on B200 at 1024 KITTI frames (e2e frames/s): 1 stream 51-61k, 2 streams 76k, 4 streams x 48-frame
    // chunks 80.4k, 8 streams no better; chunks of 128+ or a growing schedule lose (ORBGPU_CHUNK / ORBGPU_STREAMS override).
    static const int chunk_env = []() { const char* e = getenv("ORBGPU_CHUNK"); return e ? atoi(e) : 0; }();
    std::vector<std::pair<int, int>> chunks;   // (first frame, frames)
    {
        const int chunk_pref = chunk_env > 0 ? chunk_env : 48;
        const int chunk = batch <= chunk_pref + chunk_pref / 2 ? batch : chunk_pref;
        for (int f0 = 0; f0 < batch; f0 += chunk) chunks.push_back(std::make_pair(f0, std::min(chunk, batch - f0)));
    }
    const int nchunks = (int)chunks.size();
    while ((int)ex->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        OG_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        ex->ev_in.push_back(a);
        OG_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ex->ev_out.push_back(b);
    }
    // work queued earlier on the compute stream (device-resident calls) must not be overtaken by the copies
    OG_CUDA(cudaEventRecord(ex->ev_begin, st));
    OG_CUDA(cudaStreamWaitEvent(ex->s_h2d, ex->ev_begin, 0));
    OG_CUDA(cudaStreamWaitEvent(ex->s_d2h, ex->ev_begin, 0));
    static const int n_streams = []() { const char* e = getenv("ORBGPU_STREAMS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= orbgpu_extractor::kStreams ? v : 4; }();
    for (int k = 1; k < n_streams; ++k) OG_CUDA(cudaStreamWaitEvent(ex->compute_stream(k), ex->ev_begin, 0));
    const size_t fbytes = (size_t)width * height;
    const bool packed = row_stride == (size_t)width && frame_stride == fbytes;
    for (int c = 0; c < nchunks; ++c) {
        const int f0 = chunks[c].first, nb = chunks[c].second;
        if (packed) {
            OG_CUDA(cudaMemcpyAsync(ex->d_images + f0 * fbytes, images + f0 * fbytes, fbytes * nb, cudaMemcpyHostToDevice, ex->s_h2d));
        } else {
            for (int f = f0; f < f0 + nb; ++f)
                OG_CUDA(cudaMemcpy2DAsync(ex->d_images + f * fbytes, width, images + (size_t)f * frame_stride, row_stride, width, height,
                                          cudaMemcpyHostToDevice, ex->s_h2d));
        }
        cudaStream_t cs = ex->compute_stream(c % n_streams);
        OG_CUDA(cudaEventRecord(ex->ev_in[c], ex->s_h2d));
        OG_CUDA(cudaStreamWaitEvent(cs, ex->ev_in[c], 0));
        rc = launch_extract(ex, ex->d_images, nb, width, fbytes, ex->d_kp, ex->d_desc, ex->kp_cap, ex->d_counts, f0, cs);
        if (rc) return rc;
        OG_CUDA(cudaEventRecord(ex->ev_out[c], cs));
        OG_CUDA(cudaStreamWaitEvent(ex->s_d2h, ex->ev_out[c], 0));
        // D2H of this chunk's slots
        OG_CUDA(cudaMemcpyAsync(counts + f0, ex->d_counts + f0, (size_t)nb * 4, cudaMemcpyDeviceToHost, ex->s_d2h));
        const size_t kb = (size_t)ex->kp_cap * sizeof(og::KeyPoint), db = (size_t)ex->kp_cap * 32;
        if (kp_capacity == ex->kp_cap) {
            OG_CUDA(cudaMemcpyAsync(kp_out + (size_t)f0 * kp_capacity, ex->d_kp + (size_t)f0 * ex->kp_cap, kb * nb, cudaMemcpyDeviceToHost, ex->s_d2h));
            OG_CUDA(cudaMemcpyAsync(desc_out + (size_t)f0 * kp_capacity * 32, ex->d_desc + (size_t)f0 * ex->kp_cap * 32, db * nb, cudaMemcpyDeviceToHost, ex->s_d2h));
        } else {
            OG_CUDA(cudaMemcpy2DAsync(kp_out + (size_t)f0 * kp_capacity, (size_t)kp_capacity * sizeof(og::KeyPoint), ex->d_kp + (size_t)f0 * ex->kp_cap, kb, kb, nb,
                                      cudaMemcpyDeviceToHost, ex->s_d2h));
            OG_CUDA(cudaMemcpy2DAsync(desc_out + (size_t)f0 * kp_capacity * 32, (size_t)kp_capacity * 32, ex->d_desc + (size_t)f0 * ex->kp_cap * 32, db, db, nb,
                                      cudaMemcpyDeviceToHost, ex->s_d2h));
        }
    }
    ex->last_batch = batch;
    ex->last_launches *= nchunks;
    OG_CUDA(cudaStreamSynchronize(ex->s_d2h));
    for (int k = 1; k < n_streams; ++k) OG_CUDA(cudaStreamSynchronize(ex->compute_stream(k)));
    OG_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

int orbgpu_extract_batch_color(orbgpu_extractor* ex, const uint8_t* images, int batch, int width, int height, int channels, int rgb_order,
                               size_t row_stride, size_t frame_stride, uint8_t* gray_out, orbgpu_keypoint* kp_out, uint8_t* desc_out,
                               int kp_capacity, int32_t* counts) {
    OG_NVTX("orbgpu_extract_batch_color");
    if (channels != 3 && channels != 4) return fail(ORBGPU_ERR_ARG, "colour input must have 3 or 4 interleaved channels");
    int rc = check_call(ex, batch, width, height, row_stride / (size_t)channels, kp_capacity);
    if (rc) return rc;
    if (row_stride < (size_t)width * channels) return fail(ORBGPU_ERR_ARG, "row_stride smaller than width * channels");
    if (!images || !kp_out || !desc_out || !counts) return fail(ORBGPU_ERR_ARG, "null pointer");
    cudaStream_t st = ex->stream;
    const size_t in_bytes = (size_t)(batch - 1) * frame_stride + (size_t)(height - 1) * row_stride + (size_t)width * channels;
    if (in_bytes > ex->color_cap) {
        if (ex->d_color) cudaFree(ex->d_color);
        ex->d_color = nullptr;
        ex->color_cap = 0;
        OG_CUDA(cudaMalloc((void**)&ex->d_color, in_bytes));
        ex->color_cap = in_bytes;
    }
    OG_CUDA(cudaMemcpyAsync(ex->d_color, images, in_bytes, cudaMemcpyHostToDevice, st));
    og::k_cvt_gray<<<dim3(((width + 3) / 4 + 255) / 256, height, batch), 256, 0, st>>>(ex->d_color, (long long)row_stride, (long long)frame_stride,
                                                                                  channels, rgb_order ? 1 : 0, width, height, ex->d_images);
    const size_t fbytes = (size_t)width * height;
    if (gray_out) OG_CUDA(cudaMemcpyAsync(gray_out, ex->d_images, fbytes * batch, cudaMemcpyDeviceToHost, st));
    rc = launch_extract(ex, ex->d_images, batch, width, fbytes, ex->d_kp, ex->d_desc, ex->kp_cap, ex->d_counts);
    if (rc) return rc;
    ++ex->last_launches;
    OG_CUDA(cudaMemcpyAsync(counts, ex->d_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    const size_t kb = (size_t)ex->kp_cap * sizeof(og::KeyPoint), db = (size_t)ex->kp_cap * 32;
    OG_CUDA(cudaMemcpy2DAsync(kp_out, (size_t)kp_capacity * sizeof(og::KeyPoint), ex->d_kp, kb, kb, batch, cudaMemcpyDeviceToHost, st));
    OG_CUDA(cudaMemcpy2DAsync(desc_out, (size_t)kp_capacity * 32, ex->d_desc, db, db, batch, cudaMemcpyDeviceToHost, st));
    OG_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

int orbgpu_extract(orbgpu_extractor* ex, const uint8_t* image, int width, int height, size_t row_stride,
                   orbgpu_keypoint* kp_out, uint8_t* desc_out, int kp_capacity, int* n_out) {
    OG_NVTX("orbgpu_extract");
    if (!n_out) return fail(ORBGPU_ERR_ARG, "null n_out");
    *n_out = 0;
    if (!image || width <= 0 || height <= 0) return ORBGPU_OK;  // empty image: silent return (:1046-1047)
    int32_t n = 0;
    int rc = orbgpu_extract_batch(ex, image, 1, width, height, row_stride, row_stride * height, kp_out, desc_out, kp_capacity, &n);
    if (rc) return rc;
    *n_out = n;
    return ORBGPU_OK;
}

int orbgpu_extractor_sync(orbgpu_extractor* ex) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    OG_CUDA(cudaSetDevice(ex->device));
    OG_CUDA(cudaStreamSynchronize(ex->stream));
    return ORBGPU_OK;
}

int orbgpu_extractor_stream(orbgpu_extractor* ex, void** stream_out) {
    if (!ex || !stream_out) return fail(ORBGPU_ERR_ARG, "null argument");
    *stream_out = (void*)ex->stream;
    return ORBGPU_OK;
}

int orbgpu_extractor_last_launches(const orbgpu_extractor* ex) { return ex ? ex->last_launches : 0; }

}  // extern "C" (re-opened below)

// Internal (not part of the C ABI): where the results of the extractor's last batch lie in HBM — consumed by
// orbgpu_frame_set_from_extraction (og_match.cu).
int og_extractor_last_results(orbgpu_extractor* ex, int* device, int* batch, int* stride, const og::KeyPoint** kp, const uint8_t** desc,
                              const int32_t** counts, cudaStream_t* stream) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    if (ex->last_batch < 1 || !ex->last_kp) return fail(ORBGPU_ERR_ARG, "the extractor has not processed a batch yet");
    *device = ex->device; *batch = ex->last_batch; *stride = ex->last_stride;
    *kp = ex->last_kp; *desc = ex->last_desc; *counts = ex->last_counts; *stream = ex->stream;
    return ORBGPU_OK;
}

extern "C" {

int orbgpu_extractor_set_eager_frame(orbgpu_extractor* ex, int enable) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    ex->eager_frame = enable != 0;
    return ORBGPU_OK;
}

int orbgpu_extractor_set_profiling(orbgpu_extractor* ex, int enable) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    OG_CUDA(cudaSetDevice(ex->device));
    if (enable && !ex->ev[0])
        for (int i = 0; i < 6; ++i) OG_CUDA(cudaEventCreate(&ex->ev[i]));
    ex->profiling = enable != 0;
    return ORBGPU_OK;
}

int orbgpu_extractor_stage_ms(orbgpu_extractor* ex, float* ms5) {
    if (!ex || !ms5 || !ex->ev[0]) return fail(ORBGPU_ERR_ARG, "profiling was not enabled");
    OG_CUDA(cudaSetDevice(ex->device));
    OG_CUDA(cudaEventSynchronize(ex->ev[5]));
    for (int i = 0; i < 5; ++i) OG_CUDA(cudaEventElapsedTime(&ms5[i], ex->ev[i], ex->ev[i + 1]));
    return ORBGPU_OK;
}

int orbgpu_extractor_level_dims(const orbgpu_extractor* ex, int level, int* width, int* height) {
    if (!ex || level < 0 || level >= ex->nlevels || ex->cur_w == 0) return fail(ORBGPU_ERR_ARG, "bad level or no frame processed yet");
    if (width) *width = ex->P.lv[level].w;
    if (height) *height = ex->P.lv[level].h;
    return ORBGPU_OK;
}

static int read_plane(orbgpu_extractor* ex, const uint8_t* base, int frame, int level, int bordered, uint8_t* out, size_t out_stride) {
    if (!ex || !out || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->last_batch)
        return fail(ORBGPU_ERR_ARG, "bad frame/level");
    OG_CUDA(cudaSetDevice(ex->device));
    const og::Level& L = ex->P.lv[level];
    const uint8_t* src = base + L.base + (long long)frame * L.frame_stride;
    const int W = bordered ? L.w + 2 * og::kEdge : L.w, H = bordered ? L.rows : L.h;
    src += bordered ? (og::kXPad - og::kEdge) : ((long long)og::kEdge * L.pitch + og::kXPad);
    if (out_stride < (size_t)W) return fail(ORBGPU_ERR_ARG, "out_stride too small");
    OG_CUDA(cudaMemcpy2DAsync(out, out_stride, src, L.pitch, W, H, cudaMemcpyDeviceToHost, ex->stream));
    OG_CUDA(cudaStreamSynchronize(ex->stream));
    return ORBGPU_OK;
}

int orbgpu_extractor_read_level(orbgpu_extractor* ex, int frame, int level, int bordered, uint8_t* out, size_t out_stride) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    if (bordered && ex->frame_pending && ex->last_batch > 0) {
        // the frame of every level of the last call, once (ORBextractor.cc:1122-1128)
        OG_CUDA(cudaSetDevice(ex->device));
        og::ExtractParams P = ex->P;
        P.batch = ex->last_batch;
        P.frame0 = 0;
        const int sides0 = 2 * P.lv[0].h;
        og::k_border_sides<<<dim3((sides0 + og::kBorderThreads - 1) / og::kBorderThreads, P.n_levels, P.batch), og::kBorderThreads, 0, ex->stream>>>(P);
        og::k_border_caps<<<dim3(2 * og::kEdge, P.n_levels, P.batch), og::kBorderThreads, 0, ex->stream>>>(P);
        OG_CUDA(cudaGetLastError());
        ex->frame_pending = false;
    }
    return read_plane(ex, ex->d_pyr, frame, level, bordered, out, out_stride);
}

int orbgpu_extractor_read_blurred(orbgpu_extractor* ex, int frame, int level, uint8_t* out, size_t out_stride) {
    if (!ex) return fail(ORBGPU_ERR_ARG, "null extractor");
    return read_plane(ex, ex->d_blur, frame, level, 0, out, out_stride);
}

int orbgpu_extractor_read_points(orbgpu_extractor* ex, int frame, int level, int stage, orbgpu_keypoint* out, int capacity, int* n_out) {
    if (!ex || !n_out || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->last_batch || (stage != 0 && stage != 1))
        return fail(ORBGPU_ERR_ARG, "bad argument");
    OG_CUDA(cudaSetDevice(ex->device));
    const og::ExtractParams& P = ex->P;
    const og::Level& L = P.lv[level];
    std::vector<uint32_t> xy;
    std::vector<uint8_t> rr;
    std::vector<float> angles;
    if (stage == 0) {
        std::vector<int32_t> cc(L.n_cells);
        OG_CUDA(cudaMemcpy(cc.data(), P.cell_count + (size_t)frame * P.total_cells + L.cell_base, (size_t)L.n_cells * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> sxy(L.cand_cap);
        std::vector<uint8_t> srr(L.cand_cap);
        OG_CUDA(cudaMemcpy(sxy.data(), P.cand_xy + (size_t)frame * P.total_cand_cap + L.cand_base, (size_t)L.cand_cap * 4, cudaMemcpyDeviceToHost));
        OG_CUDA(cudaMemcpy(srr.data(), P.cand_resp + (size_t)frame * P.total_cand_cap + L.cand_base, (size_t)L.cand_cap, cudaMemcpyDeviceToHost));
        for (int c = 0; c < L.n_cells; ++c)
            for (int k = 0; k < cc[c]; ++k) {
                xy.push_back(sxy[ex->cells[L.cell_base + c].slot + k]);
                rr.push_back(srr[ex->cells[L.cell_base + c].slot + k]);
            }
    } else {
        std::vector<int32_t> sc(ex->nlevels);
        OG_CUDA(cudaMemcpy(sc.data(), P.sel_count + (size_t)frame * ex->nlevels, (size_t)ex->nlevels * 4, cudaMemcpyDeviceToHost));
        const int n = sc[level];
        int off = 0;
        for (int l = 0; l < level; ++l) off += sc[l];
        xy.resize(n); rr.resize(n);
        if (n) {
            OG_CUDA(cudaMemcpy(xy.data(), P.sel_xy + (size_t)frame * P.total_sel_cap + L.sel_base, (size_t)n * 4, cudaMemcpyDeviceToHost));
            OG_CUDA(cudaMemcpy(rr.data(), P.sel_resp + (size_t)frame * P.total_sel_cap + L.sel_base, (size_t)n, cudaMemcpyDeviceToHost));
            std::vector<og::KeyPoint> kp(n);
            // the final records of the last call, wherever that call wrote them (own buffers or the caller's device buffers)
            OG_CUDA(cudaMemcpy(kp.data(), ex->last_kp + (size_t)frame * ex->last_stride + off, (size_t)n * sizeof(og::KeyPoint), cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; ++i) angles.push_back(kp[i].angle);
        }
    }
    *n_out = (int)xy.size();
    for (int i = 0; i < (int)xy.size() && i < capacity; ++i) {
        orbgpu_keypoint k;
        const int add = stage == 0 ? 0 : 16;
        k.x = (float)((xy[i] & 0xffffu) + add);
        k.y = (float)((xy[i] >> 16) + add);
        k.size = stage == 0 ? 7.f : L.kp_size;
        k.angle = stage == 0 ? -1.f : angles[i];
        k.response = (float)rr[i];
        k.octave = stage == 0 ? 0 : level;
        k.class_id = -1;
        out[i] = k;
    }
    return ORBGPU_OK;
}

int orbgpu_octree_last_path(const orbgpu_extractor* ex) { return ex ? ex->last_octree_direct : 0; }

int orbgpu_octree(orbgpu_extractor* ex, const orbgpu_keypoint* candidates, int n, int min_x, int max_x, int min_y, int max_y,
                  int n_features, orbgpu_keypoint* out, int capacity, int* n_out) {
    if (!ex || !n_out || n < 0 || n_features < 1) return fail(ORBGPU_ERR_ARG, "bad argument");
    OG_CUDA(cudaSetDevice(ex->device));
    const int width = max_x - min_x, height = max_y - min_y;
    if (width <= 0 || height <= 0) return fail(ORBGPU_ERR_ARG, "empty rectangle");
    const int n_ini = (int)roundf((float)width / (float)height);
    if (n_ini < 1) return fail(ORBGPU_ERR_ARG, "rectangle taller than 2:1");
    const float hx = (float)width / (float)n_ini;
    const int cap = std::max(n, 1), sel_cap = std::max(n_features + 3, 4 * n_ini), node_cap = sel_cap + 8;
    if (node_cap > 65535) return fail(ORBGPU_ERR_ARG, "n_features too large");
    std::vector<uint32_t> xy(cap);
    std::vector<uint8_t> rr(cap);
    for (int i = 0; i < n; ++i) {
        xy[i] = ((uint32_t)(int)candidates[i].y << 16) | (uint32_t)(int)candidates[i].x;
        rr[i] = (uint8_t)(int)candidates[i].response;
    }
    uint8_t* ws = nullptr;
    uint32_t *d_xy = nullptr, *d_oxy = nullptr;
    uint8_t *d_rr = nullptr, *d_orr = nullptr;
    int* d_n = nullptr;
    OG_CUDA(cudaMalloc((void**)&ws, octree_ws_bytes(cap, node_cap)));
    OG_CUDA(cudaMalloc((void**)&d_xy, (size_t)cap * 4));
    OG_CUDA(cudaMalloc((void**)&d_rr, cap));
    OG_CUDA(cudaMalloc((void**)&d_oxy, (size_t)sel_cap * 4));
    OG_CUDA(cudaMalloc((void**)&d_orr, sel_cap));
    OG_CUDA(cudaMalloc((void**)&d_n, 8));
    OG_CUDA(cudaMemcpy(d_xy, xy.data(), (size_t)cap * 4, cudaMemcpyHostToDevice));
    OG_CUDA(cudaMemcpy(d_rr, rr.data(), cap, cudaMemcpyHostToDevice));
    // ORBGPU_OCT_DIRECT=0: the division-pass state machine alone (the general path the direct construction falls back to)
    static const int direct_env = []() { const char* e = getenv("ORBGPU_OCT_DIRECT"); return e ? atoi(e) : 1; }();
    const int Dh = og::ot2_depth(n_ini, og::kOt2Budget);
    const int kcap = 4096;
    int direct_bytes = direct_env ? (int)og::ot2_smem_bytes(n_ini, Dh, std::max(node_cap, og::kOctThreads), kcap) : 0;
    if (direct_bytes > og::kOctSmem) direct_bytes = 0;
    OG_CUDA(cudaFuncSetAttribute(og::k_octree_single, cudaFuncAttributeMaxDynamicSharedMemorySize, og::kOctSmem));
    og::k_octree_single<<<1, og::kOctThreads, direct_bytes, ex->stream>>>(ws, cap, node_cap, d_xy, d_rr, n, n_ini, hx, height, n_features, d_oxy, d_orr,
                                                                         sel_cap, d_n, direct_bytes, kcap);
    OG_CUDA(cudaGetLastError());
    OG_CUDA(cudaStreamSynchronize(ex->stream));
    int mm[2] = {0, 0};
    OG_CUDA(cudaMemcpy(mm, d_n, 8, cudaMemcpyDeviceToHost));
    const int m = mm[0];
    ex->last_octree_direct = mm[1];
    std::vector<uint32_t> oxy(std::max(m, 1));
    std::vector<uint8_t> orr(std::max(m, 1));
    if (m) {
        OG_CUDA(cudaMemcpy(oxy.data(), d_oxy, (size_t)m * 4, cudaMemcpyDeviceToHost));
        OG_CUDA(cudaMemcpy(orr.data(), d_orr, m, cudaMemcpyDeviceToHost));
    }
    cudaFree(ws); cudaFree(d_xy); cudaFree(d_rr); cudaFree(d_oxy); cudaFree(d_orr); cudaFree(d_n);
    *n_out = m;
    for (int i = 0; i < m && i < capacity; ++i) {
        orbgpu_keypoint k;
        memset(&k, 0, sizeof(k));
        k.x = (float)(oxy[i] & 0xffffu);
        k.y = (float)(oxy[i] >> 16);
        k.response = (float)orr[i];
        out[i] = k;
    }
    return ORBGPU_OK;
}

// ---- Frame::ComputeStereoMatches (Frame.cc:501-675) ---------------------------------------------------------------
static int stereo_launch(orbgpu_extractor* L, orbgpu_extractor* R, float mb, float mbf, float* d_u, float* d_d, int out_stride) {
    if (!L || !R) return fail(ORBGPU_ERR_ARG, "null extractor");
    if (L->device != R->device) return fail(ORBGPU_ERR_ARG, "left and right extractor live on different devices");
    if (L->last_batch < 1 || L->last_batch != R->last_batch || L->cur_w != R->cur_w || L->cur_h != R->cur_h || L->nlevels != R->nlevels ||
        L->last_stride != R->last_stride || !L->last_kp || !R->last_kp)
        return fail(ORBGPU_ERR_ARG, "stereo matching needs the left and right extractor to have just processed equally sized batches of equally sized images");
    // the reference indexes both pyramids with the frame's single scale table (Frame.cc:598-616): the two extractors must agree on it
    for (int l = 0; l < L->nlevels; ++l)
        if (L->scale[l] != R->scale[l] || L->P.lv[l].w != R->P.lv[l].w || L->P.lv[l].h != R->P.lv[l].h)
            return fail(ORBGPU_ERR_ARG, "stereo matching needs the left and right extractor to share the scale factor and the level geometry");
    if (!(mb > 0.f) || out_stride < L->last_stride) return fail(ORBGPU_ERR_ARG, "bad mb or output stride");
    if (L->last_stride > (int)og::kPosMask) return fail(ORBGPU_ERR_ARG, "too many key points per frame");
    OG_CUDA(cudaSetDevice(L->device));
    const int B = L->last_batch, n_rows = L->P.lv[0].h;
    const int band = 2 * (int)std::ceil(2.0 * R->scale[R->nlevels - 1]) + 3;
    const size_t items_cap = (size_t)R->last_stride * band;
    auto grow = [&](void** p, size_t& cap, size_t want, size_t elt) -> cudaError_t {
        if (want <= cap) return cudaSuccess;
        if (*p) cudaFree(*p);
        *p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(p, want * elt);
        if (e == cudaSuccess) cap = want;
        return e;
    };
    OG_CUDA(cudaStreamSynchronize(L->stream));   // scratch may still be in use by an earlier stereo call
    OG_CUDA(grow((void**)&L->d_st_rows, L->st_cap_rows, (size_t)B * (n_rows + 1), 4));
    OG_CUDA(grow((void**)&L->d_st_items, L->st_cap_items, (size_t)B * items_cap, 4));
    if ((size_t)B * L->last_stride > L->st_cap_kp) {
        if (L->d_st_sad) cudaFree(L->d_st_sad);
        if (L->d_st_u) cudaFree(L->d_st_u);
        if (L->d_st_d) cudaFree(L->d_st_d);
        L->d_st_sad = nullptr; L->d_st_u = nullptr; L->d_st_d = nullptr; L->st_cap_kp = 0;
        OG_CUDA(cudaMalloc((void**)&L->d_st_sad, (size_t)B * L->last_stride * 4));
        OG_CUDA(cudaMalloc((void**)&L->d_st_u, (size_t)B * L->last_stride * 4));
        OG_CUDA(cudaMalloc((void**)&L->d_st_d, (size_t)B * L->last_stride * 4));
        L->st_cap_kp = (size_t)B * L->last_stride;
    }
    if (!L->ev_peer) OG_CUDA(cudaEventCreateWithFlags(&L->ev_peer, cudaEventDisableTiming));
    OG_CUDA(cudaEventRecord(L->ev_peer, R->stream));
    OG_CUDA(cudaStreamWaitEvent(L->stream, L->ev_peer, 0));

    og::StereoArgs A;
    A.PL = L->P; A.PR = R->P;
    A.PL.frame0 = 0; A.PR.frame0 = 0;
    A.kpL = L->last_kp; A.kpR = R->last_kp;
    A.descL = L->last_desc; A.descR = R->last_desc;
    A.cntL = L->last_counts; A.cntR = R->last_counts;
    A.kp_stride = L->last_stride;
    A.n_rows = n_rows;
    A.max_d = mbf / mb;   // :527-528
    A.mbf = mbf;
    for (int l = 0; l < og::kMaxLevels; ++l) A.inv_scale[l] = l < L->nlevels ? L->inv_scale[l] : 0.f;
    A.row_start = L->d_st_rows;
    A.row_items = L->d_st_items;
    A.items_cap = (int)items_cap;
    A.u_right = d_u ? d_u : L->d_st_u;
    A.depth = d_d ? d_d : L->d_st_d;
    A.sad = L->d_st_sad;
    A.out_stride = d_u ? out_stride : L->last_stride;
    cudaStream_t st = L->stream;
    const size_t smem = (size_t)(n_rows + 1) * 4;
    if (smem > 48 * 1024) return fail(ORBGPU_ERR_ARG, "image too tall for the row index");
    og::k_stereo_rows<<<B, og::kStereoThreads, smem, st>>>(A);
    og::k_stereo_match<<<dim3((L->last_stride + og::kStereoThreads / 32 - 1) / (og::kStereoThreads / 32), B), og::kStereoThreads, 0, st>>>(A);
    og::k_stereo_filter<<<B, og::kStereoThreads, 0, st>>>(A);
    OG_CUDA(cudaGetLastError());
    L->last_launches = 3;
    return ORBGPU_OK;
}

int orbgpu_stereo_matches_dev(orbgpu_extractor* left, orbgpu_extractor* right, float mb, float mbf, float* u_right_dev, float* depth_dev,
                              int out_stride) {
    OG_NVTX("orbgpu_stereo_matches_dev");
    if (!u_right_dev || !depth_dev) return fail(ORBGPU_ERR_ARG, "null output");
    return stereo_launch(left, right, mb, mbf, u_right_dev, depth_dev, out_stride);
}

int orbgpu_stereo_matches(orbgpu_extractor* left, orbgpu_extractor* right, float mb, float mbf, float* u_right, float* depth, int out_stride) {
    OG_NVTX("orbgpu_stereo_matches");
    if (!u_right || !depth) return fail(ORBGPU_ERR_ARG, "null output");
    int rc = stereo_launch(left, right, mb, mbf, nullptr, nullptr, out_stride);
    if (rc) return rc;
    const int B = left->last_batch, s = left->last_stride;
    OG_CUDA(cudaMemcpy2DAsync(u_right, (size_t)out_stride * 4, left->d_st_u, (size_t)s * 4, (size_t)s * 4, B, cudaMemcpyDeviceToHost, left->stream));
    OG_CUDA(cudaMemcpy2DAsync(depth, (size_t)out_stride * 4, left->d_st_d, (size_t)s * 4, (size_t)s * 4, B, cudaMemcpyDeviceToHost, left->stream));
    OG_CUDA(cudaStreamSynchronize(left->stream));
    return ORBGPU_OK;
}

}  // extern "C"

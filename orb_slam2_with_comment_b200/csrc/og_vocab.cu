// og_vocab.cu — ORBVocabulary::transform on the device (DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>), sm_100a.
//
// What the reference does per frame (Frame::ComputeBoW, Frame.cc:425-432 -> TemplatedVocabulary.h:1127-1197): every
// descriptor walks the vocabulary tree from the root, at each node taking the child with the smallest FORB::distance (first
// minimum wins, :1237-1247), until it reaches a leaf (a word).  BowVector[word] accumulates the word's weight
// (BowVector::addWeight, BowVector.cpp:36-48) and is L1-normalised (:64-90); FeatureVector[node at level L - levelsup]
// collects the feature indices (FeatureVector.cpp:30-45).
//
// Here: k_voc_descend gives every feature a sub-warp group (16 or 32 lanes, one child per lane, group-wide minimum of
// dist << 8 | child position = the reference's strict-`<` scan); k_voc_frame sorts one frame's (node, index) and (word, index)
// keys in shared memory (bitonic), turns the runs into the two maps and normalises — the floating-point sums run
// sequentially in ascending word order like the std::map iteration of the reference, so the doubles are bit-exact;
// k_voc_offsets scans the per-frame sizes; k_voc_compact writes the batch-wide CSR arrays.
#include "og_nvtx.h"
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/orbgpu.h"
#include "og_match.cuh"

int og_fail(int code, const std::string& msg);  // og_capi.cu (thread-local last error)

namespace og {

struct VocView {
    const int32_t* child_off;   // [n_nodes + 1]
    const int32_t* child_ids;   // children in push_back order (TemplatedVocabulary.h:1392)
    const uint8_t* node_desc;   // [n_nodes][32]
    const int32_t* node_word;   // word id of a leaf, -1 for inner nodes
    const double* node_weight;
};

constexpr int kVocThreads = 256;
constexpr unsigned long long kNoKey = ~0ull;

// One group of G lanes per feature.
template <int G>
__global__ void __launch_bounds__(kVocThreads) k_voc_descend(VocView V, const uint8_t* __restrict__ desc, int n_features, int nid_level,
                                                             uint32_t* __restrict__ feat_word, uint32_t* __restrict__ feat_node,
                                                             double* __restrict__ feat_weight) {
    const int gid = (blockIdx.x * kVocThreads + threadIdx.x) / G;
    if (gid >= n_features) return;   // whole groups leave together
    const int j = threadIdx.x & (G - 1);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
    const Desc f = load_desc(desc, gid);
    int cur = 0, level = 0, nid = nid_level <= 0 ? 0 : -1;
    for (;;) {
        const int off = __ldg(V.child_off + cur), cnt = __ldg(V.child_off + cur + 1) - off;
        if (cnt == 0) break;
        ++level;
        uint32_t key = 0xffffffffu;
        int c = 0;
        if (j < cnt) {
            c = __ldg(V.child_ids + off + j);
            key = ((uint32_t)hamming256(f, load_desc(V.node_desc, c)) << 8) | (uint32_t)j;
        }
        uint32_t best = key;
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(gmask, best, d, G));
        cur = __shfl_sync(gmask, c, (int)(best & 0xffu), G);
        if (level == nid_level) nid = cur;
    }
    if (j == 0) {
        feat_word[gid] = (uint32_t)__ldg(V.node_word + cur);
        feat_node[gid] = (uint32_t)(nid < 0 ? cur : nid);
        feat_weight[gid] = __ldg(V.node_weight + cur);
    }
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* keys, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int h = threadIdx.x; h < (P >> 1); h += kVocThreads) {   // one compare-exchange per thread step
                const int i = ((h & ~(j - 1)) << 1) | (h & (j - 1)), ixj = i | j;
                const unsigned long long a = keys[i], b = keys[ixj];
                if ((a > b) == ((i & k) == 0)) {
                    keys[i] = b;
                    keys[ixj] = a;
                }
            }
            __syncthreads();
        }
    }
}

// exclusive prefix of a flag over the block (in thread order) + block total; `wsum` = shared int[kVocThreads / 32]
__device__ __forceinline__ int block_flag_scan(bool flag, int* wsum, int& total) {
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();   // wsum free from the previous round
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kVocThreads / 32; ++i) {
        const int v = wsum[i];
        if (i < (int)w) before += v;
        tot += v;
    }
    total = tot;
    return before + __popc(bal & ((1u << lane) - 1u));
}

struct FrameArgs {
    const int32_t* kp_off;
    const uint32_t *feat_word, *feat_node;
    const double* feat_weight;
    // per-frame segments, frame f at offset kp_off[f]
    int32_t *seg_node_id, *seg_node_start, *seg_feat;
    uint32_t* seg_word;
    double* seg_val;
    int32_t* counts;   // [n_frames][3]: nodes, words, valid features
    int accumulate;    // TF_IDF / TF: addWeight; IDF / BINARY: addIfNotExist
    int norm;          // 0 none (DOT_PRODUCT), 1 L1, 2 L2
    // frames with more descriptors than a CTA's shared memory can sort (p_max > kVocSmemFrame): the same arrays in HBM, frame f at
    // gws + f * gws_stride (null: dynamic shared memory)
    unsigned char* gws;
    size_t gws_stride;
};
constexpr int kVocSmemFrame = 8192;   // descriptors of one frame (rounded up to a power of two) that fit the shared-memory sort

// One block per frame.  Shared: keys[P] (u64) | vals[P] (double, also used as int head positions).
__global__ void __launch_bounds__(kVocThreads) k_voc_frame(FrameArgs A, int p_max) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = A.gws ? A.gws + (size_t)blockIdx.x * A.gws_stride : smem_dyn;   // block-private either way
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    double* vals = reinterpret_cast<double*>(smem_raw + (size_t)p_max * 8);
    int* hpos = reinterpret_cast<int*>(smem_raw + (size_t)p_max * 16);   // [p_max + 1]
    __shared__ int wsum[kVocThreads / 32];
    __shared__ double s_norm;
    const int f = blockIdx.x, t = threadIdx.x;
    const int base = A.kp_off[f], n = A.kp_off[f + 1] - base;
    int P = 1;
    while (P < n) P <<= 1;

    // ---- FeatureVector: stable grouping of the valid feature indices by node id
    int nvalid = 0;
    for (int i0 = 0; i0 < P; i0 += kVocThreads) {
        const int i = i0 + t;
        const bool ok = i < n && A.feat_weight[base + i] > 0.0;
        if (i < P) keys[i] = ok ? ((unsigned long long)A.feat_node[base + i] << 32) | (unsigned)i : kNoKey;
        nvalid += __syncthreads_count(ok);
    }
    __syncthreads();
    bitonic_sort_u64(keys, P);
    int n_nodes = 0;
    for (int i0 = 0; i0 < nvalid; i0 += kVocThreads) {
        const int i = i0 + t;
        const bool in = i < nvalid;
        const unsigned long long k = in ? keys[i] : 0;
        const bool head = in && (i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(k >> 32));
        int tot;
        const int r = n_nodes + block_flag_scan(head, wsum, tot);
        if (head) {
            A.seg_node_id[base + r] = (int32_t)(k >> 32);
            A.seg_node_start[base + r] = i;
        }
        if (in) A.seg_feat[base + i] = (int32_t)(uint32_t)k;
        n_nodes += tot;
    }
    __syncthreads();

    // ---- BowVector: runs of equal word ids
    for (int i = t; i < P; i += kVocThreads) {
        const bool ok = i < n && A.feat_weight[base + i] > 0.0;
        keys[i] = ok ? ((unsigned long long)A.feat_word[base + i] << 32) | (unsigned)i : kNoKey;
    }
    __syncthreads();
    bitonic_sort_u64(keys, P);
    int n_words = 0;
    for (int i0 = 0; i0 < nvalid; i0 += kVocThreads) {
        const int i = i0 + t;
        const bool in = i < nvalid;
        const bool head = in && (i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(keys[i] >> 32));
        int tot;
        const int r = n_words + block_flag_scan(head, wsum, tot);
        if (head) hpos[r] = i;
        n_words += tot;
    }
    if (t == 0) hpos[n_words] = nvalid;
    __syncthreads();
    for (int r = t; r < n_words; r += kVocThreads) {
        const int i = hpos[r], cnt = hpos[r + 1] - i;
        const double w = A.feat_weight[base + (int)(uint32_t)keys[i]];   // every feature of a word carries the word's weight
        double v = w;
        if (A.accumulate)
            for (int c = 1; c < cnt; ++c) v += w;                        // addWeight once per feature, in feature order
        vals[r] = v;
    }
    __syncthreads();
    if (A.accumulate && A.norm == 0 && n_words > 0) {                    // TemplatedVocabulary.h:1166-1172
        const double nd = (double)n_words;
        for (int r = t; r < n_words; r += kVocThreads) vals[r] /= nd;
        __syncthreads();
    }
    if (A.norm != 0) {                                                    // BowVector::normalize, BowVector.cpp:64-90
        if (t == 0) {
            double norm = 0.0;
            if (A.norm == 1) {
                for (int r = 0; r < n_words; ++r) norm += fabs(vals[r]);
            } else {
                for (int r = 0; r < n_words; ++r) norm += vals[r] * vals[r];
                norm = sqrt(norm);
            }
            s_norm = norm;
        }
        __syncthreads();
        const double norm = s_norm;
        if (norm > 0.0)
            for (int r = t; r < n_words; r += kVocThreads) vals[r] /= norm;
        __syncthreads();
    }
    for (int r = t; r < n_words; r += kVocThreads) {
        A.seg_word[base + r] = (uint32_t)(keys[hpos[r]] >> 32);
        A.seg_val[base + r] = vals[r];
    }
    if (t == 0) {
        A.counts[3 * f] = n_nodes;
        A.counts[3 * f + 1] = n_words;
        A.counts[3 * f + 2] = nvalid;
    }
}

// Exclusive scan of the three per-frame sizes (single block).
__global__ void __launch_bounds__(1024) k_voc_offsets(const int32_t* __restrict__ counts, int n_frames, int32_t* __restrict__ node_off,
                                                      int32_t* __restrict__ word_off, int32_t* __restrict__ valid_off) {
    __shared__ int wtot[3][32];
    __shared__ int run[3];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t < 3) run[t] = 0;
    __syncthreads();
    for (int f0 = 0; f0 < n_frames; f0 += 1024) {
        const int f = f0 + t;
        int v[3], inc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            v[c] = f < n_frames ? counts[3 * f + c] : 0;
            inc[c] = v[c];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, inc[c], d);
                if (lane >= d) inc[c] += o;
            }
            if (lane == 31) wtot[c][w] = inc[c];
        }
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int x = wtot[c][lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += o;
                }
                wtot[c][lane] = x;   // inclusive over warps
            }
        }
        __syncthreads();
        int ex[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) ex[c] = run[c] + (w ? wtot[c][w - 1] : 0) + inc[c] - v[c];
        if (f < n_frames) {
            node_off[f] = ex[0];
            word_off[f] = ex[1];
            valid_off[f] = ex[2];
        }
        __syncthreads();
        if (t < 3) run[t] += wtot[t][31];
        __syncthreads();
    }
    if (t == 0) {
        node_off[n_frames] = run[0];
        word_off[n_frames] = run[1];
        valid_off[n_frames] = run[2];
    }
}

struct CompactArgs {
    const int32_t* kp_off;
    const int32_t *seg_node_id, *seg_node_start, *seg_feat;
    const uint32_t* seg_word;
    const double* seg_val;
    const int32_t *node_off, *word_off, *valid_off;
    uint32_t* bv_word;
    double* bv_value;
    int32_t *fv_node_id, *fv_feat_off, *fv_feat;
    int n_frames;
};

__global__ void __launch_bounds__(kVocThreads) k_voc_compact(CompactArgs A) {
    const int f = blockIdx.x, t = threadIdx.x;
    const int base = A.kp_off[f];
    const int n0 = A.node_off[f], nn = A.node_off[f + 1] - n0;
    const int w0 = A.word_off[f], nw = A.word_off[f + 1] - w0;
    const int v0 = A.valid_off[f], nv = A.valid_off[f + 1] - v0;
    for (int r = t; r < nn; r += kVocThreads) {
        if (A.fv_node_id) A.fv_node_id[n0 + r] = A.seg_node_id[base + r];
        if (A.fv_feat_off) A.fv_feat_off[n0 + r] = v0 + A.seg_node_start[base + r];
    }
    if (A.fv_feat)
        for (int i = t; i < nv; i += kVocThreads) A.fv_feat[v0 + i] = A.seg_feat[base + i];
    for (int r = t; r < nw; r += kVocThreads) {
        if (A.bv_word) A.bv_word[w0 + r] = A.seg_word[base + r];
        if (A.bv_value) A.bv_value[w0 + r] = A.seg_val[base + r];
    }
    if (f == A.n_frames - 1 && t == 0 && A.fv_feat_off) A.fv_feat_off[n0 + nn] = v0 + nv;
}

}  // namespace og

// ------------------------------------------------------------------------------------------------ host side
namespace {

#define OGV_CUDA(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return og_fail(ORBGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
    } while (0)

struct Buf {
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

// The read-only tree in HBM, shared by a vocabulary handle and its forks (orbgpu_vocabulary_fork); freed with the last of them.
struct VocTree {
    int device = 0;
    int32_t *d_child_off = nullptr, *d_child_ids = nullptr, *d_node_word = nullptr;
    uint8_t* d_node_desc = nullptr;
    double* d_node_weight = nullptr;
    ~VocTree() {
        cudaSetDevice(device);
        cudaFree(d_child_off);
        cudaFree(d_child_ids);
        cudaFree(d_node_word);
        cudaFree(d_node_desc);
        cudaFree(d_node_weight);
    }
};

struct orbgpu_vocabulary {
    int device = 0;
    int k = 0, L = 0, scoring = 0, weighting = 0;
    int n_nodes = 0, n_words = 0, max_children = 0;
    std::shared_ptr<VocTree> tree;
    // per-handle execution context: stream + grow-only scratch.  The host-pointer transform holds `mu` from its first copy to
    // its last, so concurrent callers of ONE handle are serialised (never interleaved on the stream); forks run concurrently.
    std::mutex mu;
    cudaStream_t stream = nullptr;
    int last_launches = 0;
    int frame_smem_set = 0;
    Buf b_kp_off, b_desc, b_feat_word, b_feat_node, b_feat_w, b_seg_node_id, b_seg_node_start, b_seg_feat, b_seg_word, b_seg_val, b_counts,
        b_node_off, b_word_off, b_valid_off, b_frame_ws, b_out[7];
};

extern "C" {

int orbgpu_vocabulary_destroy(orbgpu_vocabulary* v) {
    if (!v) return ORBGPU_OK;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    Buf* bs[] = {&v->b_kp_off, &v->b_desc, &v->b_feat_word, &v->b_feat_node, &v->b_feat_w, &v->b_seg_node_id, &v->b_seg_node_start,
                 &v->b_seg_feat, &v->b_seg_word, &v->b_seg_val, &v->b_counts, &v->b_node_off, &v->b_word_off, &v->b_valid_off, &v->b_frame_ws};
    for (Buf* b : bs) b->release();
    for (Buf& b : v->b_out) b.release();
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
    return ORBGPU_OK;
}

int orbgpu_vocabulary_create(orbgpu_vocabulary** out, int device, int k, int L, int scoring, int weighting, int n_records,
                             const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    if (n_records < 1 || !parent || !is_leaf || !desc || !weight) return og_fail(ORBGPU_ERR_ARG, "vocabulary: empty or null arrays");
    if (scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) return og_fail(ORBGPU_ERR_ARG, "vocabulary: bad scoring / weighting type");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return og_fail(ORBGPU_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return og_fail(ORBGPU_ERR_ARG, "device index out of range");
    // children lists in push_back order = ascending record order (TemplatedVocabulary.h:1392); words in leaf order (:1407-1414)
    const int n = n_records + 1;
    std::vector<int32_t> cnt(n + 1, 0), word(n, -1);
    for (int r = 0; r < n_records; ++r) {
        if (parent[r] < 0 || parent[r] > r) return og_fail(ORBGPU_ERR_ARG, "vocabulary: a parent id must precede its child");
        ++cnt[parent[r] + 1];
    }
    int max_children = 0;
    for (int i = 0; i < n; ++i) max_children = std::max(max_children, cnt[i + 1]);
    if (max_children > 32) return og_fail(ORBGPU_ERR_ARG, "vocabulary: more than 32 children per node");
    for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    std::vector<int32_t> ids(n_records), fill(cnt.begin(), cnt.end() - 1);
    int n_words = 0;
    for (int r = 0; r < n_records; ++r) {
        ids[fill[parent[r]]++] = r + 1;
        if (is_leaf[r]) word[r + 1] = n_words++;
    }
    for (int r = 0; r < n_records; ++r)
        if ((cnt[r + 2] - cnt[r + 1] == 0) != (is_leaf[r] != 0))
            return og_fail(ORBGPU_ERR_ARG, "vocabulary: is_leaf disagrees with the tree (a leaf with children or an inner node without)");
    if (cnt[1] == 0) return og_fail(ORBGPU_ERR_ARG, "vocabulary: the root has no children");
    std::vector<uint8_t> nd((size_t)n * 32, 0);
    std::vector<double> nw(n, 0.0);
    std::copy(desc, desc + (size_t)n_records * 32, nd.begin() + 32);
    std::copy(weight, weight + n_records, nw.begin() + 1);

    OGV_CUDA(cudaSetDevice(device));
    orbgpu_vocabulary* v = new orbgpu_vocabulary();
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->n_nodes = n; v->n_words = n_words; v->max_children = max_children;
    v->tree = std::make_shared<VocTree>();
    VocTree* T = v->tree.get();
    T->device = device;
    cudaError_t ce = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&T->d_child_off, (size_t)(n + 1) * 4);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&T->d_child_ids, (size_t)n_records * 4);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&T->d_node_word, (size_t)n * 4);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&T->d_node_desc, (size_t)n * 32);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&T->d_node_weight, (size_t)n * 8);
    if (ce == cudaSuccess) ce = cudaMemcpy(T->d_child_off, cnt.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(T->d_child_ids, ids.data(), (size_t)n_records * 4, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(T->d_node_word, word.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(T->d_node_desc, nd.data(), (size_t)n * 32, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(T->d_node_weight, nw.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        std::string msg = std::string("vocabulary upload failed: ") + cudaGetErrorString(ce);
        orbgpu_vocabulary_destroy(v);
        return og_fail(ORBGPU_ERR_CUDA, msg);
    }
    *out = v;
    return ORBGPU_OK;
}

int orbgpu_vocabulary_fork(const orbgpu_vocabulary* v, orbgpu_vocabulary** out) {
    if (!v || !out) return og_fail(ORBGPU_ERR_ARG, "null argument");
    *out = nullptr;
    OGV_CUDA(cudaSetDevice(v->device));
    orbgpu_vocabulary* f = new orbgpu_vocabulary();
    f->device = v->device; f->k = v->k; f->L = v->L; f->scoring = v->scoring; f->weighting = v->weighting;
    f->n_nodes = v->n_nodes; f->n_words = v->n_words; f->max_children = v->max_children;
    f->tree = v->tree;
    const cudaError_t ce = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        delete f;
        return og_fail(ORBGPU_ERR_CUDA, std::string("vocabulary fork: ") + cudaGetErrorString(ce));
    }
    *out = f;
    return ORBGPU_OK;
}

int orbgpu_vocabulary_info(const orbgpu_vocabulary* v, int* n_nodes, int* n_words) {
    if (!v) return og_fail(ORBGPU_ERR_ARG, "null vocabulary");
    if (n_nodes) *n_nodes = v->n_nodes;
    if (n_words) *n_words = v->n_words;
    return ORBGPU_OK;
}

int orbgpu_vocabulary_sync(orbgpu_vocabulary* v) {
    if (!v) return og_fail(ORBGPU_ERR_ARG, "null vocabulary");
    OGV_CUDA(cudaSetDevice(v->device));
    OGV_CUDA(cudaStreamSynchronize(v->stream));
    return ORBGPU_OK;
}

int orbgpu_vocabulary_last_launches(const orbgpu_vocabulary* v) { return v ? v->last_launches : 0; }

int orbgpu_vocabulary_stream(orbgpu_vocabulary* v, void** stream_out) {
    if (!v || !stream_out) return og_fail(ORBGPU_ERR_ARG, "null argument");
    *stream_out = (void*)v->stream;
    return ORBGPU_OK;
}

int orbgpu_bow_transform_dev(orbgpu_vocabulary* v, int n_frames, const int32_t* kp_off_dev, int n_features, int max_per_frame,
                             const uint8_t* desc_dev, int levelsup, int32_t* bv_off, uint32_t* bv_word, double* bv_value,
                             int32_t* fv_node_off, int32_t* fv_node_id, int32_t* fv_feat_off, int32_t* fv_feat,
                             uint32_t* word_of_feature, uint32_t* node_of_feature) {
    OG_NVTX("orbgpu_bow_transform_dev");
    if (!v) return og_fail(ORBGPU_ERR_ARG, "null vocabulary");
    if (n_frames < 0 || n_features < 0 || (n_frames > 0 && !kp_off_dev) || (n_features > 0 && !desc_dev))
        return og_fail(ORBGPU_ERR_ARG, "bow_transform: bad arguments");
    OGV_CUDA(cudaSetDevice(v->device));
    v->last_launches = 0;
    if (n_frames == 0) return ORBGPU_OK;
    cudaStream_t st = v->stream;
    const size_t nf = (size_t)std::max(n_features, 1);
    void *fw, *fn, *fwt, *sni, *sns, *sf, *sw, *sv, *cn, *no, *wo, *vo;
    if (!word_of_feature) { OGV_CUDA(v->b_feat_word.grab(nf * 4, &fw)); } else fw = word_of_feature;
    if (!node_of_feature) { OGV_CUDA(v->b_feat_node.grab(nf * 4, &fn)); } else fn = node_of_feature;
    OGV_CUDA(v->b_feat_w.grab(nf * 8, &fwt));
    OGV_CUDA(v->b_seg_node_id.grab(nf * 4, &sni));
    OGV_CUDA(v->b_seg_node_start.grab(nf * 4, &sns));
    OGV_CUDA(v->b_seg_feat.grab(nf * 4, &sf));
    OGV_CUDA(v->b_seg_word.grab(nf * 4, &sw));
    OGV_CUDA(v->b_seg_val.grab(nf * 8, &sv));
    OGV_CUDA(v->b_counts.grab((size_t)n_frames * 12, &cn));
    if (!fv_node_off) { OGV_CUDA(v->b_node_off.grab((size_t)(n_frames + 1) * 4, &no)); } else no = fv_node_off;
    if (!bv_off) { OGV_CUDA(v->b_word_off.grab((size_t)(n_frames + 1) * 4, &wo)); } else wo = bv_off;
    OGV_CUDA(v->b_valid_off.grab((size_t)(n_frames + 1) * 4, &vo));

    const VocTree* T = v->tree.get();
    og::VocView V = {T->d_child_off, T->d_child_ids, T->d_node_desc, T->d_node_word, T->d_node_weight};
    const int nid_level = v->L - levelsup;
    if (n_features > 0) {
        if (v->max_children <= 16) {
            const int blocks = (int)(((size_t)n_features * 16 + og::kVocThreads - 1) / og::kVocThreads);
            og::k_voc_descend<16><<<blocks, og::kVocThreads, 0, st>>>(V, desc_dev, n_features, nid_level, (uint32_t*)fw, (uint32_t*)fn, (double*)fwt);
        } else {
            const int blocks = (int)(((size_t)n_features * 32 + og::kVocThreads - 1) / og::kVocThreads);
            og::k_voc_descend<32><<<blocks, og::kVocThreads, 0, st>>>(V, desc_dev, n_features, nid_level, (uint32_t*)fw, (uint32_t*)fn, (double*)fwt);
        }
        ++v->last_launches;
    }
    int p_max = 1;
    while (p_max < max_per_frame) p_max <<= 1;
    size_t smem = (size_t)p_max * 16 + (size_t)(p_max + 1) * 4;
    void* gws = nullptr;
    size_t gws_stride = 0;
    if (p_max > og::kVocSmemFrame) {
        // a frame too large for the shared-memory sort (the reference's transform has no size limit): the same kernel on a
        // block-private workspace in HBM - slower per frame, only taken by batches that contain such a frame
        gws_stride = (smem + 255) & ~(size_t)255;
        OGV_CUDA(v->b_frame_ws.grab(gws_stride * (size_t)n_frames, &gws));
        smem = 0;
    } else if ((int)smem > v->frame_smem_set && smem > 48 * 1024) {
        OGV_CUDA(cudaFuncSetAttribute(og::k_voc_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, og::kVocSmemFrame * 16 + (og::kVocSmemFrame + 1) * 4));
        v->frame_smem_set = og::kVocSmemFrame * 16 + (og::kVocSmemFrame + 1) * 4;
    }
    og::FrameArgs A = {kp_off_dev, (const uint32_t*)fw, (const uint32_t*)fn, (const double*)fwt, (int32_t*)sni, (int32_t*)sns, (int32_t*)sf,
                       (uint32_t*)sw, (double*)sv, (int32_t*)cn, (v->weighting == 0 || v->weighting == 1) ? 1 : 0,
                       v->scoring == 5 ? 0 : (v->scoring == 1 ? 2 : 1), (unsigned char*)gws, gws_stride};
    og::k_voc_frame<<<n_frames, og::kVocThreads, smem, st>>>(A, p_max);
    og::k_voc_offsets<<<1, 1024, 0, st>>>((const int32_t*)cn, n_frames, (int32_t*)no, (int32_t*)wo, (int32_t*)vo);
    og::CompactArgs C = {kp_off_dev, (const int32_t*)sni, (const int32_t*)sns, (const int32_t*)sf, (const uint32_t*)sw, (const double*)sv,
                         (const int32_t*)no, (const int32_t*)wo, (const int32_t*)vo, bv_word, bv_value, fv_node_id, fv_feat_off, fv_feat, n_frames};
    og::k_voc_compact<<<n_frames, og::kVocThreads, 0, st>>>(C);
    v->last_launches += 3;
    OGV_CUDA(cudaGetLastError());
    return ORBGPU_OK;
}

int orbgpu_bow_transform(orbgpu_vocabulary* v, int n_frames, const int32_t* kp_off, const uint8_t* desc, int levelsup,
                         int32_t* bv_off, uint32_t* bv_word, double* bv_value, int32_t* fv_node_off, int32_t* fv_node_id,
                         int32_t* fv_feat_off, int32_t* fv_feat, uint32_t* word_of_feature, uint32_t* node_of_feature) {
    OG_NVTX("orbgpu_bow_transform");
    if (!v) return og_fail(ORBGPU_ERR_ARG, "null vocabulary");
    if (n_frames < 0 || (n_frames > 0 && !kp_off)) return og_fail(ORBGPU_ERR_ARG, "bow_transform: bad arguments");
    if (n_frames == 0) return ORBGPU_OK;
    int max_per = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (kp_off[f + 1] < kp_off[f]) return og_fail(ORBGPU_ERR_ARG, "bow_transform: kp_off must be non-decreasing");
        max_per = std::max(max_per, kp_off[f + 1] - kp_off[f]);
    }
    const int n = kp_off[n_frames] - kp_off[0];
    if (kp_off[0] != 0) return og_fail(ORBGPU_ERR_ARG, "bow_transform: kp_off[0] must be 0");
    if (n > 0 && !desc) return og_fail(ORBGPU_ERR_ARG, "bow_transform: null descriptors");
    std::lock_guard<std::mutex> lock(v->mu);   // one host call at a time per handle (ORBVocabulary is shared between threads)
    OGV_CUDA(cudaSetDevice(v->device));
    cudaStream_t st = v->stream;
    void *dko, *dd;
    OGV_CUDA(v->b_kp_off.grab((size_t)(n_frames + 1) * 4, &dko));
    OGV_CUDA(v->b_desc.grab((size_t)std::max(n, 1) * 32, &dd));
    OGV_CUDA(cudaMemcpyAsync(dko, kp_off, (size_t)(n_frames + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n > 0) OGV_CUDA(cudaMemcpyAsync(dd, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    const size_t nn = (size_t)std::max(n, 1);
    void* o[9];
    const size_t bytes[9] = {(size_t)(n_frames + 1) * 4, nn * 4, nn * 8, (size_t)(n_frames + 1) * 4, nn * 4, (nn + 1) * 4, nn * 4, nn * 4, nn * 4};
    void* host[9] = {bv_off, bv_word, bv_value, fv_node_off, fv_node_id, fv_feat_off, fv_feat, word_of_feature, node_of_feature};
    static_assert(sizeof(v->b_out) / sizeof(v->b_out[0]) == 7, "seven pooled outputs + two per-feature arrays");
    for (int i = 0; i < 7; ++i) {
        o[i] = nullptr;
        if (host[i] || i == 0 || i == 3) OGV_CUDA(v->b_out[i].grab(bytes[i], &o[i]));   // the offset arrays are always needed for the copies back
    }
    o[7] = o[8] = nullptr;
    if (word_of_feature) OGV_CUDA(v->b_feat_word.grab(bytes[7], &o[7]));
    if (node_of_feature) OGV_CUDA(v->b_feat_node.grab(bytes[8], &o[8]));
    int rc = orbgpu_bow_transform_dev(v, n_frames, (const int32_t*)dko, n, max_per, (const uint8_t*)dd, levelsup, (int32_t*)o[0], (uint32_t*)o[1],
                                      (double*)o[2], (int32_t*)o[3], (int32_t*)o[4], (int32_t*)o[5], (int32_t*)o[6], (uint32_t*)o[7], (uint32_t*)o[8]);
    if (rc != ORBGPU_OK) return rc;
    // sizes of the compact arrays
    int32_t tot_words = 0, tot_nodes = 0;
    OGV_CUDA(cudaMemcpyAsync(&tot_words, (int32_t*)o[0] + n_frames, 4, cudaMemcpyDeviceToHost, st));
    OGV_CUDA(cudaMemcpyAsync(&tot_nodes, (int32_t*)o[3] + n_frames, 4, cudaMemcpyDeviceToHost, st));
    OGV_CUDA(cudaStreamSynchronize(st));
    int32_t tot_valid = 0;
    if (fv_feat_off || fv_feat) {
        if (o[5]) {
            OGV_CUDA(cudaMemcpyAsync(&tot_valid, (int32_t*)o[5] + tot_nodes, 4, cudaMemcpyDeviceToHost, st));
            OGV_CUDA(cudaStreamSynchronize(st));
        } else {
            tot_valid = n;
        }
    }
    const size_t used[9] = {(size_t)(n_frames + 1) * 4, (size_t)tot_words * 4, (size_t)tot_words * 8, (size_t)(n_frames + 1) * 4, (size_t)tot_nodes * 4,
                            (size_t)(tot_nodes + 1) * 4, (size_t)tot_valid * 4, (size_t)n * 4, (size_t)n * 4};
    for (int i = 0; i < 9; ++i)
        if (host[i] && o[i] && used[i]) OGV_CUDA(cudaMemcpyAsync(host[i], o[i], used[i], cudaMemcpyDeviceToHost, st));
    OGV_CUDA(cudaStreamSynchronize(st));
    return ORBGPU_OK;
}

}  // extern "C"

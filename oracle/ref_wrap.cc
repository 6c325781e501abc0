// ref_wrap — TEST INFRASTRUCTURE.  C-ABI wrapper around the reference's own ORBextractor, compiled
// from the source where it lies (-I$(REF)/src -I$(REF)/include, see oracle/Makefile) against the cv::
// shim.  The reference translation unit is pulled in unmodified with #include; nothing is copied.
//
//   ORBREF_STABLE_TIEBREAK   defined: the one documented patch — the node sort at ORBextractor.cc:684
//                            (default pair<int,ExtractorNode*> ordering => ties by heap address) becomes a
//                            stable sort on the count only.  Output: oracle/_ref/liborbref.so (parity).
//                            undefined: verbatim behaviour -> oracle/_ref/liborbref_verbatim.so, used
//                            only to report the reference's own nondeterminism envelope.
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#include <chrono>
#include <thread>
#include <utility>
#include <vector>

#ifdef ORBREF_STABLE_TIEBREAK
namespace orbref_patch {
template <class It> inline void node_sort(It first, It last) {
    typedef typename std::iterator_traits<It>::value_type V;
    std::stable_sort(first, last, [](const V& a, const V& b) { return a.first < b.first; });
}
}  // namespace orbref_patch
#define sort orbref_patch::node_sort
#endif

#include "ORBextractor.cc"   // the reference translation unit, resolved through -I$(REF)/src

#ifdef ORBREF_STABLE_TIEBREAK
#undef sort
#endif

namespace {
struct Access : public ORB_SLAM2::ORBextractor {
    Access(int a, float b, int c, int d, int e) : ORB_SLAM2::ORBextractor(a, b, c, d, e) {}
    using ORB_SLAM2::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM2::ORBextractor::umax;
    using ORB_SLAM2::ORBextractor::DistributeOctTree;
};
}  // namespace

extern "C" {

void* orbref_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Access(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbref_destroy(void* h) { delete (Access*)h; }

void orbref_tables(void* h, float* scales, int* featPerLevel, int* umax) {
    Access* e = (Access*)h;
    const int n = e->GetLevels();
    std::vector<float> s = e->GetScaleFactors(), is = e->GetInverseScaleFactors(), g = e->GetScaleSigmaSquares(), ig = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; ++i) {
        scales[i] = s[i]; scales[n + i] = is[i]; scales[2 * n + i] = g[i]; scales[3 * n + i] = ig[i];
        featPerLevel[i] = e->mnFeaturesPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax[i] = e->umax[i];
}

// kp_out: cv::KeyPoint layout (28 B each); returns the number of keypoints produced
int orbref_extract(void* h, const uint8_t* img, int w, int hgt, int stride, void* kp_out, int cap, uint8_t* desc_out) {
    Access* e = (Access*)h;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    (*e)(image, cv::Mat(), kps, desc);
    const int n = (int)kps.size();
    const int m = n < cap ? n : cap;
    static_assert(sizeof(cv::KeyPoint) == 28, "KeyPoint layout");
    if (kp_out && m) memcpy(kp_out, kps.data(), (size_t)m * sizeof(cv::KeyPoint));
    if (desc_out && m) for (int i = 0; i < m; ++i) memcpy(desc_out + (size_t)i * 32, desc.ptr(i), 32);
    return n;
}

void orbref_level_dims(void* h, int level, int* w, int* hgt) {
    Access* e = (Access*)h;
    *w = e->mvImagePyramid[level].cols; *hgt = e->mvImagePyramid[level].rows;
}
void orbref_get_level(void* h, int level, int bordered, uint8_t* out) {
    Access* e = (Access*)h;
    const cv::Mat& m = e->mvImagePyramid[level];
    if (bordered) {
        const int W = m.cols + 38, H = m.rows + 38;
        const uchar* base = m.data - 19 * (size_t)m.step - 19;
        for (int y = 0; y < H; ++y) memcpy(out + (size_t)y * W, base + (size_t)y * (size_t)m.step, W);
    } else {
        for (int y = 0; y < m.rows; ++y) memcpy(out + (size_t)y * m.cols, m.ptr(y), m.cols);
    }
}

int orbref_octree(const void* cand, int n, int minX, int maxX, int minY, int maxY, int N, void* out, int cap) {
    Access e(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> K((const cv::KeyPoint*)cand, (const cv::KeyPoint*)cand + n);
    int level = 0;
    std::vector<cv::KeyPoint> r = e.DistributeOctTree(K, minX, maxX, minY, maxY, N, level);
    const int m = (int)r.size() < cap ? (int)r.size() : cap;
    if (out && m) memcpy(out, r.data(), (size_t)m * sizeof(cv::KeyPoint));
    return (int)r.size();
}

// CPU baseline: `nframes` frames (contiguous, w*hgt each) processed `reps` times by `nthreads`
// std::threads, one extractor instance per thread, one frame at a time per thread.  Returns seconds of
// wall time for the best repetition; *total_kp receives the keypoint count of one pass.
double orbref_bench(const uint8_t* frames, int nframes, int w, int hgt, int nfeatures, float scaleFactor, int nlevels,
                    int iniTh, int minTh, int nthreads, int reps, long* total_kp) {
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        std::vector<long> counts(nthreads, 0);
        std::vector<std::thread> th;
        auto t0 = std::chrono::steady_clock::now();
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                ORB_SLAM2::ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
                std::vector<cv::KeyPoint> kps;
                for (int f = t; f < nframes; f += nthreads) {
                    cv::Mat image(hgt, w, CV_8UC1, (void*)(frames + (size_t)f * w * hgt), (size_t)w);
                    cv::Mat desc;
                    ex(image, cv::Mat(), kps, desc);
                    counts[t] += (long)kps.size();
                }
            });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best) best = s;
        if (total_kp) { long c = 0; for (long v : counts) c += v; *total_kp = c; }
    }
    return best;
}

}  // extern "C"

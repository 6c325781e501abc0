// shell_test.cc — exercises the C++ drop-in shells (csrc/host) the way the reference's callers do, and checks every
// result against the CPU oracle (oracle/_build/liborboracle.so).  TEST INFRASTRUCTURE.
//   exit 0: all comparisons passed on the GPU        exit 3: no CUDA device, the shells failed loudly (no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ORBVocabulary.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "orbgpu.h"

using namespace ORB_SLAM2;

// ---- oracle entry points (oracle/orb_oracle.cc, oracle/match_oracle.cc) ----
extern "C" {
void* orbo_create(int, float, int, int, int);
void orbo_destroy(void*);
int orbo_extract(void*, const uint8_t*, int, int, int, void*, int, uint8_t*);
int orbo_get_level(void*, int, int, uint8_t*);
void orbs_stereo_matches(const orbgpu_keypoint*, const uint8_t*, int, const orbgpu_keypoint*, const uint8_t*, int, const uint8_t* const*,
                         const uint8_t* const*, const int32_t*, const int32_t*, const float*, const float*, int, float, float, float*, float*,
                         int32_t*);
void orbm_search_by_projection(const orbgpu_frame_set*, const orbgpu_mappoint_set*, const float*, int, float, float, int32_t*, int32_t*,
                               int32_t*, int32_t*, int32_t*);
void orbm_search_windowed(const orbgpu_frame_set*, const orbgpu_window_query_set*, int, int, int, int32_t*, int32_t*, int32_t*, int32_t*);
void orbm_search_for_triangulation(const orbgpu_frame_set*, const orbgpu_frame_set*, int, const int32_t*, const int32_t*, const float*,
                                   const float*, const float*, const float*, int, int, int, const int64_t*, int32_t*, int32_t*, int32_t*);
void orbm_search_by_bow(const orbgpu_frame_set*, const orbgpu_frame_set*, int, const int32_t*, const int32_t*, float, int, int, int, int,
                        const int64_t*, int32_t*, int32_t*, int32_t*);
void cvl_gemm3_f32(const float*, const float*, const float*, float*);
void cvl_gemm3t_neg_f32(const float*, const float*, float*);
double cvl_norm3_f32(const float*);
double cvl_dot3_f32(const float*, const float*);
void orbm_search_window_best(const orbgpu_frame_set*, const orbgpu_window_query_set*, const float*, int, int32_t*, int32_t*);
void orbm_search_for_initialization(const orbgpu_frame_set*, const orbgpu_window_query_set*, float, int, int32_t*, int32_t*);
void* orbo_voc_create(int, int, int, int, int, const int32_t*, const uint8_t*, const uint8_t*, const double*);
void orbo_voc_free(void*);
int orbo_voc_transform(void*, const uint8_t*, int, int, int*, uint32_t*, double*, int*, uint32_t*, int32_t*, uint32_t*, uint32_t*, uint32_t*);
}

float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;

static uint32_t g_seed = 12345;
static uint32_t rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }
static float frand(float a, float b) { return a + (b - a) * (rnd() % 100000) / 100000.0f; }

static int fails = 0;
#define EXPECT(c, ...) do { if (!(c)) { ++fails; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

static cv::Mat synth_image(int w, int h) {
    cv::Mat im(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) im.at<uchar>(y, x) = 128;
    for (int k = 0; k < w * h / 700; ++k) {
        const int x0 = rnd() % w, y0 = rnd() % h, rw = 4 + rnd() % 50, rh = 4 + rnd() % 50, g = rnd() % 256;
        for (int y = y0; y < y0 + rh && y < h; ++y) for (int x = x0; x < x0 + rw && x < w; ++x) im.at<uchar>(y, x) = (uchar)g;
    }
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        int v = im.at<uchar>(y, x) + (int)(rnd() % 7) - 3;
        im.at<uchar>(y, x) = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    return im;
}

// what Frame::ExtractORB does (Frame.cc:249-255)
struct MiniFrame {
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    void ExtractORB(ORBextractor* ex, const cv::Mat& im) { (*ex)(im, cv::Mat(), mvKeys, mDescriptors); }
};

static void test_extractor() {
    const int W = 752, H = 480;
    ORBextractor ex(1200, 1.2f, 8, 20, 7);   // Tracking.cc:120-126 style construction
    EXPECT(ex.GetLevels() == 8 && std::fabs(ex.GetScaleFactor() - 1.2f) < 1e-6f, "getters");
    EXPECT(ex.GetScaleFactors().size() == 8 && ex.GetInverseScaleSigmaSquares().size() == 8, "tables");
    void* o = orbo_create(1200, 1.2f, 8, 20, 7);
    for (int it = 0; it < 3; ++it) {
        cv::Mat im = synth_image(W, H);
        MiniFrame f;
        f.ExtractORB(&ex, im);
        std::vector<cv::KeyPoint> ok(1200 + 100);
        std::vector<uint8_t> od(ok.size() * 32);
        const int n = orbo_extract(o, im.ptr(0), W, H, (int)im.step, ok.data(), (int)ok.size(), od.data());
        EXPECT(n == (int)f.mvKeys.size() && f.mDescriptors.rows == n, "keypoint count %d vs oracle %d", (int)f.mvKeys.size(), n);
        int bad_desc = 0;
        for (int i = 0; i < n && i < (int)f.mvKeys.size(); ++i) {
            const cv::KeyPoint &a = f.mvKeys[i], &b = ok[i];
            EXPECT(a.pt.x == b.pt.x && a.pt.y == b.pt.y && a.octave == b.octave && a.response == b.response && a.size == b.size &&
                   a.class_id == b.class_id && std::fabs(a.angle - b.angle) <= 1e-3f, "keypoint %d differs", i);
            bad_desc += std::memcmp(f.mDescriptors.ptr(i), &od[(size_t)i * 32], 32) != 0;
        }
        EXPECT(bad_desc <= 1, "%d descriptor rows differ", bad_desc);
        // mvImagePyramid: interior views of bordered buffers, as ComputePyramid leaves them
        for (int l = 0; l < 8; ++l) {
            const cv::Mat& lv = ex.mvImagePyramid[l];
            std::vector<uint8_t> ol((size_t)lv.rows * lv.cols);
            orbo_get_level(o, l, 0, ol.data());
            int diff = 0;
            for (int y = 0; y < lv.rows; ++y) diff += std::memcmp(lv.ptr(y), &ol[(size_t)y * lv.cols], lv.cols) != 0;
            EXPECT(diff == 0, "pyramid level %d: %d rows differ", l, diff);
        }
    }
    // empty image: silent return, outputs untouched (reference :1046-1047)
    MiniFrame e;
    e.mvKeys.resize(3);
    e.ExtractORB(&ex, cv::Mat());
    EXPECT(e.mvKeys.size() == 3, "empty image must leave the outputs alone");
    orbo_destroy(o);
}

// what the stereo Frame constructor does (Frame.cc:78-100): both extractors, then ComputeStereoMatches
static void test_stereo() {
    const int W = 752, H = 480;
    ORBextractor exL(1200, 1.2f, 8, 20, 7), exR(1200, 1.2f, 8, 20, 7);
    cv::Mat left = synth_image(W, H), right(H, W, CV_8UC1);
    for (int y = 0; y < H; ++y) {
        const int d = 6 + (y / 60) % 9;
        for (int x = 0; x < W; ++x) right.at<uchar>(y, x) = left.at<uchar>(y, (x + d) % W);
    }
    MiniFrame fl, fr;
    fl.ExtractORB(&exL, left);
    fr.ExtractORB(&exR, right);
    const int N = (int)fl.mvKeys.size(), Nr = (int)fr.mvKeys.size();
    std::vector<float> mvuRight, mvDepth;
    const float mb = 0.11f, mbf = 47.9f;
    ORBextractor::ComputeStereoMatches(&exL, &exR, N, mb, mbf, mvuRight, mvDepth);
    // oracle on the same key points / descriptors / pyramids (copied out of the ROI views into packed buffers)
    std::vector<std::vector<uint8_t> > pl(8), pr(8);
    const uint8_t *ppl[8], *ppr[8];
    int32_t lw[8], lh[8];
    for (int l = 0; l < 8; ++l) {
        const cv::Mat &a = exL.mvImagePyramid[l], &b = exR.mvImagePyramid[l];
        lw[l] = a.cols; lh[l] = a.rows;
        pl[l].resize((size_t)a.rows * a.cols); pr[l].resize((size_t)a.rows * a.cols);
        for (int y = 0; y < a.rows; ++y) { std::memcpy(&pl[l][(size_t)y * a.cols], a.ptr(y), a.cols); std::memcpy(&pr[l][(size_t)y * a.cols], b.ptr(y), a.cols); }
        ppl[l] = pl[l].data(); ppr[l] = pr[l].data();
    }
    std::vector<uint8_t> dl((size_t)N * 32), dr((size_t)Nr * 32);
    for (int i = 0; i < N; ++i) std::memcpy(&dl[(size_t)i * 32], fl.mDescriptors.ptr(i), 32);
    for (int i = 0; i < Nr; ++i) std::memcpy(&dr[(size_t)i * 32], fr.mDescriptors.ptr(i), 32);
    std::vector<float> sc = exL.GetScaleFactors(), isc = exL.GetInverseScaleFactors(), eu(N), ed(N);
    orbs_stereo_matches((const orbgpu_keypoint*)fl.mvKeys.data(), dl.data(), N, (const orbgpu_keypoint*)fr.mvKeys.data(), dr.data(), Nr, ppl, ppr, lw,
                        lh, sc.data(), isc.data(), 8, mb, mbf, eu.data(), ed.data(), nullptr);
    int matched = 0;
    for (int i = 0; i < N; ++i) matched += eu[i] >= 0;
    EXPECT(mvuRight == eu && mvDepth == ed && matched > N / 4, "ComputeStereoMatches: %d of %d matched, vectors %s", matched, N,
           mvuRight == eu ? "equal" : "differ");
}

// The stereo Frame constructor runs the left and right extractor on two threads (Frame.cc:78-81), and matchers run on the
// Tracking / LocalMapping / LoopClosing threads: concurrent use of distinct handles must give the serial results.
static void test_two_threads() {
    const int W = 640, H = 480;
    cv::Mat imA = synth_image(W, H), imB = synth_image(W, H);
    ORBextractor exA(1000, 1.2f, 8, 20, 7), exB(1000, 1.2f, 8, 20, 7);
    MiniFrame refA, refB;
    refA.ExtractORB(&exA, imA);
    refB.ExtractORB(&exB, imB);
    int bad = 0;
    for (int round = 0; round < 20; ++round) {
        MiniFrame a, b;
        std::thread tl(&MiniFrame::ExtractORB, &a, &exA, imA);
        std::thread tr(&MiniFrame::ExtractORB, &b, &exB, imB);
        tl.join();
        tr.join();
        bad += a.mvKeys.size() != refA.mvKeys.size() || b.mvKeys.size() != refB.mvKeys.size();
        for (size_t i = 0; i < a.mvKeys.size() && i < refA.mvKeys.size(); ++i)
            bad += std::memcmp(&a.mvKeys[i], &refA.mvKeys[i], sizeof(cv::KeyPoint)) != 0 || std::memcmp(a.mDescriptors.ptr((int)i), refA.mDescriptors.ptr((int)i), 32) != 0;
        for (size_t i = 0; i < b.mvKeys.size() && i < refB.mvKeys.size(); ++i)
            bad += std::memcmp(&b.mvKeys[i], &refB.mvKeys[i], sizeof(cv::KeyPoint)) != 0 || std::memcmp(b.mDescriptors.ptr((int)i), refB.mDescriptors.ptr((int)i), 32) != 0;
    }
    EXPECT(bad == 0 && refA.mvKeys.size() > 500, "two extractor threads: %d differences from the serial results", bad);
}

// ---- matcher fixtures ----
static std::vector<float> g_sf, g_s2;
static cv::Mat random_desc(int n) {
    cv::Mat d(n, 32, CV_8UC1);
    for (int i = 0; i < n; ++i) for (int b = 0; b < 32; ++b) d.at<uchar>(i, b) = (uchar)(rnd() & 255);
    return d;
}
static void flip(uchar* row, int nbits) { for (int k = 0; k < nbits; ++k) { int b = rnd() % 256; row[b >> 3] ^= (uchar)(1 << (b & 7)); } }
static std::vector<cv::KeyPoint> random_keys(int n, int W, int H) {
    std::vector<cv::KeyPoint> k(n);
    for (int i = 0; i < n; ++i) {
        const int o = rnd() % 8;
        k[i] = cv::KeyPoint(std::floor(frand(20, W / g_sf[o] - 20)) * g_sf[o], std::floor(frand(20, H / g_sf[o] - 20)) * g_sf[o],
                            31 * g_sf[o], frand(0, 360), (float)(rnd() % 100), o, -1);
    }
    return k;
}
struct Flat {   // independent flattening for the oracle
    int32_t kp_off[2], node_off[2];
    std::vector<int32_t> node_id, feat_off, feat;
    std::vector<uint8_t> desc, flags;
    orbgpu_frame_set s;
    Flat(const std::vector<cv::KeyPoint>& keys, const cv::Mat& d, const DBoW2::FeatureVector* fv) {
        std::memset(&s, 0, sizeof(s));
        kp_off[0] = 0; kp_off[1] = (int)keys.size();
        s.n_frames = 1; s.kp_off = kp_off; s.keys_un = (const orbgpu_keypoint*)keys.data();
        desc.resize(keys.size() * 32);
        for (size_t i = 0; i < keys.size(); ++i) std::memcpy(&desc[i * 32], d.ptr((int)i), 32);
        s.desc = desc.data();
        flags.assign(keys.size(), 0);
        s.kp_flags = flags.data();
        if (fv) {
            node_off[0] = 0; feat_off.push_back(0);
            for (auto& kv : *fv) { node_id.push_back((int)kv.first); for (unsigned v : kv.second) feat.push_back((int)v); feat_off.push_back((int)feat.size()); }
            node_off[1] = (int)node_id.size();
            s.fv_node_off = node_off; s.fv_node_id = node_id.data(); s.fv_feat_off = feat_off.data(); s.fv_feat = feat.data();
        }
    }
};

static void test_matcher() {
    const int W = 640, H = 480, N = 900;
    g_sf.assign(8, 1.f); g_s2.assign(8, 1.f);
    for (int i = 1; i < 8; ++i) { g_sf[i] = (float)(g_sf[i - 1] * (double)1.2f); g_s2[i] = g_sf[i] * g_sf[i]; }
    Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = W; Frame::mnMaxY = H;
    Frame::mfGridElementWidthInv = 64.0f / W; Frame::mfGridElementHeightInv = 48.0f / H;

    // two key frames; the second re-observes ~60 % of the first
    std::vector<cv::KeyPoint> k1 = random_keys(N, W, H), k2 = random_keys(N, W, H);
    cv::Mat d1 = random_desc(N), d2 = random_desc(N);
    DBoW2::FeatureVector fv1, fv2;
    std::vector<int> node1(N), node2(N);
    for (int i = 0; i < N; ++i) node1[i] = rnd() % 40, node2[i] = rnd() % 40;
    for (int j = 0; j < N; ++j)
        if (rnd() % 10 < 6) {
            const int i = rnd() % N;
            std::memcpy(d2.ptr(j), d1.ptr(i), 32);
            flip(d2.ptr(j), 12);
            k2[j] = k1[i];
            k2[j].pt.x += frand(-1.5f, 1.5f);
            k2[j].angle = std::fmod(k1[i].angle + frand(-10, 10) + 360.f, 360.f);
            node2[j] = node1[i];
        }
    for (int i = 0; i < N; ++i) { fv1[node1[i]].push_back(i); fv2[node2[i]].push_back(i); }
    std::vector<float> ur1(N, -1.f), ur2(N, -1.f);
    for (int i = 0; i < N; ++i) { if (rnd() % 4 == 0) ur1[i] = k1[i].pt.x - 5; if (rnd() % 4 == 0) ur2[i] = k2[i].pt.x - 5; }
    KeyFrame kf1(k1, ur1, d1, g_sf, g_s2, 517.3f, 516.5f, 318.6f, 255.3f), kf2(k2, ur2, d2, g_sf, g_s2, 517.3f, 516.5f, 318.6f, 255.3f);
    kf1.mFeatVec = fv1; kf2.mFeatVec = fv2;
    std::vector<MapPoint> pool(2 * N);
    for (int i = 0; i < N; ++i) {
        if (rnd() % 10 < 7) { kf1.mvpMapPoints[i] = &pool[i]; pool[i].mbBad = rnd() % 20 == 0; pool[i].nObs = 1 + rnd() % 3; }
        if (rnd() % 10 < 7) { kf2.mvpMapPoints[i] = &pool[N + i]; pool[N + i].mbBad = rnd() % 20 == 0; pool[N + i].nObs = 1; }
    }

    // --- SearchByBoW(KF, KF) ---
    {
        ORBmatcher m(0.8f, true);
        std::vector<MapPoint*> got;
        const int n = m.SearchByBoW(&kf1, &kf2, got);
        Flat a(k1, d1, &fv1), b(k2, d2, &fv2);
        for (int i = 0; i < N; ++i) { a.flags[i] = kf1.mvpMapPoints[i] && !kf1.mvpMapPoints[i]->mbBad; b.flags[i] = kf2.mvpMapPoints[i] && !kf2.mvpMapPoints[i]->mbBad; }
        std::vector<int32_t> m12(N, -1); int32_t nm = 0; const int32_t z = 0; const int64_t z64 = 0;
        orbm_search_by_bow(&a.s, &b.s, 1, &z, &z, 0.8f, 1, 50, 0, 1, &z64, m12.data(), nullptr, &nm);
        EXPECT(n == nm && nm > 20, "SearchByBoW(KF,KF): %d matches vs oracle %d", n, nm);
        for (int i = 0; i < N; ++i) EXPECT(got[i] == (m12[i] >= 0 ? kf2.mvpMapPoints[m12[i]] : nullptr), "SearchByBoW(KF,KF) entry %d", i);
    }
    // --- SearchByBoW(KF, Frame) ---
    Frame F;
    F.N = N; F.mvKeys = k2; F.mvKeysUn = k2; F.mvuRight = ur2; F.mFeatVec = fv2; F.mDescriptors = d2; F.mvScaleFactors = g_sf;
    F.mvpMapPoints.assign(N, nullptr);
    {
        ORBmatcher m(0.7f, true);
        std::vector<MapPoint*> got;
        const int n = m.SearchByBoW(&kf1, F, got);
        Flat a(k1, d1, &fv1), b(k2, d2, &fv2);
        for (int i = 0; i < N; ++i) a.flags[i] = kf1.mvpMapPoints[i] && !kf1.mvpMapPoints[i]->mbBad;
        std::vector<int32_t> m12(N, -1); int32_t nm = 0; const int32_t z = 0; const int64_t z64 = 0;
        orbm_search_by_bow(&a.s, &b.s, 1, &z, &z, 0.7f, 1, 50, 1, 0, &z64, m12.data(), nullptr, &nm);
        std::vector<MapPoint*> exp(N, nullptr);
        for (int i = 0; i < N; ++i) if (m12[i] >= 0) exp[m12[i]] = kf1.mvpMapPoints[i];
        EXPECT(n == nm && nm > 20, "SearchByBoW(KF,F): %d matches vs oracle %d", n, nm);
        EXPECT(got == exp, "SearchByBoW(KF,F) vpMapPointMatches differ");
    }
    // --- SearchForTriangulation ---
    {
        kf1.Ow = cv::Mat(3, 1, CV_32FC1); kf2.Rcw = cv::Mat(3, 3, CV_32FC1); kf2.tcw = cv::Mat(3, 1, CV_32FC1);
        const float ow[3] = {0.4f, 0.03f, 0.2f}, tc[3] = {-0.5f, 0.01f, 0.9f};
        const float R[9] = {0.9998f, 0.f, 0.02f, 0.f, 1.f, 0.f, -0.02f, 0.f, 0.9998f};
        for (int i = 0; i < 3; ++i) { kf1.Ow.at<float>(i, 0) = ow[i]; kf2.tcw.at<float>(i, 0) = tc[i]; for (int j = 0; j < 3; ++j) kf2.Rcw.at<float>(i, j) = R[3 * i + j]; }
        cv::Mat F12(3, 3, CV_32FC1);
        const float f[9] = {0.f, -1e-4f, 0.02f, 1e-4f, 0.f, -0.6f, -0.02f, 0.6f, 0.f};   // mostly horizontal epipolar lines
        for (int i = 0; i < 9; ++i) F12.at<float>(i / 3, i % 3) = f[i];
        for (int only = 0; only < 2; ++only) {
            ORBmatcher m(0.6f, only == 0);
            std::vector<std::pair<size_t, size_t> > pairs;
            const int n = m.SearchForTriangulation(&kf1, &kf2, F12, pairs, only != 0);
            Flat a(k1, d1, &fv1), b(k2, d2, &fv2);
            for (int i = 0; i < N; ++i) { a.flags[i] = kf1.mvpMapPoints[i] != nullptr; b.flags[i] = kf2.mvpMapPoints[i] != nullptr; }
            a.s.u_right = ur1.data(); b.s.u_right = ur2.data();
            float C2[3];
            cvl_gemm3_f32(R, ow, tc, C2);
            const float invz = 1.0f / C2[2];
            const float ep[2] = {517.3f * C2[0] * invz + 318.6f, 516.5f * C2[1] * invz + 255.3f};
            std::vector<int32_t> m12(N, -1); int32_t nm = 0; const int32_t z = 0; const int64_t z64 = 0;
            orbm_search_for_triangulation(&a.s, &b.s, 1, &z, &z, f, ep, g_sf.data(), g_s2.data(), 8, only, only == 0, &z64, m12.data(), nullptr, &nm);
            std::vector<std::pair<size_t, size_t> > exp;
            for (int i = 0; i < N; ++i) if (m12[i] >= 0) exp.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
            EXPECT(n == nm && pairs == exp, "SearchForTriangulation(onlyStereo=%d): %d pairs vs oracle %d", only, n, nm);
            if (!only) EXPECT(nm > 5, "SearchForTriangulation: degenerate fixture (%d)", nm);
        }
    }
    // --- SearchByProjection(Frame, local map) ---
    {
        const int M = 3000;
        std::vector<MapPoint> mps(M);
        std::vector<MapPoint*> vp(M);
        for (int q = 0; q < M; ++q) {
            MapPoint& p = mps[q];
            vp[q] = &p;
            p.mDescriptor = cv::Mat(1, 32, CV_8UC1);
            if (rnd() % 10 < 6) {
                const int i = rnd() % N;
                std::memcpy(p.mDescriptor.ptr(0), d2.ptr(i), 32);
                flip(p.mDescriptor.ptr(0), 10);
                p.mTrackProjX = k2[i].pt.x + frand(-2, 2); p.mTrackProjY = k2[i].pt.y + frand(-2, 2); p.mnTrackScaleLevel = k2[i].octave;
            } else {
                for (int b = 0; b < 32; ++b) p.mDescriptor.at<uchar>(0, b) = (uchar)(rnd() & 255);
                p.mTrackProjX = frand(0, W); p.mTrackProjY = frand(0, H); p.mnTrackScaleLevel = rnd() % 8;
            }
            p.mTrackProjXR = p.mTrackProjX - 5; p.mTrackViewCos = rnd() % 2 ? 0.9999f : 0.99f;
            p.mbTrackInView = rnd() % 20 != 0; p.mbBad = rnd() % 25 == 0; p.nObs = rnd() % 5 == 0 ? 0 : 2;
        }
        for (int i = 0; i < N; ++i) F.mvpMapPoints[i] = (rnd() % 10 == 0) ? &pool[i] : nullptr;   // some keypoints are taken already
        for (int i = 0; i < N; ++i) if (rnd() % 7 == 0) pool[i].nObs = 0;
        // oracle first (the shell mutates F.mvpMapPoints)
        Flat a(k2, d2, nullptr);
        for (int i = 0; i < N; ++i) a.flags[i] = F.mvpMapPoints[i] ? (F.mvpMapPoints[i]->nObs > 0 ? 1 : 2) : 0;
        a.s.u_right = ur2.data();
        const float grid[4] = {0, 0, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
        a.s.grid = grid;
        int32_t mp_off[2] = {0, M};
        std::vector<float> px(M), py(M), pxr(M), vc(M);
        std::vector<int32_t> lv(M);
        std::vector<uint8_t> fl(M), dd((size_t)M * 32);
        for (int q = 0; q < M; ++q) {
            px[q] = mps[q].mTrackProjX; py[q] = mps[q].mTrackProjY; pxr[q] = mps[q].mTrackProjXR; vc[q] = mps[q].mTrackViewCos;
            lv[q] = mps[q].mnTrackScaleLevel;
            fl[q] = (uint8_t)((mps[q].mbTrackInView ? 1 : 0) | (mps[q].mbBad ? 2 : 0) | (mps[q].nObs > 0 ? 4 : 0));
            std::memcpy(&dd[(size_t)q * 32], mps[q].mDescriptor.ptr(0), 32);
        }
        orbgpu_mappoint_set ms = {mp_off, px.data(), py.data(), pxr.data(), vc.data(), lv.data(), fl.data(), dd.data()};
        std::vector<int32_t> kpm(N, -1); int32_t nm = 0;
        orbm_search_by_projection(&a.s, &ms, g_sf.data(), 8, 3.0f, 0.8f, kpm.data(), nullptr, nullptr, nullptr, &nm);
        std::vector<MapPoint*> exp = F.mvpMapPoints;
        for (int i = 0; i < N; ++i) if (kpm[i] >= 0) exp[i] = vp[kpm[i]];
        ORBmatcher m(0.8f, true);
        const int n = m.SearchByProjection(F, vp, 3.0f);
        EXPECT(n == nm && nm > 50, "SearchByProjection: %d matches vs oracle %d", n, nm);
        EXPECT(F.mvpMapPoints == exp, "SearchByProjection: F.mvpMapPoints differ");
    }
}

// --- SearchByProjection(CurrentFrame, LastFrame, th, bMono): TrackWithMotionModel's matcher ---
static void test_track_last_frame() {
    const int W = 640, H = 480, N = 1000;
    Frame::fx = 517.3f; Frame::fy = 516.5f; Frame::cx = 318.6f; Frame::cy = 255.3f;
    Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = W; Frame::mnMaxY = H;
    Frame::mfGridElementWidthInv = 64.0f / W; Frame::mfGridElementHeightInv = 48.0f / H;
    for (int variant = 0; variant < 3; ++variant) {   // 0: mono, 1: stereo forward motion, 2: stereo backward motion
        Frame last, cur;
        last.N = N; cur.N = N;
        last.mvKeys = random_keys(N, W, H); last.mvKeysUn = last.mvKeys;
        last.mDescriptors = random_desc(N);
        last.mvScaleFactors = g_sf; cur.mvScaleFactors = g_sf;
        last.mTcw = cv::Mat(4, 4, CV_32FC1); cur.mTcw = cv::Mat(4, 4, CV_32FC1);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { last.mTcw.at<float>(i, j) = i == j; cur.mTcw.at<float>(i, j) = i == j; }
        const float ang = 0.01f;
        cur.mTcw.at<float>(0, 0) = std::cos(ang); cur.mTcw.at<float>(0, 2) = std::sin(ang); cur.mTcw.at<float>(2, 0) = -std::sin(ang); cur.mTcw.at<float>(2, 2) = std::cos(ang);
        cur.mTcw.at<float>(0, 3) = 0.02f; cur.mTcw.at<float>(2, 3) = variant == 1 ? -0.3f : (variant == 2 ? 0.3f : 0.01f);
        cur.mb = 0.08f; cur.mbf = 40.f; last.mb = 0.08f; last.mbf = 40.f;
        // map points: back-projection of the last frame's key points at random depths
        std::vector<MapPoint> pool(N);
        last.mvpMapPoints.assign(N, nullptr); last.mvbOutlier.assign(N, false);
        for (int i = 0; i < N; ++i) {
            if (rnd() % 10 < 8) {
                MapPoint& p = pool[i];
                const float z = frand(2.f, 12.f);
                p.mWorldPos = cv::Mat(3, 1, CV_32FC1);
                p.mWorldPos.at<float>(0, 0) = (last.mvKeys[i].pt.x - Frame::cx) / Frame::fx * z;
                p.mWorldPos.at<float>(1, 0) = (last.mvKeys[i].pt.y - Frame::cy) / Frame::fy * z;
                p.mWorldPos.at<float>(2, 0) = rnd() % 50 == 0 ? -z : z;
                p.mDescriptor = cv::Mat(1, 32, CV_8UC1);
                std::memcpy(p.mDescriptor.ptr(0), last.mDescriptors.ptr(i), 32);
                p.nObs = rnd() % 4 == 0 ? 0 : 2;
                last.mvpMapPoints[i] = &p;
                last.mvbOutlier[i] = rnd() % 15 == 0;
            }
        }
        // current frame: the same scene seen from the new pose (+ noise), plus unrelated key points
        cur.mvKeys = random_keys(N, W, H);
        cur.mDescriptors = random_desc(N);
        for (int i = 0; i < N; ++i) {
            MapPoint* p = last.mvpMapPoints[i];
            if (!p || rnd() % 10 >= 7) continue;
            float xc[3];
            for (int r = 0; r < 3; ++r) { double a = 0; for (int c = 0; c < 3; ++c) a += (double)cur.mTcw.at<float>(r, c) * p->mWorldPos.at<float>(c, 0); xc[r] = (float)(a + cur.mTcw.at<float>(r, 3)); }
            if (xc[2] <= 0.1f) continue;
            const int j = rnd() % N;
            cur.mvKeys[j] = last.mvKeys[i];
            cur.mvKeys[j].pt.x = Frame::fx * xc[0] / xc[2] + Frame::cx + frand(-2, 2);
            cur.mvKeys[j].pt.y = Frame::fy * xc[1] / xc[2] + Frame::cy + frand(-2, 2);
            cur.mvKeys[j].angle = std::fmod(last.mvKeys[i].angle + frand(-12, 12) + 360.f, 360.f);
            std::memcpy(cur.mDescriptors.ptr(j), last.mDescriptors.ptr(i), 32);
            flip(cur.mDescriptors.ptr(j), 10);
        }
        cur.mvKeysUn = cur.mvKeys;
        cur.mvuRight.assign(N, -1.f);
        if (variant) for (int i = 0; i < N; ++i) if (rnd() % 2) cur.mvuRight[i] = cur.mvKeys[i].pt.x - frand(2, 18);
        std::vector<MapPoint> held(N);
        cur.mvpMapPoints.assign(N, nullptr);
        for (int i = 0; i < N; ++i) if (rnd() % 12 == 0) { cur.mvpMapPoints[i] = &held[i]; held[i].nObs = rnd() % 2; }
        const bool bMono = variant == 0;
        const float th = bMono ? 15.f : 7.f;

        // ---- oracle: the same projection arithmetic, then the windowed search port
        Flat a(cur.mvKeysUn, cur.mDescriptors, nullptr);
        for (int i = 0; i < N; ++i) a.flags[i] = cur.mvpMapPoints[i] ? (cur.mvpMapPoints[i]->nObs > 0 ? 1 : 2) : 0;
        a.s.u_right = cur.mvuRight.data();
        const float grid[4] = {0, 0, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
        a.s.grid = grid;
        float twc[3], tlc[3];
        float Rc[9], tcv[3], Rl[9], tl[3];
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) { Rc[3 * i + j] = cur.mTcw.at<float>(i, j); Rl[3 * i + j] = last.mTcw.at<float>(i, j); } tcv[i] = cur.mTcw.at<float>(i, 3); tl[i] = last.mTcw.at<float>(i, 3); }
        cvl_gemm3t_neg_f32(Rc, tcv, twc);
        cvl_gemm3_f32(Rl, twc, tl, tlc);
        const bool fwd = tlc[2] > cur.mb && !bMono, bwd = -tlc[2] > cur.mb && !bMono;
        EXPECT(variant == 0 || (variant == 1) == fwd, "variant %d: forward=%d backward=%d", variant, fwd, bwd);
        int32_t q_off[2] = {0, N};
        std::vector<float> qu(N), qv(N), qr(N), qur(N), qa(N);
        std::vector<int32_t> lo(N), hi(N);
        std::vector<uint8_t> qf(N, 0), qd((size_t)N * 32, 0);
        for (int i = 0; i < N; ++i) {
            MapPoint* p = last.mvpMapPoints[i];
            if (!p || last.mvbOutlier[i]) continue;
            float xc[3];
            cvl_gemm3_f32(Rc, p->mWorldPos.ptr<float>(0), tcv, xc);
            const float invzc = 1.0 / xc[2];
            if (invzc < 0) continue;
            const float u = Frame::fx * xc[0] * invzc + Frame::cx, v = Frame::fy * xc[1] * invzc + Frame::cy;
            if (u < 0 || u > W || v < 0 || v > H) continue;
            const int o = last.mvKeys[i].octave;
            qu[i] = u; qv[i] = v; qr[i] = th * g_sf[o]; qur[i] = u - cur.mbf * invzc; qa[i] = last.mvKeysUn[i].angle;
            if (fwd) { lo[i] = o; hi[i] = -1; } else if (bwd) { lo[i] = 0; hi[i] = o; } else { lo[i] = o - 1; hi[i] = o + 1; }
            qf[i] = (uint8_t)(1 | (p->nObs > 0 ? 4 : 0));
            std::memcpy(&qd[(size_t)i * 32], p->mDescriptor.ptr(0), 32);
        }
        orbgpu_window_query_set qs = {q_off, qu.data(), qv.data(), qr.data(), lo.data(), hi.data(), qur.data(), qf.data(), qd.data(), qa.data()};
        std::vector<int32_t> kpm(N, -1); int32_t nm = 0;
        orbm_search_windowed(&a.s, &qs, 100, 0, 1, kpm.data(), nullptr, nullptr, &nm);
        std::vector<MapPoint*> exp = cur.mvpMapPoints;
        for (int i = 0; i < N; ++i) { if (kpm[i] >= 0) exp[i] = last.mvpMapPoints[kpm[i]]; else if (kpm[i] == -2) exp[i] = nullptr; }

        ORBmatcher m(0.9f, true);
        const int n = m.SearchByProjection(cur, last, th, bMono);
        EXPECT(n == nm && nm > 100, "SearchByProjection(Frame,Frame) variant %d: %d matches vs oracle %d", variant, n, nm);
        EXPECT(cur.mvpMapPoints == exp, "SearchByProjection(Frame,Frame) variant %d: mvpMapPoints differ", variant);
    }
}

// --- SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): Relocalization's matcher (Tracking.cc:1435-1475) ---
static void test_relocalization_search() {
    const int W = 640, H = 480, N = 1000;
    Frame::fx = 517.3f; Frame::fy = 516.5f; Frame::cx = 318.6f; Frame::cy = 255.3f;
    Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = W; Frame::mnMaxY = H;
    Frame::mfGridElementWidthInv = 64.0f / W; Frame::mfGridElementHeightInv = 48.0f / H;
    std::vector<cv::KeyPoint> kkeys = random_keys(N, W, H);
    cv::Mat kdesc = random_desc(N);
    KeyFrame kf(kkeys, std::vector<float>(N, -1.f), kdesc, g_sf, g_s2, Frame::fx, Frame::fy, Frame::cx, Frame::cy);
    Frame cur;
    cur.N = N;
    cur.mvScaleFactors = g_sf;
    cur.mfLogScaleFactor = std::log(1.2f);
    cur.mnScaleLevels = 8;
    cur.mTcw = cv::Mat(4, 4, CV_32FC1);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) cur.mTcw.at<float>(i, j) = i == j;
    const float ang = -0.015f;
    cur.mTcw.at<float>(0, 0) = std::cos(ang); cur.mTcw.at<float>(0, 2) = std::sin(ang); cur.mTcw.at<float>(2, 0) = -std::sin(ang); cur.mTcw.at<float>(2, 2) = std::cos(ang);
    cur.mTcw.at<float>(0, 3) = -0.04f; cur.mTcw.at<float>(1, 3) = 0.01f; cur.mTcw.at<float>(2, 3) = 0.05f;
    float Rc[9], tcv[3], Ow[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Rc[3 * i + j] = cur.mTcw.at<float>(i, j); tcv[i] = cur.mTcw.at<float>(i, 3); }
    cvl_gemm3t_neg_f32(Rc, tcv, Ow);
    // the key frame's map points: back-projection of its key points (key frame at the origin), distance ranges around the octave
    std::vector<MapPoint> pool(N);
    std::set<MapPoint*> found;
    for (int i = 0; i < N; ++i) {
        if (rnd() % 10 >= 8) continue;
        MapPoint& p = pool[i];
        const float z = frand(2.f, 12.f);
        p.mWorldPos = cv::Mat(3, 1, CV_32FC1);
        p.mWorldPos.at<float>(0, 0) = (kkeys[i].pt.x - Frame::cx) / Frame::fx * z;
        p.mWorldPos.at<float>(1, 0) = (kkeys[i].pt.y - Frame::cy) / Frame::fy * z;
        p.mWorldPos.at<float>(2, 0) = z;
        const float d = std::sqrt(p.mWorldPos.at<float>(0, 0) * p.mWorldPos.at<float>(0, 0) + p.mWorldPos.at<float>(1, 0) * p.mWorldPos.at<float>(1, 0) + z * z);
        p.mfMaxDistance = d * std::pow(1.2f, (float)kkeys[i].octave - 0.5f) * (rnd() % 25 == 0 ? 40.f : 1.f);   // a few out of range
        p.mfMinDistance = p.mfMaxDistance / std::pow(1.2f, 7.f);
        p.mDescriptor = cv::Mat(1, 32, CV_8UC1);
        std::memcpy(p.mDescriptor.ptr(0), kdesc.ptr(i), 32);
        p.mbBad = rnd() % 30 == 0;
        kf.mvpMapPoints[i] = &p;
        if (rnd() % 12 == 0) found.insert(&p);
    }
    // current frame: the same points seen from the new pose (+ noise), plus unrelated key points; some key points already hold a MapPoint
    cur.mvKeys = random_keys(N, W, H);
    cur.mDescriptors = random_desc(N);
    for (int i = 0; i < N; ++i) {
        MapPoint* p = kf.mvpMapPoints[i];
        if (!p || rnd() % 10 >= 8) continue;
        float xc[3];
        cvl_gemm3_f32(Rc, p->mWorldPos.ptr<float>(0), tcv, xc);
        if (xc[2] <= 0.1f) continue;
        const int j = rnd() % N;
        cur.mvKeys[j] = kkeys[i];
        cur.mvKeys[j].pt.x = Frame::fx * xc[0] / xc[2] + Frame::cx + frand(-2, 2);
        cur.mvKeys[j].pt.y = Frame::fy * xc[1] / xc[2] + Frame::cy + frand(-2, 2);
        cur.mvKeys[j].angle = std::fmod(kkeys[i].angle + frand(-12, 12) + 360.f, 360.f);
        std::memcpy(cur.mDescriptors.ptr(j), kdesc.ptr(i), 32);
        flip(cur.mDescriptors.ptr(j), 10);
    }
    cur.mvKeysUn = cur.mvKeys;
    std::vector<MapPoint> held(N);
    cur.mvpMapPoints.assign(N, nullptr);
    for (int i = 0; i < N; ++i) if (rnd() % 12 == 0) cur.mvpMapPoints[i] = &held[i];
    const float th = 10.f;
    const int ORBdist = 100;

    // ---- oracle: the same projection arithmetic, then the windowed search port (any MapPoint blocks a key point)
    Flat a(cur.mvKeysUn, cur.mDescriptors, nullptr);
    for (int i = 0; i < N; ++i) a.flags[i] = cur.mvpMapPoints[i] ? 2 : 0;
    const float grid[4] = {0, 0, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    a.s.grid = grid;
    int32_t q_off[2] = {0, N};
    std::vector<float> qu(N), qv(N), qr(N), qa(N);
    std::vector<int32_t> lo(N), hi(N);
    std::vector<uint8_t> qf(N, 0), qd((size_t)N * 32, 0);
    int live = 0;
    for (int i = 0; i < N; ++i) {
        MapPoint* p = kf.mvpMapPoints[i];
        if (!p || p->mbBad || found.count(p)) continue;
        float xc[3];
        cvl_gemm3_f32(Rc, p->mWorldPos.ptr<float>(0), tcv, xc);
        const float invzc = 1.0 / xc[2];
        const float u = Frame::fx * xc[0] * invzc + Frame::cx, v = Frame::fy * xc[1] * invzc + Frame::cy;
        if (u < 0 || u > W || v < 0 || v > H) continue;
        const float po[3] = {p->mWorldPos.at<float>(0, 0) - Ow[0], p->mWorldPos.at<float>(1, 0) - Ow[1], p->mWorldPos.at<float>(2, 0) - Ow[2]};
        const float dist3D = (float)cvl_norm3_f32(po);
        if (dist3D < 0.8f * p->mfMinDistance || dist3D > 1.2f * p->mfMaxDistance) continue;
        const int lvl = p->PredictScale(dist3D, &cur);
        qu[i] = u; qv[i] = v; qr[i] = th * g_sf[lvl]; qa[i] = kkeys[i].angle;
        lo[i] = lvl - 1; hi[i] = lvl + 1;
        qf[i] = 1;
        ++live;
        std::memcpy(&qd[(size_t)i * 32], p->mDescriptor.ptr(0), 32);
    }
    orbgpu_window_query_set qs = {q_off, qu.data(), qv.data(), qr.data(), lo.data(), hi.data(), nullptr, qf.data(), qd.data(), qa.data()};
    std::vector<int32_t> kpm(N, -1); int32_t nm = 0;
    orbm_search_windowed(&a.s, &qs, ORBdist, 1, 1, kpm.data(), nullptr, nullptr, &nm);
    std::vector<MapPoint*> exp = cur.mvpMapPoints;
    for (int i = 0; i < N; ++i) { if (kpm[i] >= 0) exp[i] = kf.mvpMapPoints[kpm[i]]; else if (kpm[i] == -2) exp[i] = nullptr; }

    ORBmatcher m(0.9f, true);
    const int n = m.SearchByProjection(cur, &kf, found, th, ORBdist);
    EXPECT(n == nm && nm > 100, "SearchByProjection(Frame,KeyFrame): %d matches vs oracle %d (%d live queries)", n, nm, live);
    EXPECT(cur.mvpMapPoints == exp, "SearchByProjection(Frame,KeyFrame): mvpMapPoints differ");
    printf("SearchByProjection(Frame,KeyFrame): %d matches of %d live queries, oracle equal: %s\n", n, live, cur.mvpMapPoints == exp ? "yes" : "no");
}

// --- Fuse(pKF, vpMapPoints, th): LocalMapping::SearchInNeighbors' matcher (LocalMapping.cc:589-640) -------------------------
struct FuseWorld {
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    std::vector<float> ur;
    KeyFrame *kf, *other;
    std::vector<MapPoint> held, cand;
    std::vector<MapPoint*> vp;
    FuseWorld() : kf(nullptr), other(nullptr) {}
    ~FuseWorld() { delete kf; delete other; }
    int id(MapPoint* p) const {
        if (!p) return -1;
        if (!held.empty() && p >= &held[0] && p <= &held[held.size() - 1]) return 1000 + (int)(p - &held[0]);
        return 5000 + (int)(p - &cand[0]);
    }
};
static void make_fuse_world(FuseWorld& w, uint32_t seed) {
    g_seed = seed;
    const int W = 640, H = 480, N = 800, M = 700;
    w.keys = random_keys(N, W, H);
    w.desc = random_desc(N);
    w.ur.assign(N, -1.f);
    for (int i = 0; i < N; ++i) if (rnd() % 2) w.ur[i] = w.keys[i].pt.x - frand(2, 18);
    const float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f;
    w.kf = new KeyFrame(w.keys, w.ur, w.desc, g_sf, g_s2, fx, fy, cx, cy);
    w.other = new KeyFrame(w.keys, w.ur, w.desc, g_sf, g_s2, fx, fy, cx, cy);
    KeyFrame& kf = *w.kf;
    kf.mbf = 40.f; kf.mfLogScaleFactor = std::log(1.2f); kf.mnScaleLevels = 8;
    kf.mvInvLevelSigma2.resize(8);
    for (int l = 0; l < 8; ++l) kf.mvInvLevelSigma2[l] = 1.0f / g_s2[l];
    kf.mnMinX = 0; kf.mnMinY = 0; kf.mnMaxX = W; kf.mnMaxY = H;
    kf.mfGridElementWidthInv = 64.0f / W; kf.mfGridElementHeightInv = 48.0f / H;
    const float ang = 0.02f;
    float R[9] = {std::cos(ang), 0.f, std::sin(ang), 0.f, 1.f, 0.f, -std::sin(ang), 0.f, std::cos(ang)}, t[3] = {0.05f, -0.02f, 0.1f}, Ow[3];
    cvl_gemm3t_neg_f32(R, t, Ow);
    kf.Rcw = cv::Mat(3, 3, CV_32FC1); kf.tcw = cv::Mat(3, 1, CV_32FC1); kf.Ow = cv::Mat(3, 1, CV_32FC1);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) kf.Rcw.at<float>(i, j) = R[3 * i + j]; kf.tcw.at<float>(i, 0) = t[i]; kf.Ow.at<float>(i, 0) = Ow[i]; }
    // some key points of the key frame already hold a MapPoint (seen from `other` too)
    w.held.resize(N);
    for (int i = 0; i < N; ++i) {
        if (rnd() % 3) continue;
        MapPoint& p = w.held[i];
        p.AddObservation(w.kf, i);
        if (rnd() % 2) { p.AddObservation(w.other, i); w.other->mvpMapPoints[i] = &p; }
        if (rnd() % 3 == 0) p.nObs += 2;
        p.mbBad = rnd() % 25 == 0;
        kf.mvpMapPoints[i] = &p;
    }
    // candidate points: world positions that project onto key points of the key frame (+ noise), a few impossible ones
    w.cand.resize(M);
    w.vp.assign(M, nullptr);
    for (int k = 0; k < M; ++k) {
        MapPoint& p = w.cand[k];
        const int j = rnd() % N;
        const float z = (rnd() % 30 == 0 ? -1.f : 1.f) * frand(2.f, 12.f);
        const float u = w.keys[j].pt.x + frand(-1.5f, 1.5f), v = w.keys[j].pt.y + frand(-1.5f, 1.5f);
        const float pc[3] = {(u - cx) / fx * z - t[0], (v - cy) / fy * z - t[1], z - t[2]};
        p.mWorldPos = cv::Mat(3, 1, CV_32FC1);
        for (int r = 0; r < 3; ++r) p.mWorldPos.at<float>(r, 0) = R[r] * pc[0] + R[3 + r] * pc[1] + R[6 + r] * pc[2];   // R^T (pc - t)
        float po[3], d = 0;
        for (int r = 0; r < 3; ++r) { po[r] = p.mWorldPos.at<float>(r, 0) - Ow[r]; d += po[r] * po[r]; }
        d = std::sqrt(d);
        p.mNormalVector = cv::Mat(3, 1, CV_32FC1);
        for (int r = 0; r < 3; ++r) p.mNormalVector.at<float>(r, 0) = po[r] / d * (rnd() % 20 == 0 ? -1.f : 1.f);
        p.mfMaxDistance = d * std::pow(1.2f, (float)w.keys[j].octave - 0.5f) * (rnd() % 25 == 0 ? 40.f : 1.f);
        p.mfMinDistance = p.mfMaxDistance / std::pow(1.2f, 7.f);
        p.mDescriptor = cv::Mat(1, 32, CV_8UC1);
        std::memcpy(p.mDescriptor.ptr(0), w.desc.ptr(j), 32);
        flip(p.mDescriptor.ptr(0), rnd() % 6 == 0 ? 90 : 12);
        p.nObs = rnd() % 5;
        p.mbBad = rnd() % 30 == 0;
        if (rnd() % 20 == 0) p.AddObservation(w.kf, (size_t)(rnd() % N));   // already seen in this key frame
        if (rnd() % 15) w.vp[k] = &p;
        if (k > 3 && rnd() % 12 == 0) w.vp[k] = w.vp[k - 3];                  // the same point twice in the list
    }
}
static void test_fuse() {
    FuseWorld A, B;
    make_fuse_world(A, 4242);
    make_fuse_world(B, 4242);
    const float th = 3.0f;
    ORBmatcher m;
    const int nA = m.Fuse(A.kf, A.vp, th);
    // world B: the reference's loop (ORBmatcher.cc:977-1137) restated around the oracle's best-only search
    KeyFrame& kf = *B.kf;
    const int N = kf.N, M = (int)B.vp.size();
    float R[9], t[3], Ow[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[3 * i + j] = kf.Rcw.at<float>(i, j); t[i] = kf.tcw.at<float>(i, 0); Ow[i] = kf.Ow.at<float>(i, 0); }
    Flat a(kf.mvKeysUn, kf.mDescriptors, nullptr);
    a.s.kp_flags = nullptr;
    a.s.u_right = kf.mvuRight.data();
    const float grid[4] = {0, 0, kf.mfGridElementWidthInv, kf.mfGridElementHeightInv};
    a.s.grid = grid;
    int32_t q_off[2] = {0, M};
    std::vector<float> qu(M), qv(M), qr(M), qur(M);
    std::vector<int32_t> lo(M), hi(M);
    std::vector<uint8_t> qf(M, 0), qd((size_t)M * 32, 0);
    for (int i = 0; i < M; ++i) {
        MapPoint* p = B.vp[i];
        if (!p) continue;
        float pc[3];
        cvl_gemm3_f32(R, p->mWorldPos.ptr<float>(0), t, pc);
        if (pc[2] < 0.0f) continue;
        const float invz = 1 / pc[2];
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = kf.fx * x + kf.cx, v = kf.fy * y + kf.cy;
        if (!kf.IsInImage(u, v)) continue;
        const float po[3] = {p->mWorldPos.at<float>(0, 0) - Ow[0], p->mWorldPos.at<float>(1, 0) - Ow[1], p->mWorldPos.at<float>(2, 0) - Ow[2]};
        const float dist3D = (float)cvl_norm3_f32(po);
        if (dist3D < p->GetMinDistanceInvariance() || dist3D > p->GetMaxDistanceInvariance()) continue;
        const float nrm[3] = {p->mNormalVector.at<float>(0, 0), p->mNormalVector.at<float>(1, 0), p->mNormalVector.at<float>(2, 0)};
        if (cvl_dot3_f32(po, nrm) < 0.5 * dist3D) continue;
        const int lvl = p->PredictScale(dist3D, &kf);
        qu[i] = u; qv[i] = v; qur[i] = u - kf.mbf * invz; qr[i] = th * g_sf[lvl];
        lo[i] = lvl - 1; hi[i] = lvl;
        qf[i] = 1;
        std::memcpy(&qd[(size_t)i * 32], p->mDescriptor.ptr(0), 32);
    }
    orbgpu_window_query_set qs = {q_off, qu.data(), qv.data(), qr.data(), lo.data(), hi.data(), qur.data(), qf.data(), qd.data(), nullptr};
    std::vector<int32_t> best(M, -1), bd(M, 256);
    orbm_search_window_best(&a.s, &qs, kf.mvInvLevelSigma2.data(), 0, best.data(), bd.data());
    int nB = 0, replaced_a = 0, replaced_b = 0, added = 0;
    for (int i = 0; i < M; ++i) {
        MapPoint* pMP = B.vp[i];
        if (!pMP || !qf[i]) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(&kf)) continue;
        if (bd[i] > 50) continue;
        MapPoint* pMPinKF = kf.GetMapPoint(best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) { pMP->Replace(pMPinKF); ++replaced_a; }
                else { pMPinKF->Replace(pMP); ++replaced_b; }
            }
        } else {
            pMP->AddObservation(&kf, best[i]);
            kf.AddMapPoint(pMP, best[i]);
            ++added;
        }
        nB++;
    }
    int bad = 0;
    for (int i = 0; i < N; ++i) {
        bad += A.id(A.kf->mvpMapPoints[i]) != B.id(B.kf->mvpMapPoints[i]);
        bad += A.id(A.other->mvpMapPoints[i]) != B.id(B.other->mvpMapPoints[i]);
        bad += A.held[i].mbBad != B.held[i].mbBad || A.held[i].nObs != B.held[i].nObs || A.id(A.held[i].mpReplaced) != B.id(B.held[i].mpReplaced);
    }
    for (int k = 0; k < M; ++k)
        bad += A.cand[k].mbBad != B.cand[k].mbBad || A.cand[k].nObs != B.cand[k].nObs || A.id(A.cand[k].mpReplaced) != B.id(B.cand[k].mpReplaced);
    EXPECT(nA == nB && bad == 0 && added > 30 && replaced_a > 5 && replaced_b > 5, "Fuse: %d fused (oracle loop %d), %d state differences; added %d, replaced %d / %d", nA, nB,
           bad, added, replaced_a, replaced_b);
    printf("Fuse: %d fused (%d added, %d + %d replaced), map state equal to the restated loop: %s\n", nA, added, replaced_a, replaced_b, bad ? "no" : "yes");
}

// ---- ORBmatcher::SearchForInitialization as Tracking::MonocularInitialization calls it (Tracking.cc:609-612) ----------------
static void test_search_for_initialization() {
    const int N = 900;
    Frame F1, F2;
    F1.mvKeysUn = random_keys(N, 640, 480);
    F1.mDescriptors = random_desc(N);
    // F2 sees most of F1's points again a few pixels away, with a few descriptor bits flipped; several F1 points share a descriptor
    F2.mvKeysUn = random_keys(N, 640, 480);
    F2.mDescriptors = random_desc(N);
    for (int i = 0; i < N; ++i) {
        if (i % 7 == 3) std::memcpy(F1.mDescriptors.ptr(i), F1.mDescriptors.ptr(i - 1), 32);   // rivals for one F2 point
        if (i % 5 != 0) {
            F2.mvKeysUn[i] = F1.mvKeysUn[i];
            F2.mvKeysUn[i].pt.x += frand(-6, 6);
            F2.mvKeysUn[i].pt.y += frand(-6, 6);
            F2.mvKeysUn[i].angle = F1.mvKeysUn[i].angle + frand(-10, 10);
            if (F2.mvKeysUn[i].angle < 0) F2.mvKeysUn[i].angle += 360.f;
            if (F2.mvKeysUn[i].angle >= 360.f) F2.mvKeysUn[i].angle -= 360.f;
            std::memcpy(F2.mDescriptors.ptr(i), F1.mDescriptors.ptr(i), 32);
            for (int b = 0; b < 12; ++b) F2.mDescriptors.at<uchar>(i, rnd() % 32) ^= (uchar)(1u << (rnd() % 8));
        }
    }
    Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = 640; Frame::mnMaxY = 480;
    Frame::mfGridElementWidthInv = 64.f / 640.f; Frame::mfGridElementHeightInv = 48.f / 480.f;
    std::vector<cv::Point2f> prev(N), prev0;
    for (int i = 0; i < N; ++i) prev[i] = F1.mvKeysUn[i].pt;
    prev0 = prev;
    std::vector<int> m12;
    ORBmatcher matcher(0.9f, true);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
    // oracle on the same flattened inputs
    std::vector<uint8_t> fl(N);
    std::vector<float> qu(N), qv(N), qr(N, 100.f), qa(N);
    std::vector<int32_t> lo(N, 0), hi(N, 0);
    for (int i = 0; i < N; ++i) { qu[i] = prev0[i].x; qv[i] = prev0[i].y; qa[i] = F1.mvKeysUn[i].angle; fl[i] = F1.mvKeysUn[i].octave > 0 ? 0 : 1; }
    const int32_t off[2] = {0, N};
    orbgpu_frame_set fs;
    std::memset(&fs, 0, sizeof fs);
    fs.n_frames = 1; fs.kp_off = off; fs.keys_un = (const orbgpu_keypoint*)F2.mvKeysUn.data(); fs.desc = F2.mDescriptors.ptr(0);
    const float grid[4] = {0, 0, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    fs.grid = grid;
    orbgpu_window_query_set qs;
    std::memset(&qs, 0, sizeof qs);
    qs.q_off = off; qs.u = qu.data(); qs.v = qv.data(); qs.radius = qr.data(); qs.min_level = lo.data(); qs.max_level = hi.data();
    qs.flags = fl.data(); qs.desc = F1.mDescriptors.ptr(0); qs.angle = qa.data();
    std::vector<int32_t> e12(N, -1);
    int32_t en = 0;
    orbm_search_for_initialization(&fs, &qs, 0.9f, 1, e12.data(), &en);
    int bad = 0, moved = 0;
    for (int i = 0; i < N; ++i) {
        bad += m12[i] != e12[i];
        if (m12[i] >= 0) { moved += prev[i].x == F2.mvKeysUn[m12[i]].pt.x && prev[i].y == F2.mvKeysUn[m12[i]].pt.y; }
        else bad += !(prev[i].x == prev0[i].x && prev[i].y == prev0[i].y);
    }
    EXPECT(n == en && bad == 0 && moved == n + 0 * moved && n > 50, "SearchForInitialization: %d matches (oracle %d), %d entries differ, %d vbPrevMatched updated", n, en, bad, moved);
    printf("SearchForInitialization: %d matches, oracle equal: %s\n", n, bad ? "no" : "yes");
}

// ---- ORBVocabulary::transform as Frame::ComputeBoW calls it (Frame.cc:425-432) -------------------------------------------
struct VocRecords {
    int k, L;
    std::vector<int32_t> parent;
    std::vector<uint8_t> is_leaf, desc;
    std::vector<double> weight;
};
static void grow(VocRecords& V, int pid, const uint8_t* pdesc, int level) {
    std::vector<int> ids;
    const int nch = level == 2 ? V.k - 1 : V.k;   // one level with fewer children
    for (int c = 0; c < nch; ++c) {
        uint8_t d[32];
        for (int b = 0; b < 32; ++b) d[b] = level == 1 ? (uint8_t)rnd() : (uint8_t)(pdesc[b] ^ ((rnd() % 5 == 0) ? (1u << (rnd() % 8)) : 0u));
        V.parent.push_back(pid);
        V.is_leaf.push_back(level == V.L);
        V.desc.insert(V.desc.end(), d, d + 32);
        V.weight.push_back(level == V.L ? (rnd() % 17 == 0 ? 0.0 : 0.25 + (rnd() % 100000) / 17000.0) : 0.0);
        ids.push_back((int)V.parent.size());
    }
    if (level < V.L)
        for (size_t i = 0; i < ids.size(); ++i) grow(V, ids[i], &V.desc[(size_t)(ids[i] - 1) * 32], level + 1);
}
static std::string write_voc(const VocRecords& V) {
    char name[] = "/tmp/orbgpu_vocXXXXXX";
    const int fd = mkstemp(name);
    FILE* f = fdopen(fd, "w");
    fprintf(f, "%d %d  0 0", V.k, V.L);
    for (size_t r = 0; r < V.parent.size(); ++r) {
        fprintf(f, "\n%d %d ", V.parent[r], (int)V.is_leaf[r]);
        for (int b = 0; b < 32; ++b) fprintf(f, "%d ", (int)V.desc[r * 32 + b]);
        fprintf(f, " %.17g", V.weight[r]);
    }
    fclose(f);
    return name;
}
static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& D) {   // Converter.cc:29-37
    std::vector<cv::Mat> v;
    v.reserve(D.rows);
    for (int j = 0; j < D.rows; ++j) v.push_back(D.row(j));
    return v;
}
static void test_vocabulary() {
    VocRecords V;
    V.k = 6; V.L = 4;
    const uint8_t zero[32] = {0};
    grow(V, 0, zero, 1);
    const std::string path = write_voc(V);
    ORBVocabulary voc;
    EXPECT(voc.loadFromTextFile(path), "loadFromTextFile failed");
    remove(path.c_str());
    void* o = orbo_voc_create(V.k, V.L, 0, 0, (int)V.parent.size(), V.parent.data(), V.is_leaf.data(), V.desc.data(), V.weight.data());
    for (int n : {0, 1, 1500}) {
        Frame F;
        F.mDescriptors = cv::Mat(n, 32, CV_8U);
        for (int i = 0; i < n; ++i) {
            const size_t leaf = rnd() % V.parent.size();
            for (int b = 0; b < 32; ++b) F.mDescriptors.at<uchar>(i, b) = (uint8_t)(V.desc[leaf * 32 + b] ^ ((rnd() % 9 == 0) ? (1u << (rnd() % 8)) : 0u));
        }
        voc.transform(toDescriptorVector(F.mDescriptors), F.mBowVec, F.mFeatVec, 2);   // Frame::ComputeBoW with levelsup 2 of 4 levels
        std::vector<uint32_t> bw(n + 1), fn(n + 1), ff(n + 1);
        std::vector<double> bv(n + 1);
        std::vector<int32_t> fo(n + 2);
        int nb = 0, nf = 0;
        orbo_voc_transform(o, n ? F.mDescriptors.ptr(0) : zero, n, 2, &nb, bw.data(), bv.data(), &nf, fn.data(), fo.data(), ff.data(), 0, 0);
        EXPECT((int)F.mBowVec.size() == nb && (int)F.mFeatVec.size() == nf, "vocabulary n=%d: %zu words / %zu nodes, oracle %d / %d", n,
               F.mBowVec.size(), F.mFeatVec.size(), nb, nf);
        int i = 0, bad = 0;
        for (DBoW2::BowVector::const_iterator it = F.mBowVec.begin(); it != F.mBowVec.end() && i < nb; ++it, ++i)
            bad += it->first != bw[i] || it->second != bv[i];
        int j = 0;
        for (DBoW2::FeatureVector::const_iterator it = F.mFeatVec.begin(); it != F.mFeatVec.end() && j < nf; ++it, ++j) {
            bad += it->first != fn[j] || (int)it->second.size() != fo[j + 1] - fo[j];
            for (size_t q = 0; q < it->second.size() && (int)q < fo[j + 1] - fo[j]; ++q) bad += it->second[q] != ff[fo[j] + q];
        }
        EXPECT(bad == 0, "vocabulary n=%d: %d entries differ from the oracle", n, bad);
        if (n == 1500) printf("vocabulary: %zu words, %zu nodes for %d features (oracle equal: %s)\n", F.mBowVec.size(), F.mFeatVec.size(), n, bad ? "no" : "yes");
    }
    orbo_voc_free(o);
}

// Tracking (Frame::ComputeBoW, Tracking.cc:874) and LocalMapping (KeyFrame::ComputeBoW, LocalMapping.cc:164) call transform on ONE
// shared ORBVocabulary from two threads: every thread must get the result of its own descriptors, for frames of different sizes
// (different scratch sizes, so a shared scratch buffer would be reallocated under the other thread).
static void test_vocabulary_two_threads() {
    VocRecords V;
    V.k = 5; V.L = 4;
    const uint8_t zero[32] = {0};
    grow(V, 0, zero, 1);
    const std::string path = write_voc(V);
    ORBVocabulary voc;
    EXPECT(voc.loadFromTextFile(path), "loadFromTextFile failed");
    remove(path.c_str());
    const int kFrames = 24;
    std::vector<cv::Mat> D(kFrames);
    for (int f = 0; f < kFrames; ++f) {
        const int n = 200 + (int)(rnd() % 1800);
        D[f] = cv::Mat(n, 32, CV_8U);
        for (int i = 0; i < n; ++i) {
            const size_t leaf = rnd() % V.parent.size();
            for (int b = 0; b < 32; ++b) D[f].at<uchar>(i, b) = (uint8_t)(V.desc[leaf * 32 + b] ^ ((rnd() % 9 == 0) ? (1u << (rnd() % 8)) : 0u));
        }
    }
    std::vector<DBoW2::BowVector> bv1(kFrames), bv2(kFrames);
    std::vector<DBoW2::FeatureVector> fv1(kFrames), fv2(kFrames);
    for (int f = 0; f < kFrames; ++f) voc.transform(toDescriptorVector(D[f]), bv1[f], fv1[f], 2);   // serial results
    std::string err;
    auto work = [&](int first) {
        try {
            for (int rep = 0; rep < 6; ++rep)
                for (int f = first; f < kFrames; f += 2) voc.transform(toDescriptorVector(D[f]), bv2[f], fv2[f], 2);
        } catch (const std::exception& e) { err = e.what(); }
    };
    std::thread a(work, 0), b(work, 1);
    a.join(); b.join();
    int bad = 0;
    for (int f = 0; f < kFrames; ++f) bad += !(bv1[f] == bv2[f]) || !(fv1[f] == fv2[f]);
    EXPECT(err.empty() && bad == 0, "vocabulary on two threads: %d of %d frames differ from the serial results %s", bad, kFrames, err.c_str());
    printf("vocabulary on two threads: %d frames x 6 rounds, %d differences\n", kFrames, bad);
}

int main() {
    int ndev = 0;
    if (orbgpu_device_count(&ndev) != 0 || ndev == 0) {
        // no GPU: every shell must fail loudly, never fall back
        int thrown = 0;
        try { ORBextractor ex(1000, 1.2f, 8, 20, 7); MiniFrame f; f.ExtractORB(&ex, synth_image(640, 480)); } catch (const std::runtime_error& e) { ++thrown; printf("extractor: %s\n", e.what()); }
        try {
            g_sf.assign(8, 1.f); g_s2.assign(8, 1.f);
            std::vector<cv::KeyPoint> k = random_keys(10, 640, 480);
            cv::Mat d = random_desc(10);
            KeyFrame a(k, std::vector<float>(10, -1.f), d, g_sf, g_s2, 1, 1, 0, 0), b(k, std::vector<float>(10, -1.f), d, g_sf, g_s2, 1, 1, 0, 0);
            ORBmatcher m; std::vector<MapPoint*> out; m.SearchByBoW(&a, &b, out);
        } catch (const std::runtime_error& e) { ++thrown; printf("matcher: %s\n", e.what()); }
        try {
            VocRecords V; V.k = 3; V.L = 2; const uint8_t z[32] = {0}; grow(V, 0, z, 1);
            const std::string path = write_voc(V);
            ORBVocabulary voc;
            struct Rm { std::string p; ~Rm() { remove(p.c_str()); } } rm = {path};
            voc.loadFromTextFile(path);
        } catch (const std::runtime_error& e) { ++thrown; printf("vocabulary: %s\n", e.what()); }
        return thrown == 3 ? 3 : 1;
    }
    test_extractor();
    test_stereo();
    test_two_threads();
    test_matcher();
    test_track_last_frame();
    test_vocabulary();
    test_vocabulary_two_threads();
    test_search_for_initialization();
    test_relocalization_search();
    test_fuse();
    printf(fails ? "shell_test: %d FAILURES\n" : "shell_test: all shell results equal the oracle (%d failures)\n", fails);
    return fails ? 1 : 0;
}

// ORBmatcher_gpu.cc — the four Hamming-path search functions of ORB_SLAM2::ORBmatcher as members of the reference's own
// class (declared in the reference's include/ORBmatcher.h), implemented on the GPU through the C ABI of include/orbgpu.h:
//
//   SearchByProjection(Frame&, const vector<MapPoint*>&, th)      replaces src/ORBmatcher.cc:59-155
//   SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)            replaces :211-344
//   SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)         replaces :635-768
//   SearchForTriangulation(KeyFrame*, KeyFrame*, F12, pairs, ..)  replaces :783-975
//   SearchByProjection(Frame&, const Frame&, th, bMono)           replaces :1540-1685 (pose arithmetic here, search on the GPU)
//   SearchForInitialization(Frame&, Frame&, vbPrevMatched, ...)   replaces :493-632
//   SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, d)  replaces :1711-1849 (Relocalization)
//   Fuse(KeyFrame*, const vector<MapPoint*>&, th)                  replaces :977-1137 (search on the GPU, map updates here)
//   SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)    replaces :359-491 (LoopClosing)
//   Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)             replaces :1139-1283 (LoopClosing)
//   SearchBySim3(KeyFrame*, KeyFrame*, vpMatches12, s12, R12, t12, th)  replaces :1285-1520 (LoopClosing)
//
// Integration: compile this file into the ORB_SLAM2 library and remove (or #ifdef out) those four bodies from the
// reference's ORBmatcher.cc; everything else of that file — the constructor, DescriptorDistance, the other search
// functions — stays the reference's CPU code.  Each function flattens the Frame / KeyFrame / MapPoint
// members the reference body reads into the views of orbgpu.h, makes ONE call, and writes the result back into the
// caller's containers exactly where the reference does.  No distance is computed on the host.
#include <cmath>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "ORBmatcher.h"
#include "orbgpu.h"

namespace ORB_SLAM2 {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbgpu_keypoint), "cv::KeyPoint must be the 28-byte record the C ABI reads");

namespace {

void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string("ORBmatcher (GPU): ") + what + ": " + orbgpu_last_error());
}

// ORBmatcher objects are short-lived stack objects (one per call site invocation); the device handle (stream, scratch)
// is kept per host thread instead — Tracking, LocalMapping and LoopClosing each get their own.
struct ThreadMatcher {
    orbgpu_matcher* m;
    ThreadMatcher() : m(nullptr) {}
    ~ThreadMatcher() { if (m) orbgpu_matcher_destroy(m); }
};
orbgpu_matcher* matcher() {
    thread_local ThreadMatcher t;
    if (!t.m) check(orbgpu_matcher_create(&t.m, orbgpu_default_device()), "cannot create the device matcher");
    return t.m;
}

// N x 32 descriptor rows as one contiguous block (DescriptorDistance reads rows as 8 int32, ORBmatcher.cc:1903-1904)
const uint8_t* rows32(const cv::Mat& d, std::vector<uint8_t>& tmp) {
    if (d.rows == 0) return tmp.data();
    if (d.isContinuous()) return d.ptr(0);
    tmp.resize((size_t)d.rows * 32);
    for (int i = 0; i < d.rows; ++i) std::memcpy(tmp.data() + (size_t)i * 32, d.ptr(i), 32);
    return tmp.data();
}

// std::map<NodeId, vector<unsigned>> -> CSR (ascending node ids, features in vector order)
struct FlatFeatVec {
    int32_t node_off[2];
    std::vector<int32_t> node_id, feat_off, feat;
    explicit FlatFeatVec(const DBoW2::FeatureVector& fv) {
        node_off[0] = 0;
        feat_off.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            node_id.push_back((int32_t)it->first);
            for (size_t k = 0; k < it->second.size(); ++k) feat.push_back((int32_t)it->second[k]);
            feat_off.push_back((int32_t)feat.size());
        }
        node_off[1] = (int32_t)node_id.size();
    }
    void attach(orbgpu_frame_set& s) const {
        s.fv_node_off = node_off;
        s.fv_node_id = node_id.data();
        s.fv_feat_off = feat_off.data();
        s.fv_feat = feat.data();
    }
};

orbgpu_frame_set one_frame(const int32_t* kp_off, const std::vector<cv::KeyPoint>& keys, const uint8_t* desc) {
    orbgpu_frame_set s;
    std::memset(&s, 0, sizeof(s));
    s.n_frames = 1;
    s.kp_off = kp_off;
    s.keys_un = reinterpret_cast<const orbgpu_keypoint*>(keys.data());
    s.desc = desc;
    return s;
}

const int32_t kZero = 0;
const int64_t kZero64 = 0;

}  // namespace

int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    const int N = (int)F.mvKeysUn.size(), M = (int)vpMapPoints.size();
    if (N == 0 || M == 0) return 0;
    const int32_t kp_off[2] = {0, N}, mp_off[2] = {0, M};
    std::vector<uint8_t> tmp, kflags(N, 0);
    orbgpu_frame_set fs = one_frame(kp_off, F.mvKeysUn, rows32(F.mDescriptors, tmp));
    for (int i = 0; i < N; ++i)   // :108-110: a keypoint holding a MapPoint with observations is skipped
        if (F.mvpMapPoints[i]) kflags[i] = F.mvpMapPoints[i]->Observations() > 0 ? 1 : 2;
    fs.kp_flags = kflags.data();
    fs.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
    const float grid[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    fs.grid = grid;

    std::vector<float> px(M), py(M), pxr(M), vc(M);
    std::vector<int32_t> lvl(M);
    std::vector<uint8_t> mflags(M, 0), mdesc((size_t)M * 32, 0);
    for (int q = 0; q < M; ++q) {
        MapPoint* p = vpMapPoints[q];
        if (!p->mbTrackInView) continue;                       // :70-71
        if (p->isBad()) { mflags[q] = 1 | 2; continue; }       // :73-74
        mflags[q] = (uint8_t)(1 | (p->Observations() > 0 ? 4 : 0));
        px[q] = p->mTrackProjX; py[q] = p->mTrackProjY; pxr[q] = p->mTrackProjXR;
        vc[q] = p->mTrackViewCos; lvl[q] = p->mnTrackScaleLevel;
        const cv::Mat d = p->GetDescriptor();                  // :93
        std::memcpy(&mdesc[(size_t)q * 32], d.ptr(0), 32);
    }
    orbgpu_mappoint_set ms;
    ms.mp_off = mp_off; ms.proj_x = px.data(); ms.proj_y = py.data(); ms.proj_xr = pxr.data(); ms.view_cos = vc.data();
    ms.level = lvl.data(); ms.flags = mflags.data(); ms.desc = mdesc.data();

    std::vector<int32_t> kp_match(N, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_by_projection(matcher(), &fs, &ms, F.mvScaleFactors.data(), (int)F.mvScaleFactors.size(), th, mfNNratio,
                                      kp_match.data(), nullptr, nullptr, nullptr, &nmatches),
          "SearchByProjection");
    for (int i = 0; i < N; ++i)
        if (kp_match[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[kp_match[i]];   // :149
    return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const int N1 = (int)pKF->mvKeysUn.size(), N2 = (int)F.mvKeys.size();
    if (N1 == 0 || N2 == 0) return 0;
    const int32_t off1[2] = {0, N1}, off2[2] = {0, N2};
    std::vector<uint8_t> t1, t2, flags1(N1, 0);
    for (int i = 0; i < N1; ++i) flags1[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) ? 1 : 0;   // :233-238
    orbgpu_frame_set s1 = one_frame(off1, pKF->mvKeysUn, rows32(pKF->mDescriptors, t1));
    orbgpu_frame_set s2 = one_frame(off2, F.mvKeys, rows32(F.mDescriptors, t2));   // the rotation check reads F.mvKeys (:294)
    s1.kp_flags = flags1.data();
    const FlatFeatVec fv1(pKF->mFeatVec), fv2(F.mFeatVec);
    fv1.attach(s1);
    fv2.attach(s2);
    std::vector<int32_t> match12(N1, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_by_bow(matcher(), &s1, &s2, 1, &kZero, &kZero, mfNNratio, mbCheckOrientation ? 1 : 0, TH_LOW, /*<=*/1,
                               /*require_mp2*/ 0, &kZero64, match12.data(), nullptr, &nmatches),
          "SearchByBoW(KeyFrame, Frame)");
    for (int i = 0; i < N1; ++i)
        if (match12[i] >= 0) vpMapPointMatches[match12[i]] = vpMapPointsKF[i];   // :288
    return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    const int N1 = (int)pKF1->mvKeysUn.size(), N2 = (int)pKF2->mvKeysUn.size();
    if (N1 == 0 || N2 == 0) return 0;
    const int32_t off1[2] = {0, N1}, off2[2] = {0, N2};
    std::vector<uint8_t> t1, t2, flags1(N1, 0), flags2(N2, 0);
    for (int i = 0; i < N1; ++i) flags1[i] = (vpMapPoints1[i] && !vpMapPoints1[i]->isBad()) ? 1 : 0;   // :673-677
    for (int i = 0; i < N2; ++i) flags2[i] = (vpMapPoints2[i] && !vpMapPoints2[i]->isBad()) ? 1 : 0;   // :688-694
    orbgpu_frame_set s1 = one_frame(off1, pKF1->mvKeysUn, rows32(pKF1->mDescriptors, t1));
    orbgpu_frame_set s2 = one_frame(off2, pKF2->mvKeysUn, rows32(pKF2->mDescriptors, t2));
    s1.kp_flags = flags1.data();
    s2.kp_flags = flags2.data();
    const FlatFeatVec fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    fv1.attach(s1);
    fv2.attach(s2);
    std::vector<int32_t> match12(N1, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_by_bow(matcher(), &s1, &s2, 1, &kZero, &kZero, mfNNratio, mbCheckOrientation ? 1 : 0, TH_LOW, /*<*/0,
                               /*require_mp2*/ 1, &kZero64, match12.data(), nullptr, &nmatches),
          "SearchByBoW(KeyFrame, KeyFrame)");
    for (int i = 0; i < N1; ++i)
        if (match12[i] >= 0) vpMatches12[i] = vpMapPoints2[match12[i]];   // :715
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    vMatchedPairs.clear();
    const int N1 = (int)pKF1->mvKeysUn.size(), N2 = (int)pKF2->mvKeysUn.size();
    if (N1 == 0 || N2 == 0) return 0;
    // epipole of camera 1 in image 2 (:790-799): C2 = R2w * Cw + t2w.  A 3x3 * 3x1 + 3x1 cv::Mat expression is one cv::gemm call
    // on its small-matrix path: the dot product is accumulated in float, left to right, the addend joins in double
    // (pinned to cv2 4.13: tests/golden/cvsmall_golden.npz, oracle/cvlite.cc cvl_gemm3_f32).
    const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    float C2[3];
    for (int i = 0; i < 3; ++i) {
        const float t = R2w.at<float>(i, 0) * Cw.at<float>(0, 0) + R2w.at<float>(i, 1) * Cw.at<float>(1, 0) + R2w.at<float>(i, 2) * Cw.at<float>(2, 0);
        C2[i] = (float)((double)t + (double)t2w.at<float>(i, 0));
    }
    const float invz = 1.0f / C2[2];
    const float epipole[2] = {pKF2->fx * C2[0] * invz + pKF2->cx, pKF2->fy * C2[1] * invz + pKF2->cy};
    float f12[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) f12[3 * i + j] = F12.at<float>(i, j);

    const int32_t off1[2] = {0, N1}, off2[2] = {0, N2};
    std::vector<uint8_t> t1, t2, flags1(N1, 0), flags2(N2, 0);
    for (int i = 0; i < N1; ++i) flags1[i] = pKF1->GetMapPoint(i) ? 1 : 0;   // :846
    for (int i = 0; i < N2; ++i) flags2[i] = pKF2->GetMapPoint(i) ? 1 : 0;   // :868
    orbgpu_frame_set s1 = one_frame(off1, pKF1->mvKeysUn, rows32(pKF1->mDescriptors, t1));
    orbgpu_frame_set s2 = one_frame(off2, pKF2->mvKeysUn, rows32(pKF2->mDescriptors, t2));
    s1.kp_flags = flags1.data();
    s2.kp_flags = flags2.data();
    s1.u_right = pKF1->mvuRight.empty() ? nullptr : pKF1->mvuRight.data();
    s2.u_right = pKF2->mvuRight.empty() ? nullptr : pKF2->mvuRight.data();
    const FlatFeatVec fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    fv1.attach(s1);
    fv2.attach(s2);
    std::vector<int32_t> match12(N1, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_for_triangulation(matcher(), &s1, &s2, 1, &kZero, &kZero, f12, epipole, pKF2->mvScaleFactors.data(),
                                          pKF2->mvLevelSigma2.data(), (int)pKF2->mvScaleFactors.size(), bOnlyStereo ? 1 : 0,
                                          mbCheckOrientation ? 1 : 0, &kZero64, match12.data(), nullptr, &nmatches),
          "SearchForTriangulation");
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < N1; ++i)
        if (match12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)match12[i]));   // :964-972
    return nmatches;
}

// 3x3 (rows 0-2, cols 0-2 of a 4x4 pose) times a 3-vector plus a 3-vector, the way cv::gemm evaluates the reference's
// cv::Mat expressions Rcw*x3Dw+tcw (ORBmatcher.cc:1556-1563, :1579): float dot product left to right, addend joined in double.
static void rt_apply(const cv::Mat& T, const float* x, float* out) {
    for (int i = 0; i < 3; ++i) {
        const float t = T.at<float>(i, 0) * x[0] + T.at<float>(i, 1) * x[1] + T.at<float>(i, 2) * x[2];
        out[i] = (float)((double)t + (double)T.at<float>(i, 3));
    }
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const int N = (int)CurrentFrame.mvKeysUn.size(), NL = LastFrame.N;
    if (N == 0 || NL == 0) return 0;
    // twc = -Rcw^T * tcw (a transposed operand takes cv::gemm's general path: double accumulation), tlc = Rlw * twc + tlw (:1556-1563)
    float twc[3], tlc[3];
    for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
        for (int j = 0; j < 3; ++j) acc += (double)CurrentFrame.mTcw.at<float>(j, i) * (double)CurrentFrame.mTcw.at<float>(j, 3);
        twc[i] = (float)(-acc);
    }
    rt_apply(LastFrame.mTcw, twc, tlc);
    const bool bForward = tlc[2] > CurrentFrame.mb && !bMono;     // :1566-1567
    const bool bBackward = -tlc[2] > CurrentFrame.mb && !bMono;

    const int32_t kp_off[2] = {0, N}, q_off[2] = {0, NL};
    std::vector<uint8_t> tmp, kflags(N, 0);
    orbgpu_frame_set fs = one_frame(kp_off, CurrentFrame.mvKeysUn, rows32(CurrentFrame.mDescriptors, tmp));
    for (int i = 0; i < N; ++i)
        if (CurrentFrame.mvpMapPoints[i]) kflags[i] = CurrentFrame.mvpMapPoints[i]->Observations() > 0 ? 1 : 2;   // :1619-1621
    fs.kp_flags = kflags.data();
    fs.u_right = CurrentFrame.mvuRight.empty() ? nullptr : CurrentFrame.mvuRight.data();
    const float grid[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    fs.grid = grid;

    std::vector<float> qu(NL), qv(NL), qr(NL), qur(NL), qang(NL);
    std::vector<int32_t> qlo(NL), qhi(NL);
    std::vector<uint8_t> qfl(NL, 0), qdesc((size_t)NL * 32, 0);
    for (int i = 0; i < NL; ++i) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;            // :1572-1575
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const float xw[3] = {x3Dw.at<float>(0, 0), x3Dw.at<float>(1, 0), x3Dw.at<float>(2, 0)};
        float xc[3];
        rt_apply(CurrentFrame.mTcw, xw, xc);                      // :1579
        const float invzc = 1.0 / xc[2];
        if (invzc < 0) continue;
        const float u = Frame::fx * xc[0] * invzc + Frame::cx;
        const float v = Frame::fy * xc[1] * invzc + Frame::cy;
        if (u < Frame::mnMinX || u > Frame::mnMaxX) continue;
        if (v < Frame::mnMinY || v > Frame::mnMaxY) continue;
        const int nLastOctave = LastFrame.mvKeys[i].octave;
        qu[i] = u; qv[i] = v;
        qr[i] = th * CurrentFrame.mvScaleFactors[nLastOctave];   // :1598
        if (bForward) { qlo[i] = nLastOctave; qhi[i] = -1; }      // GetFeaturesInArea(u, v, radius, nLastOctave)
        else if (bBackward) { qlo[i] = 0; qhi[i] = nLastOctave; }
        else { qlo[i] = nLastOctave - 1; qhi[i] = nLastOctave + 1; }
        qur[i] = u - CurrentFrame.mbf * invzc;                    // :1626
        qang[i] = LastFrame.mvKeysUn[i].angle;
        qfl[i] = (uint8_t)(1 | (pMP->Observations() > 0 ? 4 : 0));
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&qdesc[(size_t)i * 32], d.ptr(0), 32);
    }
    orbgpu_window_query_set qs;
    qs.q_off = q_off; qs.u = qu.data(); qs.v = qv.data(); qs.radius = qr.data(); qs.min_level = qlo.data(); qs.max_level = qhi.data();
    qs.ur = qur.data(); qs.flags = qfl.data(); qs.desc = qdesc.data(); qs.angle = qang.data();
    std::vector<int32_t> kp_match(N, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_windowed(matcher(), &fs, &qs, TH_HIGH, /*skip_any_mappoint*/ 0, mbCheckOrientation ? 1 : 0, kp_match.data(), nullptr,
                                 nullptr, &nmatches),
          "SearchByProjection(Frame, Frame)");
    for (int i = 0; i < N; ++i) {
        if (kp_match[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[kp_match[i]];   // :1644
        else if (kp_match[i] == -2) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);    // :1676
    }
    return nmatches;
}

// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1711-1849): the projection search of
// Tracking::Relocalization.  The per-map-point pose arithmetic (:1713-1757) stays here, in cv::gemm's evaluation order; the
// windowed search, the "key point already holds a MapPoint" rule (:1776-1777) and the rotation check run on the GPU.
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    const int N = (int)CurrentFrame.mvKeysUn.size();
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int NQ = (int)vpMPs.size();
    if (N == 0 || NQ == 0) return 0;
    // Ow = -Rcw^T * tcw (a transposed operand takes cv::gemm's general path: double accumulation)
    float Ow[3];
    for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
        for (int j = 0; j < 3; ++j) acc += (double)CurrentFrame.mTcw.at<float>(j, i) * (double)CurrentFrame.mTcw.at<float>(j, 3);
        Ow[i] = (float)(-acc);
    }
    const int32_t kp_off[2] = {0, N}, q_off[2] = {0, NQ};
    std::vector<uint8_t> tmp, kflags(N, 0);
    orbgpu_frame_set fs = one_frame(kp_off, CurrentFrame.mvKeysUn, rows32(CurrentFrame.mDescriptors, tmp));
    for (int i = 0; i < N; ++i)
        if (CurrentFrame.mvpMapPoints[i]) kflags[i] = 2;   // any MapPoint blocks the key point (:1776-1777)
    fs.kp_flags = kflags.data();
    const float grid[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    fs.grid = grid;
    std::vector<float> qu(NQ), qv(NQ), qr(NQ), qang(NQ);
    std::vector<int32_t> qlo(NQ), qhi(NQ);
    std::vector<uint8_t> qfl(NQ, 0), qdesc((size_t)NQ * 32, 0);
    for (int i = 0; i < NQ; ++i) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;   // :1727-1731
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const float xw[3] = {x3Dw.at<float>(0, 0), x3Dw.at<float>(1, 0), x3Dw.at<float>(2, 0)};
        float xc[3];
        rt_apply(CurrentFrame.mTcw, xw, xc);                              // :1735
        const float invzc = 1.0 / xc[2];
        const float u = Frame::fx * xc[0] * invzc + Frame::cx;
        const float v = Frame::fy * xc[1] * invzc + Frame::cy;
        if (u < Frame::mnMinX || u > Frame::mnMaxX) continue;
        if (v < Frame::mnMinY || v > Frame::mnMaxY) continue;
        // dist3D = cv::norm(x3Dw - Ow): float differences, squares summed in double (:1752-1753)
        double s2 = 0.0;
        for (int k = 0; k < 3; ++k) { const float d = xw[k] - Ow[k]; s2 += (double)d * (double)d; }
        float dist3D = (float)std::sqrt(s2);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
        qu[i] = u; qv[i] = v;
        qr[i] = th * CurrentFrame.mvScaleFactors[nPredictedLevel];       // :1764
        qlo[i] = nPredictedLevel - 1; qhi[i] = nPredictedLevel + 1;
        qang[i] = pKF->mvKeysUn[i].angle;
        qfl[i] = 1;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&qdesc[(size_t)i * 32], d.ptr(0), 32);
    }
    orbgpu_window_query_set qs;
    qs.q_off = q_off; qs.u = qu.data(); qs.v = qv.data(); qs.radius = qr.data(); qs.min_level = qlo.data(); qs.max_level = qhi.data();
    qs.ur = nullptr; qs.flags = qfl.data(); qs.desc = qdesc.data(); qs.angle = qang.data();
    std::vector<int32_t> kp_match(N, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_windowed(matcher(), &fs, &qs, ORBdist, /*skip_any_mappoint*/ 1, mbCheckOrientation ? 1 : 0, kp_match.data(), nullptr,
                                 nullptr, &nmatches),
          "SearchByProjection(Frame, KeyFrame)");
    for (int i = 0; i < N; ++i) {
        if (kp_match[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[kp_match[i]];                   // :1795
        else if (kp_match[i] == -2) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);    // :1841
    }
    return nmatches;
}

// Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:977-1137, LocalMapping::SearchInNeighbors): project the map points into the key
// frame, look for the best key point in a window with the chi-square gate — that search runs on the GPU for all points at
// once (orbgpu_search_window_best: queries are independent, candidates are not filtered by the MapPoints they hold) — then
// apply Replace / AddObservation in vector order here, with isBad() / IsInKeyFrame() evaluated at each point's turn exactly
// as the reference's loop does (an earlier Replace can turn a later point bad).
int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    const int N = (int)pKF->mvKeysUn.size(), NQ = (int)vpMapPoints.size();
    if (N == 0 || NQ == 0) return 0;
    const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Owm = pKF->GetCameraCenter();
    float R[9], t[3], Ow[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[3 * i + j] = Rcw.at<float>(i, j);
        t[i] = tcw.at<float>(i, 0);
        Ow[i] = Owm.at<float>(i, 0);
    }
    const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy, bf = pKF->mbf;
    const int32_t kp_off[2] = {0, N}, q_off[2] = {0, NQ};
    std::vector<uint8_t> tmp;
    orbgpu_frame_set fs = one_frame(kp_off, pKF->mvKeysUn, rows32(pKF->mDescriptors, tmp));
    fs.u_right = pKF->mvuRight.empty() ? nullptr : pKF->mvuRight.data();
    const float grid[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv};
    fs.grid = grid;
    std::vector<float> qu(NQ), qv(NQ), qr(NQ), qur(NQ);
    std::vector<int32_t> qlo(NQ), qhi(NQ);
    std::vector<uint8_t> qfl(NQ, 0), qdesc((size_t)NQ * 32, 0);
    for (int i = 0; i < NQ; ++i) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float xw[3] = {p3Dw.at<float>(0, 0), p3Dw.at<float>(1, 0), p3Dw.at<float>(2, 0)};
        float pc[3];
        for (int r = 0; r < 3; ++r) {   // Rcw*p3Dw + tcw: cv::gemm's small-matrix path (float dot product, addend joined in double)
            const float d = R[3 * r] * xw[0] + R[3 * r + 1] * xw[1] + R[3 * r + 2] * xw[2];
            pc[r] = (float)((double)d + (double)t[r]);
        }
        if (pc[2] < 0.0f) continue;                                   // :1011-1012
        const float invz = 1 / pc[2];
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) continue;                          // :1023-1024
        const float ur = u - bf * invz;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        const float PO[3] = {xw[0] - Ow[0], xw[1] - Ow[1], xw[2] - Ow[2]};
        const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);   // cv::norm
        if (dist3D < minDistance || dist3D > maxDistance) continue;  // :1035-1036
        const cv::Mat Pn = pMP->GetNormal();
        const double dot = (double)PO[0] * Pn.at<float>(0, 0) + (double)PO[1] * Pn.at<float>(1, 0) + (double)PO[2] * Pn.at<float>(2, 0);
        if (dot < 0.5 * dist3D) continue;                             // :1041-1042
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        qu[i] = u; qv[i] = v; qur[i] = ur;
        qr[i] = th * pKF->mvScaleFactors[nPredictedLevel];           // :1047
        qlo[i] = nPredictedLevel - 1; qhi[i] = nPredictedLevel;       // :1066-1067
        qfl[i] = 1;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&qdesc[(size_t)i * 32], d.ptr(0), 32);
    }
    orbgpu_window_query_set qs;
    qs.q_off = q_off; qs.u = qu.data(); qs.v = qv.data(); qs.radius = qr.data(); qs.min_level = qlo.data(); qs.max_level = qhi.data();
    qs.ur = qur.data(); qs.flags = qfl.data(); qs.desc = qdesc.data(); qs.angle = nullptr;
    std::vector<int32_t> best(NQ, -1), bdist(NQ, 256);
    check(orbgpu_search_window_best(matcher(), &fs, &qs, pKF->mvInvLevelSigma2.data(), (int)pKF->mvInvLevelSigma2.size(), 0, best.data(), bdist.data()),
          "Fuse");
    int nFused = 0;
    for (int i = 0; i < NQ; ++i) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || !qfl[i]) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;        // :999-1000, at this point's turn
        if (bdist[i] > TH_LOW) continue;                              // :1114
        const int bestIdx = best[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

// SearchForInitialization (ORBmatcher.cc:493-632): the monocular bootstrap match between the reference frame F1 and the current
// frame F2 inside a window around vbPrevMatched.  The stealing rule and the rotation check run on the device; vbPrevMatched is
// updated here exactly as at :626-628.
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                        int windowSize) {
    const int N1 = (int)F1.mvKeysUn.size(), N2 = (int)F2.mvKeysUn.size();
    vnMatches12 = std::vector<int>(N1, -1);
    if (N1 == 0 || N2 == 0) return 0;
    const int32_t kp_off[2] = {0, N2}, q_off[2] = {0, N1};
    std::vector<uint8_t> tmp2, tmp1;
    orbgpu_frame_set fs = one_frame(kp_off, F2.mvKeysUn, rows32(F2.mDescriptors, tmp2));
    const float grid[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    fs.grid = grid;
    std::vector<float> qu(N1), qv(N1), qr(N1, (float)windowSize), qang(N1);
    std::vector<int32_t> qlo(N1, 0), qhi(N1, 0);   // GetFeaturesInArea(x, y, windowSize, level1, level1) with level1 == 0 (:512-519)
    std::vector<uint8_t> qfl(N1);
    for (int i = 0; i < N1; ++i) {
        qu[i] = vbPrevMatched[i].x;
        qv[i] = vbPrevMatched[i].y;
        qang[i] = F1.mvKeysUn[i].angle;
        qfl[i] = F1.mvKeysUn[i].octave > 0 ? 0 : 1;   // level1 > 0: continue (:513-514)
    }
    orbgpu_window_query_set qs;
    qs.q_off = q_off; qs.u = qu.data(); qs.v = qv.data(); qs.radius = qr.data(); qs.min_level = qlo.data(); qs.max_level = qhi.data();
    qs.ur = nullptr; qs.flags = qfl.data(); qs.desc = rows32(F1.mDescriptors, tmp1); qs.angle = qang.data();
    std::vector<int32_t> m12(N1, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_for_initialization(matcher(), &fs, &qs, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches),
          "SearchForInitialization");
    for (int i = 0; i < N1; ++i) {
        vnMatches12[i] = m12[i];
        if (m12[i] >= 0) vbPrevMatched[i] = F2.mvKeysUn[m12[i]].pt;   // :626-628
    }
    return nmatches;
}

namespace {
// One side of the Sim3-guided searches: every MapPoint of `points` with a pose-independent reason to be searched (`live`) is
// projected by `project` (the reference's own cv::Mat expressions), gated exactly as the reference gates it, and turned into a
// window query on pKF's key points.  The best-candidate search of all queries is ONE device call.
struct BestSearch {
    std::vector<float> qu, qv, qr;
    std::vector<int32_t> qlo, qhi, best, bdist;
    std::vector<uint8_t> qfl, qdesc;
    explicit BestSearch(int n) : qu(n), qv(n), qr(n), qlo(n), qhi(n), best(n, -1), bdist(n, 256), qfl(n, 0), qdesc((size_t)n * 32, 0) {}
    void add(int i, float u, float v, float radius, int level, MapPoint* pMP) {
        qu[i] = u; qv[i] = v; qr[i] = radius;
        qlo[i] = level - 1; qhi[i] = level;   // kpLevel < nPredictedLevel-1 || kpLevel > nPredictedLevel: continue
        qfl[i] = 1;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&qdesc[(size_t)i * 32], d.ptr(0), 32);
    }
    void run(KeyFrame* pKF, const char* what) {
        const int N = (int)pKF->mvKeysUn.size(), NQ = (int)qu.size();
        if (N == 0 || NQ == 0) return;
        const int32_t kp_off[2] = {0, N}, q_off[2] = {0, NQ};
        std::vector<uint8_t> tmp;
        orbgpu_frame_set fs = one_frame(kp_off, pKF->mvKeysUn, rows32(pKF->mDescriptors, tmp));
        const float grid[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv};
        fs.grid = grid;
        orbgpu_window_query_set qs;
        qs.q_off = q_off; qs.u = qu.data(); qs.v = qv.data(); qs.radius = qr.data(); qs.min_level = qlo.data(); qs.max_level = qhi.data();
        qs.ur = nullptr; qs.flags = qfl.data(); qs.desc = qdesc.data(); qs.angle = nullptr;
        check(orbgpu_search_window_best(matcher(), &fs, &qs, nullptr, 0, 0, best.data(), bdist.data()), what);
    }
};

// Scw -> Rcw, tcw, Ow exactly as the reference writes it (ORBmatcher.cc:368-374, :1148-1153)
void decompose_sim3(const cv::Mat& Scw, cv::Mat& Rcw, cv::Mat& tcw, cv::Mat& Ow) {
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    Rcw = sRcw / scw;
    tcw = Scw.rowRange(0, 3).col(3) / scw;
    Ow = -Rcw.t() * tcw;
}

// The gates shared by SearchByProjection(KF, Scw), Fuse(KF, Scw) (:393-436, :1175-1216): projection, image bounds, distance
// range, viewing angle.  Returns false where the reference `continue`s.
bool project_with_normal(KeyFrame* pKF, MapPoint* pMP, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, float& u, float& v, float& dist3D) {
    cv::Mat p3Dw = pMP->GetWorldPos();
    cv::Mat p3Dc = Rcw * p3Dw + tcw;
    if (p3Dc.at<float>(2) < 0.0) return false;
    const float invz = 1 / p3Dc.at<float>(2);
    const float x = p3Dc.at<float>(0) * invz;
    const float y = p3Dc.at<float>(1) * invz;
    u = pKF->fx * x + pKF->cx;
    v = pKF->fy * y + pKF->cy;
    if (!pKF->IsInImage(u, v)) return false;
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    cv::Mat PO = p3Dw - Ow;
    dist3D = cv::norm(PO);
    if (dist3D < minDistance || dist3D > maxDistance) return false;
    cv::Mat Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist3D) return false;
    return true;
}
}  // namespace

// SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:359-491, LoopClosing::ComputeSim3): the loop map points
// projected with the Sim3 pose; a key point that already has a match (before the call or from an earlier point of this call)
// is skipped (:453-454), so the search is the greedy windowed one (orbgpu_search_windowed, every occupied key point blocked).
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
    const int N = (int)pKF->mvKeysUn.size(), NQ = (int)vpPoints.size();
    if (N == 0 || NQ == 0) return 0;
    cv::Mat Rcw, tcw, Ow;
    decompose_sim3(Scw, Rcw, tcw, Ow);
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    const int32_t kp_off[2] = {0, N}, q_off[2] = {0, NQ};
    std::vector<uint8_t> tmp, kflags(N, 0);
    orbgpu_frame_set fs = one_frame(kp_off, pKF->mvKeysUn, rows32(pKF->mDescriptors, tmp));
    for (int i = 0; i < N; ++i)
        if (vpMatched[i]) kflags[i] = 2;
    fs.kp_flags = kflags.data();
    const float grid[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv};
    fs.grid = grid;
    BestSearch Q(NQ);
    for (int i = 0; i < NQ; ++i) {
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;      // :389-390
        float u, v, dist;
        if (!project_with_normal(pKF, pMP, Rcw, tcw, Ow, u, v, dist)) continue;
        const int nPredictedLevel = pMP->PredictScale(dist, pKF);
        Q.add(i, u, v, th * pKF->mvScaleFactors[nPredictedLevel], nPredictedLevel, pMP);   // :435
    }
    orbgpu_window_query_set qs;
    qs.q_off = q_off; qs.u = Q.qu.data(); qs.v = Q.qv.data(); qs.radius = Q.qr.data(); qs.min_level = Q.qlo.data(); qs.max_level = Q.qhi.data();
    qs.ur = nullptr; qs.flags = Q.qfl.data(); qs.desc = Q.qdesc.data(); qs.angle = nullptr;
    std::vector<int32_t> kp_match(N, -1);
    int32_t nmatches = 0;
    check(orbgpu_search_windowed(matcher(), &fs, &qs, TH_LOW, /*skip_any_mappoint*/ 1, /*check_orientation*/ 0, kp_match.data(), nullptr, nullptr,
                                 &nmatches),
          "SearchByProjection(KeyFrame, Scw)");
    for (int i = 0; i < N; ++i)
        if (kp_match[i] >= 0) vpMatched[i] = vpPoints[kp_match[i]];    // :473
    return nmatches;
}

// Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1139-1283, LoopClosing::SearchAndFuse): like Fuse(pKF, vpMapPoints,
// th) without the chi-square gate; a point whose best key point already holds a MapPoint is reported for replacement instead of
// being replaced here.  The search runs on the GPU for all points at once, the map updates in vector order here (a key point
// filled by an earlier point of this call is seen by the later ones, :1262-1272).
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    const int N = (int)pKF->mvKeysUn.size(), NQ = (int)vpPoints.size();
    if (N == 0 || NQ == 0) return 0;
    cv::Mat Rcw, tcw, Ow;
    decompose_sim3(Scw, Rcw, tcw, Ow);
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();   // :1157, taken before the loop
    BestSearch Q(NQ);
    for (int i = 0; i < NQ; ++i) {
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;      // :1170-1171 (nothing in this loop changes either)
        float u, v, dist3D;
        if (!project_with_normal(pKF, pMP, Rcw, tcw, Ow, u, v, dist3D)) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        Q.add(i, u, v, th * pKF->mvScaleFactors[nPredictedLevel], nPredictedLevel, pMP);   // :1215
    }
    Q.run(pKF, "Fuse(KeyFrame, Scw)");
    int nFused = 0;
    for (int i = 0; i < NQ; ++i) {
        if (!Q.qfl[i] || Q.bdist[i] > TH_LOW) continue;                // :1259
        MapPoint* pMP = vpPoints[i];
        const int bestIdx = Q.best[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

// SearchBySim3 (ORBmatcher.cc:1285-1520, LoopClosing::ComputeSim3): the map points of each key frame projected into the other
// with the Sim3 estimate, best key point within a window (two device calls, one per direction), then the mutual-agreement
// check (:1497-1517) here.
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
    const float &fx = pKF1->fx, &fy = pKF1->fy, &cx = pKF1->cx, &cy = pKF1->cy;
    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;                    // :1304-1308
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;

    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    // direction 1 -> 2 (:1337-1401)
    BestSearch Q1(N1);
    for (int i1 = 0; i1 < N1; i1++) {
        MapPoint* pMP = vpMapPoints1[i1];
        if (!pMP || vbAlreadyMatched1[i1]) continue;
        if (pMP->isBad()) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc1 = R1w * p3Dw + t1w;
        cv::Mat p3Dc2 = sR21 * p3Dc1 + t21;
        if (p3Dc2.at<float>(2) < 0.0) continue;
        const float invz = 1.0 / p3Dc2.at<float>(2);
        const float x = p3Dc2.at<float>(0) * invz;
        const float y = p3Dc2.at<float>(1) * invz;
        const float u = fx * x + cx;
        const float v = fy * y + cy;
        if (!pKF2->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        const float dist3D = cv::norm(p3Dc2);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF2);
        Q1.add(i1, u, v, th * pKF2->mvScaleFactors[nPredictedLevel], nPredictedLevel, pMP);
    }
    Q1.run(pKF2, "SearchBySim3 (1 -> 2)");
    // direction 2 -> 1 (:1417-1481)
    BestSearch Q2(N2);
    for (int i2 = 0; i2 < N2; i2++) {
        MapPoint* pMP = vpMapPoints2[i2];
        if (!pMP || vbAlreadyMatched2[i2]) continue;
        if (pMP->isBad()) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc2 = R2w * p3Dw + t2w;
        cv::Mat p3Dc1 = sR12 * p3Dc2 + t12;
        if (p3Dc1.at<float>(2) < 0.0) continue;
        const float invz = 1.0 / p3Dc1.at<float>(2);
        const float x = p3Dc1.at<float>(0) * invz;
        const float y = p3Dc1.at<float>(1) * invz;
        const float u = fx * x + cx;
        const float v = fy * y + cy;
        if (!pKF1->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        const float dist3D = cv::norm(p3Dc1);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF1);
        Q2.add(i2, u, v, th * pKF1->mvScaleFactors[nPredictedLevel], nPredictedLevel, pMP);
    }
    Q2.run(pKF1, "SearchBySim3 (2 -> 1)");
    // agreement (:1497-1517)
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = (Q1.qfl[i1] && Q1.bdist[i1] <= TH_HIGH) ? Q1.best[i1] : -1;
        if (idx2 >= 0) {
            const int idx1 = (Q2.qfl[idx2] && Q2.bdist[idx2] <= TH_HIGH) ? Q2.best[idx2] : -1;
            if (idx1 == i1) {
                vpMatches12[i1] = vpMapPoints2[idx2];
                nFound++;
            }
        }
    }
    return nFound;
}

#if defined(ORBGPU_SHELL_STANDALONE) || defined(ORBGPU_SHELL_REPLACES_ORBMATCHER_CC)
// Builds without the reference's ORBmatcher.cc (tests; or a deployment that drops that file altogether, since all eleven search
// members live here): the remaining members of the class.
const int ORBmatcher::TH_HIGH = ORBGPU_TH_HIGH;
const int ORBmatcher::TH_LOW = ORBGPU_TH_LOW;
const int ORBmatcher::HISTO_LENGTH = ORBGPU_HISTO_LENGTH;
ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1901-1917) for the CPU callers that stay CPU code (Frame::ComputeStereoMatches when
// the extractor hook is not used, MapPoint::ComputeDistinctiveDescriptors): 256-bit Hamming distance of two descriptor rows.
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    const uint32_t* pa = a.ptr<uint32_t>();
    const uint32_t* pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; ++i) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}
#endif

}  // namespace ORB_SLAM2

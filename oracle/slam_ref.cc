// slam_ref — TEST INFRASTRUCTURE.  The reference's own ORBmatcher, Frame, KeyFrame, MapPoint, Map and ORBextractor,
// compiled from the sources where they lie (oracle/Makefile, target slamref: /root/reference/src/{ORBmatcher,Frame,KeyFrame,
// MapPoint,Map,KeyFrameDatabase,ORBextractor}.cc + Thirdparty/DBoW2, against the cv:: stand-in of shim/), behind a C wrapper
// that takes the SAME flat views as the C ABI (include/orbgpu.h) and as the port (oracle/match_oracle.cc): each entry
// builds real Frame / KeyFrame / MapPoint objects from the views, calls the reference member function, and flattens what
// that function wrote.  Nothing of the reference is copied or restated here; this file only constructs inputs and reads
// outputs.  It pins the port and the CUDA path: tests compare both with this library.
//
// Where the flat view holds values the reference computes itself from a camera pose (projections of the windowed
// searches, the epipole of SearchForTriangulation), the wrapper builds the pose that makes the reference's own float
// arithmetic reproduce those values exactly: identity rotation, zero translation, fx = fy = 1, cx = cy = 0 and world
// points (u, v, 1) — 1*u + 0*v + 0*1 and u*1 + 0 are exact in float, so the reference projects to exactly (u, v).
// The full pose arithmetic is exercised one level up (tests/cpp/ref_twin_test.cc: shell vs reference on shared worlds).
//
// The member access below (private / protected opened for THIS translation unit only) sets fields that the reference
// fills through code outside the hot path (constructors that run a full extraction, UpdateNormalAndDepth, ...).
#include "slam_world.h"

#include <chrono>

using namespace slamworld;

namespace {

std::unordered_map<MapPoint*, int> index_of(const std::vector<MapPoint*>& v) {
    std::unordered_map<MapPoint*, int> m;
    for (size_t i = 0; i < v.size(); ++i) if (v[i]) m[v[i]] = (int)i;
    return m;
}

}  // namespace

extern "C" {

int slamref_hamming(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:59-155.  The best / second-best
// distances are internal to the reference function: mp_best_* are left untouched.
void slamref_search_by_projection(const orbgpu_frame_set* fs, const orbgpu_mappoint_set* mp, const float* scale, int n_levels, float th,
                                  float nnratio, int32_t* kp_match, int32_t*, int32_t*, int32_t*, int32_t* nmatches_out) {
    for (int f = 0; f < fs->n_frames; ++f) {
        World W;
        const int k0 = fs->kp_off[f], n = fs->kp_off[f + 1] - k0;
        Frame* F = W.frame(fs, f, scale, n_levels, nullptr, identity_camera(fs->grid + 4 * f), true);
        if (fs->kp_flags)
            for (int i = 0; i < n; ++i)
                if (fs->kp_flags[k0 + i]) F->mvpMapPoints[i] = W.occupant(fs->kp_flags[k0 + i] == 1 ? 1 : 0);
        std::vector<MapPoint*> v;
        for (int q = mp->mp_off[f]; q < mp->mp_off[f + 1]; ++q) {
            MapPoint* p = W.mappoint(vec3(0, 0, 0), mp->desc + (size_t)q * 32, (mp->flags[q] & 4) ? 1 : 0, (mp->flags[q] & 2) != 0);
            p->mbTrackInView = (mp->flags[q] & 1) != 0;
            p->mTrackProjX = mp->proj_x[q]; p->mTrackProjY = mp->proj_y[q];
            p->mTrackProjXR = mp->proj_xr ? mp->proj_xr[q] : 0.f;
            p->mTrackViewCos = mp->view_cos[q];
            p->mnTrackScaleLevel = mp->level[q];
            v.push_back(p);
        }
        ORBmatcher matcher(nnratio, true);
        const int nm = matcher.SearchByProjection(*F, v, th);
        std::unordered_map<MapPoint*, int> idx = index_of(v);
        for (int i = 0; i < n; ++i) {
            std::unordered_map<MapPoint*, int>::iterator it = idx.find(F->mvpMapPoints[i]);
            if (kp_match) kp_match[k0 + i] = it == idx.end() ? -1 : it->second;
        }
        if (nmatches_out) nmatches_out[f] = nm;
    }
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1540-1685;
// skip_any == 0) and SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) (:1711-1849; skip_any != 0)
// on a world built so that every live query of the view projects to exactly (u, v) with exactly its radius and level range.
// Preconditions on the view (violations return a negative code): radius[q] == th * scale[l] for one level l; the queries of
// a frame share one level-range mode — [l-1, l+1], forward [l, -1] or backward [0, l] (first variant; the second variant
// only has [l-1, l+1]); ur[q] == u[q] - mbf; u >= mnMinX, v >= mnMinY; th_dist == TH_HIGH for the first variant.
// th and mbf come as extra arguments after the port's signature.  kp_match: index of the query whose MapPoint the key point
// holds after the call; -2 where a key point that held a MapPoint before is NULL now (reset by the rotation check); else -1
// (a reset key point that was NULL before is indistinguishable from an untouched one).
int slamref_search_windowed(const orbgpu_frame_set* fs, const orbgpu_window_query_set* qs, int th_dist, int skip_any, int check_orientation,
                            int32_t* kp_match, int32_t*, int32_t*, int32_t* nmatches_out, const float* scale, int n_levels, float th, float mbf) {
    const float log_sf = std::log(scale[1]);
    if (!skip_any && th_dist != ORBmatcher::TH_HIGH) return -6;
    for (int f = 0; f < fs->n_frames; ++f) {
        World W;
        const int k0 = fs->kp_off[f], n = fs->kp_off[f + 1] - k0;
        const int q0 = qs->q_off[f], nq = qs->q_off[f + 1] - q0;
        Camera cam = identity_camera(fs->grid + 4 * f);
        cam.mbf = mbf;
        cam.mb = 1.f;
        // level of every live query, then the frame's level-range mode
        std::vector<int> qlevel(nq, 0);
        int mode = 0;   // 0: [l-1, l+1], 1: forward [l, -1], 2: backward [0, l]
        for (int j = 0; j < nq; ++j) {
            const int q = q0 + j;
            if (!(qs->flags[q] & 1)) continue;
            int lvl = -1;
            for (int l = 0; l < n_levels; ++l) if (qs->radius[q] == th * scale[l]) { lvl = l; break; }
            if (lvl < 0) return -1;
            qlevel[j] = lvl;
            if (qs->max_level[q] == -1) mode = 1;
            else if (qs->max_level[q] == lvl && mode == 0) mode = 2;
            if (qs->u[q] < cam.minX || qs->v[q] < cam.minY) return -3;
            if (qs->ur && qs->ur[q] != qs->u[q] - mbf) return -4;
        }
        if (skip_any && mode != 0) return -7;
        for (int j = 0; j < nq; ++j) {
            const int q = q0 + j, l = qlevel[j];
            if (!(qs->flags[q] & 1)) continue;
            const int lo = mode == 1 ? l : (mode == 2 ? 0 : l - 1), hi = mode == 1 ? -1 : (mode == 2 ? l : l + 1);
            if (qs->min_level[q] != lo || qs->max_level[q] != hi) return -8;
        }
        Frame* Cur = W.frame(fs, f, scale, n_levels, nullptr, cam, true);
        if (fs->kp_flags)
            for (int i = 0; i < n; ++i)
                if (fs->kp_flags[k0 + i]) Cur->mvpMapPoints[i] = W.occupant(fs->kp_flags[k0 + i] == 1 ? 1 : 0);
        const std::vector<MapPoint*> before = Cur->mvpMapPoints;

        // source side: one key point + MapPoint per query
        orbgpu_frame_set src;
        std::memset(&src, 0, sizeof(src));
        std::vector<orbgpu_keypoint> skeys(std::max(nq, 1));
        std::vector<uint8_t> sdesc((size_t)std::max(nq, 1) * 32, 0);
        const int32_t soff[2] = {0, nq};
        for (int j = 0; j < nq; ++j) {
            std::memset(&skeys[j], 0, sizeof(orbgpu_keypoint));
            skeys[j].octave = qlevel[j];
            skeys[j].angle = qs->angle ? qs->angle[q0 + j] : 0.f;
        }
        src.n_frames = 1; src.kp_off = soff; src.keys_un = skeys.data(); src.desc = sdesc.data();
        Frame* Last = W.frame(&src, 0, scale, n_levels, nullptr, cam, false);
        std::vector<MapPoint*> qmp(nq, nullptr);
        for (int j = 0; j < nq; ++j) {
            const int q = q0 + j;
            if (!(qs->flags[q] & 1)) continue;   // a query that failed the projection tests: no MapPoint at that key point
            MapPoint* p = W.mappoint(vec3(qs->u[q], qs->v[q], 1.f), qs->desc + (size_t)q * 32, (qs->flags[q] & 4) ? 1 : 0, false);
            if (skip_any) {
                // PredictScale (MapPoint.cc:421-436) must return the query's level: the distance ratio sits mid-level
                const float dist = (float)cv::norm(vec3(qs->u[q], qs->v[q], 1.f));
                p->mfMaxDistance = dist * (float)std::pow((double)scale[1], qlevel[j] - 0.5);
                int nScale = (int)std::ceil(std::log(p->mfMaxDistance / dist) / log_sf);
                nScale = nScale < 0 ? 0 : (nScale >= n_levels ? n_levels - 1 : nScale);
                if (nScale != qlevel[j]) return -5;
            }
            Last->mvpMapPoints[j] = p;
            qmp[j] = p;
        }
        ORBmatcher matcher(0.9f, check_orientation != 0);
        int nm;
        if (!skip_any) {
            // forward / backward follow from tlc = Rlw*twc + tlw against mb (ORBmatcher.cc:1558-1561): twc = 0, so tlc = tlw
            cv::Mat Tlw = eye4();
            Tlw.at<float>(2, 3) = mode == 1 ? 2.f : (mode == 2 ? -2.f : 0.f);
            Last->mTcw = Tlw;
            nm = matcher.SearchByProjection(*Cur, *Last, th, /*bMono=*/false);
        } else {
            KeyFrame* KF = W.keyframe(Last);
            std::set<MapPoint*> none;
            nm = matcher.SearchByProjection(*Cur, KF, none, th, th_dist);
        }
        std::unordered_map<MapPoint*, int> idx = index_of(qmp);
        for (int i = 0; i < n; ++i) {
            MapPoint* p = Cur->mvpMapPoints[i];
            int out = -1;
            if (p) {
                std::unordered_map<MapPoint*, int>::iterator it = idx.find(p);
                if (it != idx.end()) out = it->second;
            } else if (before[i]) {
                out = -2;
            }
            if (kp_match) kp_match[k0 + i] = out;
        }
        if (nmatches_out) nmatches_out[f] = nm;
    }
    return 0;
}

// ORBmatcher::SearchForTriangulation (ORBmatcher.cc:783-975).  The epipole is produced by the reference from the two poses
// (:790-799); key frame 1's camera centre is set to (ex, ey, 1) and key frame 2 gets the identity camera, so the reference
// computes exactly the view's epipole.
void slamref_search_for_triangulation(const orbgpu_frame_set* s1, const orbgpu_frame_set* s2, int n_pairs, const int32_t* idx1v,
                                      const int32_t* idx2v, const float* f12, const float* epipole, const float* scale, const float* sigma2,
                                      int n_levels, int only_stereo, int check_orientation, const int64_t* match_off, int32_t* match12, int32_t*,
                                      int32_t* nmatches_out) {
    for (int p = 0; p < n_pairs; ++p) {
        World W;
        Camera cam = identity_camera(nullptr);
        const int fa = idx1v[p], fb = idx2v[p];
        const int ka = s1->kp_off[fa], na = s1->kp_off[fa + 1] - ka, kb = s2->kp_off[fb], nb = s2->kp_off[fb + 1] - kb;
        Frame* F1 = W.frame(s1, fa, scale, n_levels, sigma2, cam, false);
        Frame* F2 = W.frame(s2, fb, scale, n_levels, sigma2, cam, false);
        if (s1->kp_flags) for (int i = 0; i < na; ++i) if (s1->kp_flags[ka + i] & 1) F1->mvpMapPoints[i] = W.occupant(1);
        if (s2->kp_flags) for (int i = 0; i < nb; ++i) if (s2->kp_flags[kb + i] & 1) F2->mvpMapPoints[i] = W.occupant(1);
        KeyFrame* K1 = W.keyframe(F1);
        KeyFrame* K2 = W.keyframe(F2);
        K1->Ow = vec3(epipole[2 * p], epipole[2 * p + 1], 1.f);
        cv::Mat F(3, 3, CV_32F);
        std::memcpy(F.data, f12 + 9 * p, 36);
        ORBmatcher matcher(0.6f, check_orientation != 0);
        std::vector<std::pair<size_t, size_t> > pairs;
        const int nm = matcher.SearchForTriangulation(K1, K2, F, pairs, only_stereo != 0);
        for (int i = 0; i < na; ++i) match12[match_off[p] + i] = -1;
        for (size_t i = 0; i < pairs.size(); ++i) match12[match_off[p] + (int64_t)pairs[i].first] = (int32_t)pairs[i].second;
        if (nmatches_out) nmatches_out[p] = nm;
    }
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (ORBmatcher.cc:635-768; th_inclusive == 0 && require_mp2 != 0) and
// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (:211-344; th_inclusive != 0 && require_mp2 == 0).
int slamref_search_by_bow(const orbgpu_frame_set* s1, const orbgpu_frame_set* s2, int n_pairs, const int32_t* idx1v, const int32_t* idx2v,
                          float nnratio, int check_orientation, int th_low, int th_inclusive, int require_mp2, const int64_t* match_off,
                          int32_t* match12, int32_t*, int32_t* nmatches_out) {
    if (th_low != ORBmatcher::TH_LOW || (th_inclusive != 0) == (require_mp2 != 0)) return -1;
    const float one_scale[8] = {1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};
    for (int p = 0; p < n_pairs; ++p) {
        World W;
        Camera cam = identity_camera(nullptr);
        const int fa = idx1v[p], fb = idx2v[p];
        const int ka = s1->kp_off[fa], na = s1->kp_off[fa + 1] - ka, kb = s2->kp_off[fb], nb = s2->kp_off[fb + 1] - kb;
        Frame* F1 = W.frame(s1, fa, one_scale, 8, nullptr, cam, false);
        Frame* F2 = W.frame(s2, fb, one_scale, 8, nullptr, cam, false);
        if (s1->kp_flags) for (int i = 0; i < na; ++i) if (s1->kp_flags[ka + i] & 1) F1->mvpMapPoints[i] = W.occupant(1);
        KeyFrame* K1 = W.keyframe(F1);
        ORBmatcher matcher(nnratio, check_orientation != 0);
        for (int i = 0; i < na; ++i) match12[match_off[p] + i] = -1;
        int nm;
        if (require_mp2) {
            if (s2->kp_flags) for (int i = 0; i < nb; ++i) if (s2->kp_flags[kb + i] & 1) F2->mvpMapPoints[i] = W.occupant(1);
            KeyFrame* K2 = W.keyframe(F2);
            std::vector<MapPoint*> v12;
            nm = matcher.SearchByBoW(K1, K2, v12);
            std::unordered_map<MapPoint*, int> idx = index_of(K2->GetMapPointMatches());
            for (int i = 0; i < na; ++i) if (v12[i]) match12[match_off[p] + i] = idx[v12[i]];
        } else {
            std::vector<MapPoint*> vF;
            nm = matcher.SearchByBoW(K1, *F2, vF);
            std::unordered_map<MapPoint*, int> idx = index_of(K1->GetMapPointMatches());
            for (int i2 = 0; i2 < nb; ++i2) if (vF[i2]) match12[match_off[p] + idx[vF[i2]]] = i2;
        }
        if (nmatches_out) nmatches_out[p] = nm;
    }
    return 0;
}

// ORBmatcher::SearchForInitialization (ORBmatcher.cc:493-632): frames2 = the F2 of every pair, queries1 = F1's key points
// (u, v = vbPrevMatched, radius = windowSize — an int in the reference —, flags bit 0 = octave 0).
int slamref_search_for_initialization(const orbgpu_frame_set* fs, const orbgpu_window_query_set* qs, float nnratio, int check_orientation,
                                      int32_t* match12, int32_t* nmatches_out) {
    const float one_scale[8] = {1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};
    for (int f = 0; f < fs->n_frames; ++f) {
        World W;
        const int q0 = qs->q_off[f], nq = qs->q_off[f + 1] - q0;
        Camera cam = identity_camera(fs->grid + 4 * f);
        Frame* F2 = W.frame(fs, f, one_scale, 8, nullptr, cam, true);
        std::vector<orbgpu_keypoint> k1(std::max(nq, 1));
        std::vector<cv::Point2f> prev(nq);
        int window = nq ? (int)qs->radius[q0] : 0;
        for (int j = 0; j < nq; ++j) {
            std::memset(&k1[j], 0, sizeof(orbgpu_keypoint));
            k1[j].octave = (qs->flags[q0 + j] & 1) ? 0 : 1;
            k1[j].angle = qs->angle ? qs->angle[q0 + j] : 0.f;
            prev[j] = cv::Point2f(qs->u[q0 + j], qs->v[q0 + j]);
            if ((float)window != qs->radius[q0 + j] || qs->min_level[q0 + j] != 0 || qs->max_level[q0 + j] != 0) return -1;
        }
        orbgpu_frame_set src;
        std::memset(&src, 0, sizeof(src));
        const int32_t soff[2] = {0, nq};
        src.n_frames = 1; src.kp_off = soff; src.keys_un = k1.data(); src.desc = qs->desc + (size_t)q0 * 32;
        Frame* F1 = W.frame(&src, 0, one_scale, 8, nullptr, cam, false);
        ORBmatcher matcher(nnratio, check_orientation != 0);
        std::vector<int> v12;
        const int nm = matcher.SearchForInitialization(*F1, *F2, prev, v12, window);
        for (int j = 0; j < nq; ++j) match12[q0 + j] = v12[j];
        if (nmatches_out) nmatches_out[f] = nm;
    }
    return 0;
}

// The candidate loop of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (ORBmatcher.cc:977-1137), one query at a
// time on an empty key frame so that the only observable effect — pKF->AddMapPoint(pMP, bestIdx) when bestDist <= TH_LOW — names
// the winner.  Preconditions (asserted): radius[q] == th * scale[l], [min_level, max_level] == [l-1, l], ur[q] == u[q] - mbf.
// q_best_idx[q] = bestIdx if the reference fused (bestDist <= TH_LOW) else -1.
int slamref_fuse_best(const orbgpu_frame_set* fs, const orbgpu_window_query_set* qs, const float* scale, const float* sigma2, int n_levels,
                      float th, float mbf, int32_t* q_best_idx) {
    const float log_sf = std::log(scale[1]);
    for (int f = 0; f < fs->n_frames; ++f) {
        World W;
        const int q0 = qs->q_off[f], nq = qs->q_off[f + 1] - q0;
        Camera cam = identity_camera(fs->grid + 4 * f);
        cam.mbf = mbf;
        Frame* F = W.frame(fs, f, scale, n_levels, sigma2, cam, true);
        KeyFrame* KF = W.keyframe(F);
        ORBmatcher matcher(0.6f, true);
        for (int j = 0; j < nq; ++j) {
            const int q = q0 + j;
            q_best_idx[q] = -1;
            if (!(qs->flags[q] & 1)) continue;
            int lvl = -1;
            for (int l = 0; l < n_levels; ++l) if (qs->radius[q] == th * scale[l]) { lvl = l; break; }
            if (lvl < 0 || qs->min_level[q] != lvl - 1 || qs->max_level[q] != lvl) return -1;
            if (qs->ur && qs->ur[q] != qs->u[q] - mbf) return -4;
            if (qs->u[q] < cam.minX || qs->v[q] < cam.minY) return -3;
            cv::Mat pos = vec3(qs->u[q], qs->v[q], 1.f);
            MapPoint* p = W.mappoint(pos, qs->desc + (size_t)q * 32, 1, false);
            const float dist = (float)cv::norm(pos);
            p->mfMaxDistance = dist * (float)std::pow((double)scale[1], lvl - 0.5);
            p->mfMinDistance = 0.f;
            p->mNormalVector = pos.clone();   // PO.dot(Pn) = dist^2 >= 0.5 dist (dist >= 1)
            int nScale = (int)std::ceil(std::log(p->mfMaxDistance / dist) / log_sf);
            nScale = nScale < 0 ? 0 : (nScale >= n_levels ? n_levels - 1 : nScale);
            if (nScale != lvl) return -5;
            std::vector<MapPoint*> one(1, p);
            const int fused = matcher.Fuse(KF, one, th);
            if (fused) {
                const int idx = p->GetIndexInKeyFrame(KF);
                q_best_idx[q] = idx;
                KF->EraseMapPointMatch((size_t)idx);
            }
        }
    }
    return 0;
}

// Frame::isInFrustum (Frame.cc:274-342) with MapPoint::PredictScale (MapPoint.cc:421-436).  cam[f] as in orbgpu_is_in_frustum;
// min_distance / max_distance are the raw mfMinDistance / mfMaxDistance (the reference applies the 0.8 / 1.2 factors itself).
void slamref_is_in_frustum(int n_frames, const float* cam, float log_scale_factor, int n_levels, float viewing_cos_limit, const int32_t* mp_off,
                           const float* world_pos, const float* normal, const float* min_distance, const float* max_distance, uint8_t* in_view,
                           float* proj_x, float* proj_y, float* proj_xr, int32_t* level, float* view_cos) {
    for (int f = 0; f < n_frames; ++f) {
        World W;
        const float* c = cam + 24 * f;
        Camera cm = {c[15], c[16], c[17], c[18], c[19], 0.f, c[20], c[21], c[22], c[23], 0.f, 0.f};
        std::vector<float> scale(n_levels, 1.f);
        Frame* F = W.frame(nullptr, 0, scale.data(), n_levels, nullptr, cm, false);
        F->mfLogScaleFactor = log_scale_factor;
        F->mRcw = cv::Mat(3, 3, CV_32F); std::memcpy(F->mRcw.data, c, 36);
        F->mtcw = vec3(c[9], c[10], c[11]);
        F->mOw = vec3(c[12], c[13], c[14]);
        for (int q = mp_off[f]; q < mp_off[f + 1]; ++q) {
            MapPoint* p = W.mappoint(vec3(world_pos[3 * q], world_pos[3 * q + 1], world_pos[3 * q + 2]), nullptr, 1, false);
            p->mNormalVector = vec3(normal[3 * q], normal[3 * q + 1], normal[3 * q + 2]);
            p->mfMinDistance = min_distance[q];
            p->mfMaxDistance = max_distance[q];
            p->mTrackProjX = p->mTrackProjY = p->mTrackProjXR = p->mTrackViewCos = 0.f;
            p->mnTrackScaleLevel = 0;
            const bool ok = F->isInFrustum(p, viewing_cos_limit);
            in_view[q] = ok ? 1 : 0;
            proj_x[q] = ok ? p->mTrackProjX : 0.f; proj_y[q] = ok ? p->mTrackProjY : 0.f; proj_xr[q] = ok ? p->mTrackProjXR : 0.f;
            level[q] = ok ? p->mnTrackScaleLevel : 0; view_cos[q] = ok ? p->mTrackViewCos : 0.f;
        }
    }
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-316).  The reference iterates std::map<KeyFrame*, size_t>, i.e. in
// key-frame ADDRESS order: observation slot j of every point is held by key frame j, and the key frames are placed at ascending
// addresses, so the iteration order is the view's row order.  best_desc[p] = the descriptor the reference selected (32 B).
void slamref_distinctive_descriptors(int n_points, const int32_t* obs_off, const uint8_t* desc, uint8_t* best_desc, uint8_t* has_desc) {
    int max_obs = 0;
    for (int p = 0; p < n_points; ++p) max_obs = std::max(max_obs, obs_off[p + 1] - obs_off[p]);
    World W;
    Camera cam = identity_camera(nullptr);
    const float one = 1.f;
    // slot key frames: key frame j holds row p = the j-th observed descriptor of point p
    std::vector<KeyFrame*> slot(max_obs, nullptr);
    void* block = ::operator new(sizeof(KeyFrame) * (size_t)std::max(max_obs, 1));
    std::vector<uint8_t> rows((size_t)std::max(n_points, 1) * 32);
    std::vector<orbgpu_keypoint> keys(std::max(n_points, 1));
    std::memset(keys.data(), 0, keys.size() * sizeof(orbgpu_keypoint));
    const int32_t off[2] = {0, n_points};
    for (int j = 0; j < max_obs; ++j) {
        std::fill(rows.begin(), rows.end(), 0);
        for (int p = 0; p < n_points; ++p)
            if (j < obs_off[p + 1] - obs_off[p]) std::memcpy(&rows[(size_t)p * 32], desc + (size_t)(obs_off[p] + j) * 32, 32);
        orbgpu_frame_set s;
        std::memset(&s, 0, sizeof(s));
        s.n_frames = 1; s.kp_off = off; s.keys_un = keys.data(); s.desc = rows.data();
        Frame* F = W.frame(&s, 0, &one, 1, nullptr, cam, false);
        slot[j] = new ((char*)block + sizeof(KeyFrame) * (size_t)j) KeyFrame(*F, &W.map, nullptr);
    }
    for (int p = 0; p < n_points; ++p) {
        MapPoint* mp = W.mappoint(vec3(0, 0, 0), nullptr, 0, false);
        mp->mDescriptor = cv::Mat();
        for (int j = 0; j < obs_off[p + 1] - obs_off[p]; ++j) mp->AddObservation(slot[j], (size_t)p);
        mp->ComputeDistinctiveDescriptors();
        const cv::Mat d = mp->GetDescriptor();
        has_desc[p] = d.empty() ? 0 : 1;
        if (!d.empty()) std::memcpy(best_desc + (size_t)p * 32, d.data, 32); else std::memset(best_desc + (size_t)p * 32, 0, 32);
    }
    for (int j = 0; j < max_obs; ++j) slot[j]->~KeyFrame();
    ::operator delete(block);
}

// Frame::Frame(imLeft, imRight, ...) (Frame.cc:61-117): the reference's stereo constructor — both extractions on two threads,
// ComputeStereoMatches (Frame.cc:501-675), AssignFeaturesToGrid.  Outputs: the left key points / descriptors and mvuRight / mvDepth.
//
// Reference quirk: the constructor calls ComputeStereoMatches() (Frame.cc:88) BEFORE it sets mb = mbf / fx (:113), and
// ComputeStereoMatches reads mb as minZ (:525) — so the reference computes maxD = mbf / <whatever the Frame's storage held>.
// The value is indeterminate upstream (a stack temporary in Tracking::GrabImageStereo).  To get a defined result the Frame is
// constructed in storage whose mb slot already holds `mb_at_entry`; the constructor never writes that member before the call.
// The C ABI / port take mb as an explicit argument (orbgpu_stereo_matches), so parity is checked for the intended mbf / fx
// and for other values alike.
int slamref_stereo_frame(const uint8_t* left, const uint8_t* right, int w, int h, int nfeatures, float scale_factor, int nlevels, int ini_th,
                         int min_th, float fx, float fy, float cx, float cy, float bf, float mb_at_entry, orbgpu_keypoint* kp_out,
                         uint8_t* desc_out, float* u_right, float* depth, int capacity) {
    ORBextractor exL(nfeatures, scale_factor, nlevels, ini_th, min_th), exR(nfeatures, scale_factor, nlevels, ini_th, min_th);
    cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)w), imR(h, w, CV_8UC1, (void*)right, (size_t)w);
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
    cv::Mat dist(4, 1, CV_32F);
    for (int i = 0; i < 4; ++i) dist.at<float>(i) = 0.f;
    Frame::mbInitialComputations = true;
    void* storage = ::operator new(sizeof(Frame));
    std::memset(storage, 0, sizeof(Frame));
    std::memcpy((char*)storage + offsetof(Frame, mb), &mb_at_entry, sizeof(float));
    Frame* F = new (storage) Frame(imL, imR, 0.0, &exL, &exR, nullptr, K, dist, bf, 40.f);
    const int N = F->N;
    const int n = N < capacity ? N : capacity;
    if (n) {
        std::memcpy(kp_out, F->mvKeys.data(), (size_t)n * sizeof(orbgpu_keypoint));
        for (int i = 0; i < n; ++i) std::memcpy(desc_out + (size_t)i * 32, F->mDescriptors.ptr(i), 32);
        std::memcpy(u_right, F->mvuRight.data(), (size_t)n * 4);
        std::memcpy(depth, F->mvDepth.data(), (size_t)n * 4);
    }
    F->~Frame();
    ::operator delete(storage);
    return N;
}

// Frame::GetFeaturesInArea (Frame.cc:353-410) over the grid built by Frame::AssignFeaturesToGrid (:232-247): indices for one query.
int slamref_features_in_area(const orbgpu_frame_set* fs, int f, float x, float y, float r, int min_level, int max_level, int32_t* out, int capacity) {
    World W;
    const float one_scale[8] = {1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};
    Frame* F = W.frame(fs, f, one_scale, 8, nullptr, identity_camera(fs->grid + 4 * f), true);
    const std::vector<size_t> v = F->GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < capacity; ++i) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// CPU baseline of the matching leg (bench.py): the reference's own ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) on n_pairs
// key-frame pairs, one pair at a time per std::thread.  The KeyFrame / MapPoint objects are built first (untimed: in ORB-SLAM2 they
// exist already); the returned seconds cover the search calls only.  *matches receives the sum of the return values.
double slamref_bench_bow(const orbgpu_frame_set* s1, const orbgpu_frame_set* s2, int n_pairs, const int32_t* idx1v, const int32_t* idx2v,
                         float nnratio, int check_orientation, int nthreads, long long* matches) {
    const float one_scale[8] = {1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};
    std::vector<std::unique_ptr<World> > worlds;
    std::vector<std::pair<KeyFrame*, KeyFrame*> > kfs;
    for (int p = 0; p < n_pairs; ++p) {
        worlds.emplace_back(new World());
        World& W = *worlds.back();
        Camera cam = identity_camera(nullptr);
        const int fa = idx1v[p], fb = idx2v[p];
        const int ka = s1->kp_off[fa], na = s1->kp_off[fa + 1] - ka, kb = s2->kp_off[fb], nb = s2->kp_off[fb + 1] - kb;
        Frame* F1 = W.frame(s1, fa, one_scale, 8, nullptr, cam, false);
        Frame* F2 = W.frame(s2, fb, one_scale, 8, nullptr, cam, false);
        if (s1->kp_flags) for (int i = 0; i < na; ++i) if (s1->kp_flags[ka + i] & 1) F1->mvpMapPoints[i] = W.occupant(1);
        if (s2->kp_flags) for (int i = 0; i < nb; ++i) if (s2->kp_flags[kb + i] & 1) F2->mvpMapPoints[i] = W.occupant(1);
        kfs.push_back(std::make_pair(W.keyframe(F1), W.keyframe(F2)));
    }
    std::vector<long long> found(nthreads, 0);
    std::vector<std::thread> th;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            ORBmatcher matcher(nnratio, check_orientation != 0);
            std::vector<MapPoint*> v12;
            for (int p = t; p < n_pairs; p += nthreads) found[t] += matcher.SearchByBoW(kfs[p].first, kfs[p].second, v12);
        });
    for (std::thread& x : th) x.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long tot = 0;
    for (long long v : found) tot += v;
    if (matches) *matches = tot;
    return sec;
}

}  // extern "C"

// ref_twin_test — TEST INFRASTRUCTURE, runs on the GPU box (prebuilt into oracle/_ref/ by oracle/Makefile).
//
// The drop-in ORBmatcher shell (orb_slam2_with_comment_b200/csrc/host/ORBmatcher_gpu.cc, all eleven search members on the GPU)
// compiled against the REFERENCE'S OWN headers and linked with the reference's own Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc,
// KeyFrameDatabase.cc, ORBextractor.cc and DBoW2 — exactly the deployment form — against the reference's own ORBmatcher.cc, which
// is compiled into the same binary under another class name (oracle/ref_matcher_renamed.cc: ORBmatcherRef).  Every scenario
// builds two identical worlds of real Frame / KeyFrame / MapPoint / Map objects (poses, 3-D points, observations; normals,
// depth ranges and distinctive descriptors computed by the reference's own MapPoint code), calls the shell member on one and
// the reference member on the other, and compares the complete state of both worlds afterwards: return values, every
// mvpMapPoints slot, every MapPoint's bad flag / observations / replacement, every output vector.
#include "slam_world.h"

#undef ORBMATCHER_H
#define ORBmatcher ORBmatcherRef
#include "ORBmatcher.h"   // the same declaration once more, as ORBmatcherRef
#undef ORBmatcher

#include <cstdio>
#include <functional>
#include <numeric>
#include <type_traits>

using namespace slamworld;

namespace {

int fails = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { ++fails; printf("FAIL: "); printf(__VA_ARGS__); printf("\n"); } } while (0)

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }
    double uni() { return (next() & 0xFFFFFF) / 16777216.0; }
    double uni(double a, double b) { return a + (b - a) * uni(); }
    double gauss() { double u = 0; for (int i = 0; i < 6; ++i) u += uni(); return (u - 3.0) * 1.4142; }
    int below(int n) { return (int)(next() % (uint32_t)n); }
};

const float kFx = 517.3f, kFy = 516.5f, kCx = 318.6f, kCy = 255.3f, kMbf = 40.f;
const int kW = 640, kH = 480, kLevels = 8;
std::vector<float> g_sf;

Camera tum_camera() {
    Camera c = {kFx, kFy, kCx, kCy, kMbf, kMbf / kFx, 0.f, (float)kW, 0.f, (float)kH, 64.f / kW, 48.f / kH};
    return c;
}

struct Pose { double R[9], t[3]; };
Pose make_pose(Rng& r, double rot, double trans) {
    const double a = r.gauss() * rot, b = r.gauss() * rot, c = r.gauss() * rot;
    const double ca = cos(a), sa = sin(a), cb = cos(b), sb = sin(b), cc = cos(c), sc = sin(c);
    Pose P = {{cb * cc, -cb * sc, sb, sa * sb * cc + ca * sc, -sa * sb * sc + ca * cc, -sa * cb, -ca * sb * cc + sa * sc, ca * sb * sc + sa * cc, ca * cb},
              {r.gauss() * trans, r.gauss() * trans, r.gauss() * trans}};
    return P;
}
cv::Mat pose_mat(const Pose& P) {
    cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T.at<float>(i, j) = (float)P.R[3 * i + j];
        T.at<float>(i, 3) = (float)P.t[i];
    }
    return T;
}

struct WorldPoint { float X[3]; uint8_t desc[32]; float angle; double size; };

struct View {   // flat arrays of one camera view, and which world point each key point shows (-1: clutter)
    std::vector<orbgpu_keypoint> keys;
    std::vector<uint8_t> desc;
    std::vector<float> ur;
    std::vector<int> point;
    std::vector<int32_t> node_id, feat_off, feat;
};

View make_view(Rng& r, const std::vector<WorldPoint>& pts, const Pose& P, double stereo_frac, int clutter, double flip) {
    View V;
    struct K { orbgpu_keypoint k; uint8_t d[32]; float ur; int pt; };
    std::vector<K> ks;
    for (size_t i = 0; i < pts.size(); ++i) {
        double pc[3];
        for (int a = 0; a < 3; ++a) pc[a] = P.R[3 * a] * pts[i].X[0] + P.R[3 * a + 1] * pts[i].X[1] + P.R[3 * a + 2] * pts[i].X[2] + P.t[a];
        if (pc[2] < 0.3 || r.uni() < 0.15) continue;
        const double u = kFx * pc[0] / pc[2] + kCx + r.gauss() * 0.7, v = kFy * pc[1] / pc[2] + kCy + r.gauss() * 0.7;
        if (u < 20 || u > kW - 20 || v < 20 || v > kH - 20) continue;
        const double dist = sqrt(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
        int lvl = (int)ceil(log(pts[i].size / dist) / log(1.2) - 0.5);
        lvl = lvl < 0 ? 0 : (lvl > kLevels - 1 ? kLevels - 1 : lvl);
        K k;
        memset(&k.k, 0, sizeof(k.k));
        k.k.x = (float)u; k.k.y = (float)v; k.k.octave = lvl; k.k.size = 31.f * g_sf[lvl]; k.k.class_id = -1;
        k.k.angle = (float)fmod(pts[i].angle + r.gauss() * 5.0 + 720.0, 360.0);
        memcpy(k.d, pts[i].desc, 32);
        for (int b = 0; b < 256; ++b) if (r.uni() < flip) k.d[b >> 3] ^= (uint8_t)(1u << (b & 7));
        k.ur = r.uni() < stereo_frac ? (float)(u - kMbf / pc[2]) : -1.f;
        k.pt = (int)i;
        ks.push_back(k);
    }
    for (int c = 0; c < clutter; ++c) {
        K k;
        memset(&k.k, 0, sizeof(k.k));
        k.k.x = (float)r.uni(20, kW - 20); k.k.y = (float)r.uni(20, kH - 20); k.k.octave = r.below(kLevels);
        k.k.size = 31.f * g_sf[k.k.octave]; k.k.angle = (float)r.uni(0, 360); k.k.class_id = -1;
        for (int b = 0; b < 32; ++b) k.d[b] = (uint8_t)r.next();
        k.ur = -1.f; k.pt = -1;
        ks.push_back(k);
    }
    for (size_t i = ks.size(); i > 1; --i) std::swap(ks[i - 1], ks[r.below((int)i)]);
    std::map<int, std::vector<int> > fv;
    for (size_t i = 0; i < ks.size(); ++i) {
        V.keys.push_back(ks[i].k);
        V.desc.insert(V.desc.end(), ks[i].d, ks[i].d + 32);
        V.ur.push_back(ks[i].ur);
        V.point.push_back(ks[i].pt);
        fv[ks[i].pt >= 0 ? (ks[i].pt * 7) % 61 : r.below(61)].push_back((int)i);
    }
    V.feat_off.push_back(0);
    for (std::map<int, std::vector<int> >::iterator it = fv.begin(); it != fv.end(); ++it) {
        V.node_id.push_back(it->first);
        V.feat.insert(V.feat.end(), it->second.begin(), it->second.end());
        V.feat_off.push_back((int32_t)V.feat.size());
    }
    return V;
}

// One complete world; built twice per scenario from the same seed.
struct Scene {
    World W;
    std::vector<WorldPoint> pts;
    std::vector<Pose> poses;
    std::vector<View> views;
    std::vector<Frame*> frames;      // one per view
    std::vector<KeyFrame*> kfs;      // key frame of view v (or NULL)
    std::vector<std::vector<MapPoint*> > maps;   // maps[m][point id]: MapPoint of world point in map m (or NULL)

    // view 1's pose gets (base_x, 0, fwd_z) added to its translation BEFORE its key points are generated
    Scene(uint64_t seed, int n_points, int n_views, double stereo_frac, int clutter, double base_x = 0.0, double fwd_z = 0.0) {
        Rng r(seed);
        for (int i = 0; i < n_points; ++i) {
            WorldPoint p;
            p.X[0] = (float)r.uni(-4, 4); p.X[1] = (float)r.uni(-3, 3); p.X[2] = (float)r.uni(2, 12);
            for (int b = 0; b < 32; ++b) p.desc[b] = (uint8_t)r.next();
            if (i % 9 == 0 && i) memcpy(p.desc, pts[i - 1].desc, 32), p.desc[3] ^= 0x11;   // near-duplicate descriptors: ties and ratio tests
            p.angle = (float)r.uni(0, 360);
            p.size = p.X[2] * pow(1.2, r.uni(0.5, 6.5));
            pts.push_back(p);
        }
        for (int v = 0; v < n_views; ++v) {
            poses.push_back(make_pose(r, v ? 0.03 : 0.0, v ? 0.25 : 0.0));
            if (v == 1) { poses.back().t[0] += base_x; poses.back().t[2] += fwd_z; }
            views.push_back(make_view(r, pts, poses.back(), stereo_frac, clutter, 0.04));
            const View& V = views.back();
            orbgpu_frame_set fs;
            memset(&fs, 0, sizeof(fs));
            const int32_t off[2] = {0, (int32_t)V.keys.size()}, noff[2] = {0, (int32_t)V.node_id.size()};
            fs.n_frames = 1; fs.kp_off = off; fs.keys_un = V.keys.data(); fs.desc = V.desc.data(); fs.u_right = V.ur.data();
            fs.fv_node_off = noff; fs.fv_node_id = V.node_id.data(); fs.fv_feat_off = V.feat_off.data(); fs.fv_feat = V.feat.data();
            Frame* F = W.frame(&fs, 0, g_sf.data(), kLevels, nullptr, tum_camera(), true);
            F->SetPose(pose_mat(poses.back()));
            frames.push_back(F);
            kfs.push_back(nullptr);
        }
    }
    KeyFrame* keyframe(int v) {
        if (!kfs[v]) kfs[v] = W.keyframe(frames[v]);
        return kfs[v];
    }
    // A map: one MapPoint per world point seen by at least one of `views_of_map` (key frames), observed by all of them that see it
    // with probability `keep`; normals, depth ranges and descriptors by the reference's own MapPoint members.
    int make_map(Rng& r, const std::vector<int>& views_of_map, double keep) {
        std::vector<MapPoint*> M(pts.size(), nullptr);
        for (size_t p = 0; p < pts.size(); ++p) {
            MapPoint* mp = nullptr;
            for (size_t a = 0; a < views_of_map.size(); ++a) {
                const int v = views_of_map[a];
                const View& V = views[v];
                for (size_t i = 0; i < V.point.size(); ++i) {
                    if (V.point[i] != (int)p || r.uni() > keep) continue;
                    KeyFrame* K = keyframe(v);
                    if (K->GetMapPoint(i)) continue;
                    if (!mp) {
                        mp = new MapPoint(vec3(pts[p].X[0], pts[p].X[1], pts[p].X[2]), K, &W.map);
                        W.mps.push_back(mp);
                    }
                    mp->AddObservation(K, i);
                    K->AddMapPoint(mp, i);
                }
            }
            if (mp) {
                mp->ComputeDistinctiveDescriptors();
                mp->UpdateNormalAndDepth();
            }
            M[p] = mp;
        }
        maps.push_back(M);
        return (int)maps.size() - 1;
    }
    // complete observable state
    std::vector<long long> dump() {
        std::unordered_map<MapPoint*, int> id;
        for (size_t i = 0; i < W.mps.size(); ++i) id[W.mps[i]] = (int)i;
        std::unordered_map<KeyFrame*, int> kid;
        for (size_t i = 0; i < W.kfs.size(); ++i) kid[W.kfs[i]] = (int)i;
        std::vector<long long> d;
        auto pid = [&](MapPoint* p) -> long long { return p ? (id.count(p) ? id[p] : -7) : -1; };
        for (Frame* F : frames) for (MapPoint* p : F->mvpMapPoints) d.push_back(pid(p));
        for (KeyFrame* K : W.kfs) for (MapPoint* p : K->GetMapPointMatches()) d.push_back(pid(p));
        for (MapPoint* p : W.mps) {
            d.push_back(p->isBad()); d.push_back(p->Observations()); d.push_back(pid(p->GetReplaced()));
            std::vector<std::pair<int, long long> > obs;
            const std::map<KeyFrame*, size_t> o = p->GetObservations();
            for (std::map<KeyFrame*, size_t>::const_iterator it = o.begin(); it != o.end(); ++it) obs.push_back(std::make_pair(kid[it->first], (long long)it->second));
            std::sort(obs.begin(), obs.end());
            for (size_t i = 0; i < obs.size(); ++i) { d.push_back(obs[i].first); d.push_back(obs[i].second); }
            d.push_back(-99);
        }
        return d;
    }
    long long pid(MapPoint* p) {
        if (!p) return -1;
        for (size_t i = 0; i < W.mps.size(); ++i) if (W.mps[i] == p) return (long long)i;
        return -7;
    }
};

// Runs `body` on two identical scenes, once with the shell (GPU) and once with the reference class, and compares everything.
template <class Build, class Body>
void twin(const char* name, Build build, Body body) {
    std::unique_ptr<Scene> A(build()), B(build());
    EXPECT(A->dump() == B->dump(), "%s: the two worlds differ before the call", name);
    std::vector<long long> outA, outB;
    body(*A, (ORBmatcher*)nullptr, outA);      // the shell: searches on the GPU
    body(*B, (ORBmatcherRef*)nullptr, outB);   // the reference's ORBmatcher.cc
    const std::vector<long long> dA = A->dump(), dB = B->dump();
    size_t diff = 0;
    for (size_t i = 0; i < dA.size() && i < dB.size(); ++i) diff += dA[i] != dB[i];
    EXPECT(outA == outB, "%s: outputs differ (return value shell %lld, reference %lld)", name, outA.empty() ? -1 : outA[0], outB.empty() ? -1 : outB[0]);
    EXPECT(dA.size() == dB.size() && diff == 0, "%s: %zu world-state entries differ after the call", name, diff);
    EXPECT(!outB.empty() && outB[0] > 0, "%s: degenerate scenario (the reference found nothing)", name);
    printf("%-58s reference returns %4lld, shell %4lld, outputs %s, world state %s\n", name, outB.empty() ? -1 : outB[0], outA.empty() ? -1 : outA[0],
           outA == outB ? "equal" : "DIFFER", (dA == dB) ? "equal" : "DIFFERS");
}

std::vector<MapPoint*> all_points(Scene& S, int m) {
    std::vector<MapPoint*> v;
    for (MapPoint* p : S.maps[m]) if (p) v.push_back(p);
    return v;
}

cv::Mat sim3_of(const Pose& P, double s) {
    cv::Mat T = pose_mat(P);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) T.at<float>(i, j) = (float)(s * T.at<float>(i, j));
    return T;
}

}  // namespace

int main() {
    int ndev = 0;
    if (orbgpu_device_count(&ndev) != 0 || ndev == 0) {
        printf("ref_twin_test: no CUDA device — the shell has no CPU fallback\n");
        return 3;
    }
    g_sf.assign(kLevels, 1.f);
    for (int i = 1; i < kLevels; ++i) g_sf[i] = (float)(g_sf[i - 1] * (double)1.2f);

    // 1. SearchByProjection(Frame&, const vector<MapPoint*>&, th) — Tracking::SearchLocalPoints
    for (float th : {1.f, 3.f})
        twin("SearchByProjection(Frame, vpMapPoints, th)", [] {
            Scene* S = new Scene(11, 1400, 2, 0.3, 300);
            Rng r(5);
            S->make_map(r, {0}, 0.9);
            Frame* F = S->frames[1];
            for (size_t i = 0; i < F->mvpMapPoints.size(); ++i)   // a share of the frame's key points already tracked
                if (S->views[1].point[i] >= 0 && r.uni() < 0.2) F->mvpMapPoints[i] = S->maps[0][S->views[1].point[i]];
            for (MapPoint* p : all_points(*S, 0)) { p->mbTrackInView = false; F->isInFrustum(p, 0.5f); }
            return S; },
             [th](Scene& S, auto* tag, std::vector<long long>& out) {
                 typename std::remove_pointer<decltype(tag)>::type m(0.8f, true);
                 out.push_back(m.SearchByProjection(*S.frames[1], all_points(S, 0), th));
             });

    // 2. SearchByProjection(Frame& Current, const Frame& Last, th, bMono) — Tracking::TrackWithMotionModel
    for (int mono = 0; mono < 2; ++mono)
        twin(mono ? "SearchByProjection(Current, Last, th, bMono=true)" : "SearchByProjection(Current, Last, th, bMono=false)", [] {
            Scene* S = new Scene(12, 1200, 2, 0.4, 250, 0.0, -0.9);   // forward motion beyond the baseline: the forward level range of :1604
            Rng r(6);
            S->make_map(r, {0}, 0.85);
            Frame* L = S->frames[0];
            L->mvpMapPoints = S->keyframe(0)->GetMapPointMatches();
            for (size_t i = 0; i < L->mvbOutlier.size(); ++i) L->mvbOutlier[i] = r.uni() < 0.05;
            Frame* C = S->frames[1];
            for (size_t i = 0; i < C->mvpMapPoints.size(); ++i)
                if (S->views[1].point[i] >= 0 && r.uni() < 0.1) C->mvpMapPoints[i] = S->maps[0][S->views[1].point[i]];
            return S; },
             [mono](Scene& S, auto* tag, std::vector<long long>& out) {
                 typename std::remove_pointer<decltype(tag)>::type m(0.9f, true);
                 out.push_back(m.SearchByProjection(*S.frames[1], *S.frames[0], 15.f, mono != 0));
             });

    // 3. SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) — Tracking::Relocalization
    twin("SearchByProjection(Current, KeyFrame, sAlreadyFound, th, d)", [] {
        Scene* S = new Scene(13, 1200, 2, 0.0, 250);
        Rng r(7);
        S->make_map(r, {0}, 0.9);
        Frame* C = S->frames[1];
        for (size_t i = 0; i < C->mvpMapPoints.size(); ++i)
            if (S->views[1].point[i] >= 0 && r.uni() < 0.15) C->mvpMapPoints[i] = S->maps[0][S->views[1].point[i]];
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.9f, true);
             std::set<MapPoint*> found;
             for (MapPoint* p : S.frames[1]->mvpMapPoints) if (p) found.insert(p);
             out.push_back(m.SearchByProjection(*S.frames[1], S.keyframe(0), found, 10.f, 100));
         });

    // 4. SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) — LoopClosing::ComputeSim3
    twin("SearchByProjection(KeyFrame, Scw, vpPoints, vpMatched, th)", [] {
        Scene* S = new Scene(14, 1300, 2, 0.0, 250);
        Rng r(8);
        S->make_map(r, {0}, 0.9);
        S->keyframe(1);
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.75f, true);
             KeyFrame* K = S.keyframe(1);
             Rng r(3);
             std::vector<MapPoint*> matched(K->N, (MapPoint*)nullptr);
             for (int i = 0; i < K->N; ++i)
                 if (S.views[1].point[i] >= 0 && r.uni() < 0.2) matched[i] = S.maps[0][S.views[1].point[i]];
             out.push_back(m.SearchByProjection(K, sim3_of(S.poses[1], 1.07), all_points(S, 0), matched, 10));
             for (MapPoint* p : matched) out.push_back(S.pid(p));
         });

    // 5. SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) — Tracking::TrackReferenceKeyFrame / Relocalization
    twin("SearchByBoW(KeyFrame, Frame, vpMapPointMatches)", [] {
        Scene* S = new Scene(15, 1500, 2, 0.0, 300);
        Rng r(9);
        S->make_map(r, {0}, 0.9);
        for (MapPoint* p : all_points(*S, 0)) if (r.uni() < 0.05) p->mbBad = true;
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.7f, true);
             std::vector<MapPoint*> v;
             out.push_back(m.SearchByBoW(S.keyframe(0), *S.frames[1], v));
             for (MapPoint* p : v) out.push_back(S.pid(p));
         });

    // 6. SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) — LoopClosing::ComputeSim3
    twin("SearchByBoW(KeyFrame, KeyFrame, vpMatches12)", [] {
        Scene* S = new Scene(16, 1500, 2, 0.0, 300);
        Rng r(10);
        S->make_map(r, {0}, 0.9);
        S->make_map(r, {1}, 0.9);
        for (MapPoint* p : all_points(*S, 1)) if (r.uni() < 0.05) p->mbBad = true;
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.75f, true);
             std::vector<MapPoint*> v;
             out.push_back(m.SearchByBoW(S.keyframe(0), S.keyframe(1), v));
             for (MapPoint* p : v) out.push_back(S.pid(p));
         });

    // 7. SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — Tracking::MonocularInitialization
    twin("SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, 100)", [] {
        Scene* S = new Scene(17, 1600, 2, 0.0, 400);
        for (int v = 0; v < 2; ++v)   // the bootstrap only uses level-0 key points: make them the majority
            for (size_t i = 0; i < S->frames[v]->mvKeysUn.size(); ++i)
                if (i % 3) { S->frames[v]->mvKeysUn[i].octave = 0; S->frames[v]->mvKeys[i].octave = 0; }
        for (int v = 0; v < 2; ++v) {
            for (int i = 0; i < FRAME_GRID_COLS; ++i) for (int j = 0; j < FRAME_GRID_ROWS; ++j) S->frames[v]->mGrid[i][j].clear();
            S->frames[v]->AssignFeaturesToGrid();
        }
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.9f, true);
             std::vector<cv::Point2f> prev;
             for (const cv::KeyPoint& k : S.frames[0]->mvKeysUn) prev.push_back(k.pt);
             std::vector<int> m12;
             out.push_back(m.SearchForInitialization(*S.frames[0], *S.frames[1], prev, m12, 100));
             for (int v : m12) out.push_back(v);
             for (const cv::Point2f& p : prev) { out.push_back((long long)(p.x * 64)); out.push_back((long long)(p.y * 64)); }
         });

    // 8. SearchForTriangulation(KF1, KF2, F12, vMatchedPairs, bOnlyStereo) — LocalMapping::CreateNewMapPoints
    for (int only = 0; only < 2; ++only)
        twin(only ? "SearchForTriangulation(KF1, KF2, F12, pairs, bOnlyStereo=true)" : "SearchForTriangulation(KF1, KF2, F12, pairs, bOnlyStereo=false)", [] {
            Scene* S = new Scene(18, 1500, 2, 0.5, 300, 0.6);   // a real baseline
            Rng r(11);
            S->make_map(r, {0, 1}, 0.3);   // most key points still lack a MapPoint: those are the ones to triangulate
            return S; },
             [only](Scene& S, auto* tag, std::vector<long long>& out) {
                 typename std::remove_pointer<decltype(tag)>::type m(0.6f, false);
                 KeyFrame *K1 = S.keyframe(0), *K2 = S.keyframe(1);
                 // LocalMapping::ComputeF12 (LocalMapping.cc:669-687): F12 = K1^-T * [t12]x * R12 * K2^-1
                 cv::Mat R1w = K1->GetRotation(), t1w = K1->GetTranslation(), R2w = K2->GetRotation(), t2w = K2->GetTranslation();
                 cv::Mat R12 = R1w * R2w.t();
                 cv::Mat t12 = -R1w * R2w.t() * t2w + t1w;
                 cv::Mat tx = cv::Mat(3, 3, CV_32F);
                 const float x = t12.at<float>(0), y = t12.at<float>(1), z = t12.at<float>(2);
                 const float sk[9] = {0, -z, y, z, 0, -x, -y, x, 0};
                 memcpy(tx.data, sk, 36);
                 cv::Mat Kinv = cv::Mat::eye(3, 3, CV_32F);
                 Kinv.at<float>(0, 0) = 1.f / kFx; Kinv.at<float>(1, 1) = 1.f / kFy; Kinv.at<float>(0, 2) = -kCx / kFx; Kinv.at<float>(1, 2) = -kCy / kFy;
                 cv::Mat F12 = Kinv.t() * tx * R12 * Kinv;
                 std::vector<std::pair<size_t, size_t> > pairs;
                 out.push_back(m.SearchForTriangulation(K1, K2, F12, pairs, only != 0));
                 for (size_t i = 0; i < pairs.size(); ++i) { out.push_back((long long)pairs[i].first); out.push_back((long long)pairs[i].second); }
             });

    // 9. Fuse(KeyFrame*, const vector<MapPoint*>&, th) — LocalMapping::SearchInNeighbors
    twin("Fuse(KeyFrame, vpMapPoints, th)", [] {
        Scene* S = new Scene(19, 1300, 2, 0.4, 250);
        Rng r(12);
        S->make_map(r, {0}, 0.5);   // the target key frame's own points (duplicates of the neighbour's: Replace fires)
        S->make_map(r, {1}, 0.9);
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.6f, true);
             out.push_back(m.Fuse(S.keyframe(0), all_points(S, 1), 3.0f));
         });

    // 10. Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) — LoopClosing::SearchAndFuse
    twin("Fuse(KeyFrame, Scw, vpPoints, th, vpReplacePoint)", [] {
        Scene* S = new Scene(20, 1300, 2, 0.0, 250);
        Rng r(13);
        S->make_map(r, {0}, 0.5);
        S->make_map(r, {1}, 0.9);
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.8f, true);
             const std::vector<MapPoint*> pts = all_points(S, 1);
             std::vector<MapPoint*> repl(pts.size(), (MapPoint*)nullptr);
             out.push_back(m.Fuse(S.keyframe(0), sim3_of(S.poses[0], 0.96), pts, 4.f, repl));
             for (MapPoint* p : repl) out.push_back(S.pid(p));
         });

    // 11. SearchBySim3(KF1, KF2, vpMatches12, s12, R12, t12, th) — LoopClosing::ComputeSim3
    twin("SearchBySim3(KF1, KF2, vpMatches12, s12, R12, t12, th)", [] {
        Scene* S = new Scene(21, 1400, 2, 0.0, 250);
        Rng r(14);
        S->make_map(r, {0}, 0.9);
        S->make_map(r, {1}, 0.9);
        return S; },
         [](Scene& S, auto* tag, std::vector<long long>& out) {
             typename std::remove_pointer<decltype(tag)>::type m(0.75f, true);
             KeyFrame *K1 = S.keyframe(0), *K2 = S.keyframe(1);
             cv::Mat R12 = K1->GetRotation() * K2->GetRotation().t();
             cv::Mat t12 = -R12 * K2->GetTranslation() + K1->GetTranslation();
             Rng r(4);
             std::vector<MapPoint*> v12(K1->N, (MapPoint*)nullptr);
             for (int i = 0; i < K1->N; ++i)   // matches known from SearchByBoW
                 if (S.views[0].point[i] >= 0 && r.uni() < 0.25 && S.maps[1][S.views[0].point[i]] && K1->GetMapPoint(i)) v12[i] = S.maps[1][S.views[0].point[i]];
             const float s12 = 1.0f;
             out.push_back(m.SearchBySim3(K1, K2, v12, s12, R12, t12, 7.5f));
             for (MapPoint* p : v12) out.push_back(S.pid(p));
         });

    printf(fails ? "ref_twin_test: %d FAILURES\n" : "ref_twin_test: every shell member leaves its world exactly as the reference's ORBmatcher.cc does (%d failures)\n", fails);
    return fails ? 1 : 0;
}

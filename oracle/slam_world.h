// slam_world.h — TEST INFRASTRUCTURE.  Builds REAL reference objects (ORB_SLAM2::Frame / KeyFrame / MapPoint / Map from
// /root/reference/include, compiled from /root/reference/src where they lie) out of the flat views of include/orbgpu.h.
// Shared by oracle/slam_ref.cc (the C wrapper behind libslamref.so) and tests/cpp/ref_twin_test.cc (shell vs reference on twin
// worlds).  private / protected are opened for the including translation unit only: the fields set here are filled in the
// reference by code outside the hot path (constructors that run a full extraction, UpdateNormalAndDepth, ...).
#ifndef ORBGPU_ORACLE_SLAM_WORLD_H
#define ORBGPU_ORACLE_SLAM_WORLD_H
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#undef private
#undef protected

#include "orbgpu.h"

using namespace ORB_SLAM2;

namespace slamworld {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbgpu_keypoint), "cv::KeyPoint layout");

struct Camera {
    float fx, fy, cx, cy, mbf, mb, minX, maxX, minY, maxY, gridWInv, gridHInv;
};
inline Camera identity_camera(const float* grid4) {
    Camera c = {1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (grid4) {
        c.minX = grid4[0]; c.minY = grid4[1]; c.gridWInv = grid4[2]; c.gridHInv = grid4[3];
        c.maxX = 1e9f; c.maxY = 1e9f;   // the upper image bounds only gate projections: the flat views carry already-gated queries
    }
    return c;
}
inline void set_statics(const Camera& c) {
    Frame::fx = c.fx; Frame::fy = c.fy; Frame::cx = c.cx; Frame::cy = c.cy;
    Frame::invfx = 1.0f / c.fx; Frame::invfy = 1.0f / c.fy;
    Frame::mnMinX = c.minX; Frame::mnMaxX = c.maxX; Frame::mnMinY = c.minY; Frame::mnMaxY = c.maxY;
    Frame::mfGridElementWidthInv = c.gridWInv; Frame::mfGridElementHeightInv = c.gridHInv;
    Frame::mbInitialComputations = false;
}

inline cv::Mat eye4() { cv::Mat m = cv::Mat::eye(4, 4, CV_32F); return m; }
inline cv::Mat vec3(float x, float y, float z) { cv::Mat m(3, 1, CV_32F); m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z; return m; }
inline cv::Mat desc_row(const uint8_t* d) { cv::Mat m(1, 32, CV_8U); std::memcpy(m.data, d, 32); return m; }

// Everything a call creates; destroyed in one go.
struct World {
    Map map;
    std::vector<Frame*> frames;
    std::vector<KeyFrame*> kfs;
    std::vector<MapPoint*> mps;
    KeyFrame* anchor;   // reference key frame of the map points (MapPoint's constructor reads its ids)
    World() : anchor(nullptr) {
        const float one = 1.f;
        anchor = keyframe(frame(nullptr, 0, &one, 1, nullptr, identity_camera(nullptr), false));
    }
    ~World() {
        for (MapPoint* p : mps) delete p;
        for (KeyFrame* k : kfs) delete k;
        for (Frame* f : frames) delete f;
    }

    // A Frame whose public members hold frame `f` of the flat view.
    Frame* frame(const orbgpu_frame_set* fs, int f, const float* scale, int n_levels, const float* sigma2, const Camera& cam,
                 bool with_grid) {
        set_statics(cam);
        Frame* F = new Frame();
        frames.push_back(F);
        const int k0 = fs ? fs->kp_off[f] : 0, n = fs ? fs->kp_off[f + 1] - k0 : 0;
        F->mpORBvocabulary = nullptr; F->mpORBextractorLeft = F->mpORBextractorRight = nullptr;
        F->mTimeStamp = 0; F->mbf = cam.mbf; F->mb = cam.mb; F->mThDepth = 0; F->N = n;
        F->mnId = Frame::nNextId++; F->mpReferenceKF = nullptr;
        if (n) {
            const cv::KeyPoint* kp = reinterpret_cast<const cv::KeyPoint*>(fs->keys_un + k0);
            F->mvKeys.assign(kp, kp + n);
            F->mvKeysUn = F->mvKeys;
        }
        F->mvuRight.assign(n, -1.f);
        F->mvDepth.assign(n, -1.f);
        if (fs && fs->u_right) for (int i = 0; i < n; ++i) F->mvuRight[i] = fs->u_right[k0 + i];
        F->mDescriptors = cv::Mat(n, 32, CV_8U);
        if (n) std::memcpy(F->mDescriptors.data, fs->desc + (size_t)k0 * 32, (size_t)n * 32);
        F->mvpMapPoints.assign(n, static_cast<MapPoint*>(nullptr));
        F->mvbOutlier.assign(n, false);
        F->mnScaleLevels = n_levels;
        F->mfScaleFactor = n_levels > 1 ? scale[1] : 1.2f;
        F->mfLogScaleFactor = std::log(F->mfScaleFactor);
        F->mvScaleFactors.assign(scale, scale + n_levels);
        F->mvInvScaleFactors.resize(n_levels); F->mvLevelSigma2.resize(n_levels); F->mvInvLevelSigma2.resize(n_levels);
        for (int l = 0; l < n_levels; ++l) {
            F->mvInvScaleFactors[l] = 1.0f / scale[l];
            F->mvLevelSigma2[l] = sigma2 ? sigma2[l] : scale[l] * scale[l];
            F->mvInvLevelSigma2[l] = 1.0f / F->mvLevelSigma2[l];
        }
        F->mTcw = eye4();
        F->UpdatePoseMatrices();
        if (fs && fs->fv_node_off) {
            for (int a = fs->fv_node_off[f]; a < fs->fv_node_off[f + 1]; ++a) {
                std::vector<unsigned int>& v = F->mFeatVec[(DBoW2::NodeId)fs->fv_node_id[a]];
                for (int i = fs->fv_feat_off[a]; i < fs->fv_feat_off[a + 1]; ++i) v.push_back((unsigned int)fs->fv_feat[i]);
            }
        }
        if (with_grid) F->AssignFeaturesToGrid();
        return F;
    }
    KeyFrame* keyframe(Frame* F) {
        KeyFrame* k = new KeyFrame(*F, &map, nullptr);
        kfs.push_back(k);
        return k;
    }
    MapPoint* mappoint(const cv::Mat& pos, const uint8_t* desc, int n_obs, bool bad) {
        MapPoint* p = new MapPoint(pos, anchor, &map);
        mps.push_back(p);
        p->nObs = n_obs;
        p->mbBad = bad;
        if (desc) p->mDescriptor = desc_row(desc);
        p->mfMinDistance = 0.f;
        p->mfMaxDistance = 0.f;
        return p;
    }
    // the occupant of a key point that a flat view flags as "holds a MapPoint"
    MapPoint* occupant(int n_obs) { return mappoint(vec3(0, 0, 0), nullptr, n_obs, false); }
};

}  // namespace slamworld
#endif

// Stand-ins for the ORB-SLAM2 classes the matcher shell (csrc/host/ORBmatcher_gpu.cc) reads, declaring ONLY the members
// the four GPU-backed search functions touch, with the reference's names and types
// (/root/reference/include/Frame.h:113-188, KeyFrame.h:52-184, MapPoint.h:52-97, Thirdparty/DBoW2/DBoW2/FeatureVector.h:21).
// They exist so the shell can be compiled and exercised where neither OpenCV nor the reference's sources are
// available; in a real build the reference's own headers come first on the include path and these are never seen.
#ifndef ORBGPU_STUB_FRAME_H
#define ORBGPU_STUB_FRAME_H

#include <map>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

namespace DBoW2 {
typedef unsigned int NodeId;
typedef unsigned int WordId;    // BowVector.h:21
typedef double WordValue;       // BowVector.h:24
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
class BowVector : public std::map<WordId, WordValue> {};   // BowVector.h:58-60
}  // namespace DBoW2

namespace ORB_SLAM2 {

class KeyFrame;
class Frame;

class MapPoint {
public:
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0), mTrackViewCos(0),
                 mbBad(false), nObs(0), mfMinDistance(0), mfMaxDistance(0) {}
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }   // MapPoint.cc:392-396
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }   // MapPoint.cc:398-402
    inline int PredictScale(const float& currentDist, Frame* pF);       // MapPoint.cc:421-436
    inline int PredictScale(const float& currentDist, KeyFrame* pKF);   // MapPoint.cc:404-419
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }        // MapPoint.cc:325-329
    int GetIndexInKeyFrame(KeyFrame* pKF) {                                            // MapPoint.cc:240-246
        std::map<KeyFrame*, size_t>::iterator it = mObservations.find(pKF);
        return it == mObservations.end() ? -1 : (int)it->second;
    }
    void AddObservation(KeyFrame* pKF, size_t idx) {                                   // MapPoint.cc:85-96 (monocular count)
        if (mObservations.count(pKF)) return;
        mObservations[pKF] = idx;
        nObs++;
    }
    inline void Replace(MapPoint* pMP);                                                // MapPoint.cc:189-238
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }

    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;

    // stub state
    bool mbBad;
    int nObs;
    float mfMinDistance, mfMaxDistance;
    cv::Mat mDescriptor;
    cv::Mat mWorldPos;   // 3x1 CV_32F
    cv::Mat mNormalVector;   // 3x1 CV_32F
    std::map<KeyFrame*, size_t> mObservations;
    MapPoint* mpReplaced = 0;
};

class Frame {
public:
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors;
    float mfLogScaleFactor;
    int mnScaleLevels;
    cv::Mat mTcw;        // 4x4 CV_32F
    float mb, mbf;
    static float fx, fy, cx, cy;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};

inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = (int)std::ceil(std::log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

class KeyFrame {
public:
    KeyFrame(const std::vector<cv::KeyPoint>& keys, const std::vector<float>& uright, const cv::Mat& desc,
             const std::vector<float>& sf, const std::vector<float>& s2, float _fx, float _fy, float _cx, float _cy)
        : fx(_fx), fy(_fy), cx(_cx), cy(_cy), N((int)keys.size()), mvKeysUn(keys), mvuRight(uright), mDescriptors(desc),
          mvScaleFactors(sf), mvLevelSigma2(s2), mvpMapPoints(keys.size(), (MapPoint*)0) {}
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints() {                                                                  // KeyFrame.cc:248-261
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); ++i)
            if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }                       // KeyFrame.cc:219-223
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }   // KeyFrame.cc:624-627

    const float fx, fy, cx, cy;
    const int N;
    const std::vector<cv::KeyPoint> mvKeysUn;
    const std::vector<float> mvuRight;
    const cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    const std::vector<float> mvScaleFactors;
    const std::vector<float> mvLevelSigma2;

    float mbf = 0.f, mfLogScaleFactor = 0.f;
    int mnScaleLevels = 0;
    std::vector<float> mvInvLevelSigma2;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    float mfGridElementWidthInv = 0.f, mfGridElementHeightInv = 0.f;

    // stub state
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Ow, Rcw, tcw;   // 3x1, 3x3, 3x1 CV_32F
};

inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = (int)std::ceil(std::log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}

// MapPoint::Replace (MapPoint.cc:189-238): this point goes bad, its observations move to pMP (or are erased where pMP is seen already)
inline void MapPoint::Replace(MapPoint* pMP) {
    if (pMP == this) return;
    std::map<KeyFrame*, size_t> obs = mObservations;
    mObservations.clear();
    mbBad = true;
    mpReplaced = pMP;
    for (std::map<KeyFrame*, size_t>::iterator it = obs.begin(); it != obs.end(); ++it) {
        KeyFrame* pKF = it->first;
        if (!pMP->IsInKeyFrame(pKF)) {
            pKF->mvpMapPoints[it->second] = pMP;   // ReplaceMapPointMatch
            pMP->AddObservation(pKF, it->second);
        } else {
            pKF->mvpMapPoints[it->second] = static_cast<MapPoint*>(0);   // EraseMapPointMatch
        }
    }
}

}  // namespace ORB_SLAM2
#endif

// cv::FAST (ORBextractor.cc:809,814) and KeyPointsFilter::retainBest (only reached from the dead
// ComputeKeyPointsOld, ORBextractor.cc:1006,1024).  Definitions: oracle/cvlite_shim.cc.
#ifndef ORBGPU_SHIM_OPENCV2_FEATURES2D_HPP
#define ORBGPU_SHIM_OPENCV2_FEATURES2D_HPP
#include <opencv2/core/core.hpp>
namespace cv {
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};
}
#endif

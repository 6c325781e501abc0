// Declarations of the three imgproc calls on the hot path (ORBextractor.cc:1086,1120,1122-1128).
// Definitions: oracle/cvlite_shim.cc (restated OpenCV 4.13 arithmetic; oracle builds only).
#ifndef ORBGPU_SHIM_OPENCV2_IMGPROC_HPP
#define ORBGPU_SHIM_OPENCV2_IMGPROC_HPP
#include <opencv2/core/core.hpp>
namespace cv {
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
}
#endif

// TEST INFRASTRUCTURE (CPU oracle).  Binds the shim's cv:: free functions (shim/opencv2/**) to the
// cvlite restatements so that /root/reference/src/ORBextractor.cc compiles and runs verbatim
// without OpenCV (oracle/Makefile -> oracle/_ref/).
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#include "cvlite.h"

namespace cv {

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int) {
    // ORBextractor.cc:1120 passes a destination that already is the ROI inside the bordered level
    // buffer; create() keeps it when the geometry matches (it must NOT reallocate in that case).
    if (!(dst.data && dst.rows == dsize.height && dst.cols == dsize.width)) dst.create(dsize.height, dsize.width, src.type());
    cvl_resize_linear_u8(src.data, src.cols, src.rows, (int)(size_t)src.step, dst.data, dst.cols, dst.rows, (int)(size_t)dst.step);
}

void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int) {
    const int H = src.rows + top + bottom, W = src.cols + left + right;
    if (!(dst.data && dst.rows == H && dst.cols == W)) dst.create(H, W, src.type());
    cvl_border_reflect101_u8(src.data, src.cols, src.rows, (int)(size_t)src.step, dst.data, (int)(size_t)dst.step, top, bottom, left, right);
}

void GaussianBlur(const Mat& src, Mat& dst, Size, double, double, int) {
    if (!(dst.data && dst.rows == src.rows && dst.cols == src.cols)) dst.create(src.rows, src.cols, src.type());
    cvl_gaussian7x7_u8(src.data, src.cols, src.rows, (int)(size_t)src.step, dst.data, (int)(size_t)dst.step);
}

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    keypoints.clear();
    static thread_local std::vector<cvl_kp> buf;
    const int cap = (image.rows * image.cols) / 2 + 16;
    if ((int)buf.size() < cap) buf.resize(cap);
    const int n = cvl_fast9_16(image.data, image.cols, image.rows, (int)(size_t)image.step, threshold, nonmaxSuppression ? 1 : 0, buf.data(), cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)buf[i].x, (float)buf[i].y, 7.f, -1.f, (float)buf[i].score));
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
    // only reachable from the dead ComputeKeyPointsOld (ORBextractor.cc:855-1032)
    if (npoints >= 0 && (int)keypoints.size() > npoints) {
        std::stable_sort(keypoints.begin(), keypoints.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        keypoints.resize(npoints);
    }
}

float fastAtan2(float y, float x) { return cvl_fast_atan2(y, x); }

}  // namespace cv

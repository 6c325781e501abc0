// TEST INFRASTRUCTURE.  The reference's include/Converter.h pulls in Eigen and g2o (include/Converter.h:26-28), neither of
// which exists in this image; Frame.cc and KeyFrame.cc use exactly one member of it, toDescriptorVector (Frame.cc:429,
// KeyFrame.cc:63; defined at src/Converter.cc:27-35: one cv::Mat row header per descriptor row).  This directory precedes
// the reference's include/ on the include path of the slamref build, so those two files see this declaration instead.
#ifndef CONVERTER_H
#define CONVERTER_H
#include <opencv2/core/core.hpp>
#include <vector>
namespace ORB_SLAM2 {
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
        std::vector<cv::Mat> v;
        v.reserve(Descriptors.rows);
        for (int j = 0; j < Descriptors.rows; ++j) v.push_back(Descriptors.row(j));
        return v;
    }
};
}  // namespace ORB_SLAM2
#endif

// Legacy umbrella header named by ORBextractor.h:26.
#ifndef ORBGPU_SHIM_OPENCV_CV_H
#define ORBGPU_SHIM_OPENCV_CV_H
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv2/features2d/features2d.hpp>
#endif

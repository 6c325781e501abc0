// TEST INFRASTRUCTURE.  <opencv2/opencv.hpp> for compiling the reference's Frame.h / KeyFrame.h where they lie
// (oracle/Makefile, target slamref): everything they need is in the cv:: stand-in's core header.
#ifndef ORBGPU_SLAM_SHIM_OPENCV_HPP
#define ORBGPU_SLAM_SHIM_OPENCV_HPP
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#endif

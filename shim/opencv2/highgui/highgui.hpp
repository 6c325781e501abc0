// highgui is included by ORBextractor.cc:58 but nothing from it is used on the hot path.
#ifndef ORBGPU_SHIM_OPENCV2_HIGHGUI_HPP
#define ORBGPU_SHIM_OPENCV2_HIGHGUI_HPP
#include <opencv2/core/core.hpp>
#endif

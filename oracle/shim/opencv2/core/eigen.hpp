// oracle/shim/opencv2/core/eigen.hpp -- TEST INFRASTRUCTURE.  The reference's vins_pnp.hpp includes <opencv2/core/eigen.hpp> but uses
// nothing from it except, through the TS()/TE() timing macros of global_param.hpp:86-88, cv::getTickCount / cv::getTickFrequency and
// the int64 typedef.  OpenCV's C++ headers are not in this image, so this shim provides exactly those three names and nothing else.
#pragma once
#include <chrono>
typedef long long int64;
namespace cv {
inline int64 getTickCount() { return (int64)std::chrono::steady_clock::now().time_since_epoch().count(); }
inline double getTickFrequency() { return 1e9; }
}  // namespace cv

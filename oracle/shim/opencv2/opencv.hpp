// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE.  The reference's initial_aligment.hpp includes feature_manager.hpp, which
// includes feature_tracker.hpp, which includes <opencv2/opencv.hpp> and DECLARES members of type cv::Mat / cv::Point2f.  Nothing in
// initial_aligment.cpp touches them, and OpenCV's C++ headers are not in this image, so this shim declares the two type names (empty
// shells, never instantiated) -- enough for the unmodified initial_aligment.cpp to compile where it lies.
#pragma once
#include "core/eigen.hpp"
namespace cv {
struct Mat {};
struct Point2f { float x, y; };
}  // namespace cv

// Stand-in for <opencv2/features2d/features2d.hpp>: cv::KeyPoint only (see ../core/core.hpp).  TEST INFRASTRUCTURE.
#pragma once
#include "../core/core.hpp"
namespace cv {
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
}  // namespace cv

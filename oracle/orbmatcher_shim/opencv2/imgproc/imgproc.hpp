// Stand-in for <opencv2/imgproc/imgproc.hpp>: the drawing calls of Detector2D::draw_objects compile and do nothing (the viewer is out of scope).  TEST INFRASTRUCTURE.
#pragma once
#include <string>
#include "../core/core.hpp"
#define CV_FILLED -1
namespace cv {
enum { FONT_HERSHEY_SIMPLEX = 0 };
template <class R> inline void rectangle(Mat&, const R&, const Scalar&, int = 1) {}
inline Size getTextSize(const std::string&, int, double, int, int* baseLine) { if (baseLine) *baseLine = 0; return Size(0, 0); }
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1) {}
}  // namespace cv

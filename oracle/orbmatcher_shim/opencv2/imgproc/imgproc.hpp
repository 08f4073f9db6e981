// Stand-in for <opencv2/imgproc/imgproc.hpp>: the drawing calls of Detector2D::draw_objects compile and do nothing (the viewer is out of scope).  TEST INFRASTRUCTURE.
#pragma once
#include <cstdint>
#include <string>
#include "../core/core.hpp"
#define CV_FILLED -1
namespace cv {
enum { FONT_HERSHEY_SIMPLEX = 0 };
template <class R> inline void rectangle(Mat&, const R&, const Scalar&, int = 1) {}
inline Size getTextSize(const std::string&, int, double, int, int* baseLine) { if (baseLine) *baseLine = 0; return Size(0, 0); }
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1) {}
// ---- what src/ORBextractor.cc calls: OpenCV's algorithms are the oracle's restatements (oracle/sgs_oracle.cpp: sgo_resize, sgo_blur, sgo_fast_atan2 -- each
// pinned bit for bit against the real cv2 primitive by tests/test_oracle_golden.py), so that the reference's OWN extractor code runs on top of pinned primitives
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1 };
extern "C" int sgo_resize(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch);
extern "C" int sgo_blur(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
extern "C" float sgo_fast_atan2(float y, float x);
inline float fastAtan2(float y, float x) { return sgo_fast_atan2(y, x); }
inline void resize(InputArray src_, OutputArray dst_, Size sz, double, double, int) {
    const Mat src = src_.getMat();
    dst_.create(sz.height, sz.width, src.type());                    // a destination of the right size keeps its memory (the pyramid level is a view into its padded buffer)
    Mat dst = dst_.getMat();
    sgo_resize(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}
inline int reflect101(int i, int n) { while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i; return i; }
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int) {
    const Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    Mat keep = src.clone();                                          // src may be a view into dst (BORDER_ISOLATED use of the extractor)
    for (int y = 0; y < dst.rows; ++y)
        for (int x = 0; x < dst.cols; ++x) dst.at<uchar>(y, x) = keep.at<uchar>(reflect101(y - top, src.rows), reflect101(x - left, src.cols));
}
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size, double, double, int) {
    const Mat src = src_.getMat().clone();
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    sgo_blur(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);
}
}  // namespace cv

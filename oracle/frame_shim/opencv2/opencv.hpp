// Stand-in for <opencv2/opencv.hpp> as include/Frame.h and src/Frame.cc use it: the cv::Mat / KeyPoint stand-ins of orbmatcher_shim plus the OpenCV ALGORITHMS
// Frame.cc calls, which resolve to the oracle's restatements -- cv::calcOpticalFlowPyrLK = sgo_lk_track, cv::findFundamentalMat = sgo_find_fundamental_ransac,
// cv::undistortPoints = sgo_undistort_points (each pinned against the real cv2 primitive by tests/golden/*.npz) -- so that the reference's OWN Frame.cc
// (RGB-D constructor, RmDynamicPointWithSemanticAndGeometry, isInFrustum, the grid) runs unmodified on top of pinned primitives.  TEST INFRASTRUCTURE.
#pragma once
#include <cstdint>
#include <vector>
#include "opencv2/core/core.hpp"
#include "opencv2/features2d/features2d.hpp"
#include "opencv2/imgproc/imgproc.hpp"

#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS 2

extern "C" int sgo_lk_track(const uint8_t* I0, const uint8_t* J0, int w, int h, int pitch, const float* pts, int n, float* out);
extern "C" int sgo_find_fundamental_ransac(const float* m1, const float* m2, int n, double thresh, double confidence, int max_iters, double* F, uint8_t* mask_out, int32_t* info);
extern "C" int sgo_undistort_points(const float* xy, int n, float fxf, float fyf, float cxf, float cyf, const float* dist5, float* out_xy);

namespace cv {
enum { FM_RANSAC = 8, NORM_L1 = 2 };
struct TermCriteria { int type, maxCount; double epsilon; TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {} };

inline MatExpr operator-(const Mat& a, const MatExpr& b) { MatExpr e; e.kind = MatExpr::ADD; e.a = a; e.b = b.eval(); e.sign = -1; return e; }
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {
    double s = 0;
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}
// cv::Mat_<float>(3,1) << x, y, z   (Frame::UnprojectStereo)
template <class T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_32F) {}
    struct Comma {
        Mat m; int i;
        Comma& operator,(T v) { m.at<T>(i / m.cols, i % m.cols) = v; ++i; return *this; }
        operator Mat() const { return m; }
    };
    Comma operator<<(T v) { Comma c{*this, 0}; return c, v; }
};

// calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status, err, Size(21,21), 3, {ITER|EPS, 30, 0.01}): the oracle's restatement of video/lkpyramid.cpp
// (the parameters of the one call site of the reference are the restatement's constants; status / err are not read by the reference)
inline void calcOpticalFlowPyrLK(const Mat& prev, const Mat& next, const std::vector<Point2f>& p0, std::vector<Point2f>& p1, std::vector<uchar>& status, std::vector<float>& err,
                                 Size win, int maxLevel, TermCriteria crit) {
    assert(win.width == 21 && win.height == 21 && maxLevel == 3 && crit.maxCount == 30 && crit.epsilon == 0.01);
    const int n = (int)p0.size();
    std::vector<float> in(2 * (size_t)n + 2), out(2 * (size_t)n + 2);
    for (int i = 0; i < n; ++i) { in[2 * i] = p0[i].x; in[2 * i + 1] = p0[i].y; }
    sgo_lk_track(prev.data, next.data, prev.cols, prev.rows, (int)prev.step, in.data(), n, out.data());
    p1.resize(n); status.assign(n, 1); err.assign(n, 0.f);
    for (int i = 0; i < n; ++i) p1[i] = Point2f(out[2 * i], out[2 * i + 1]);
}
// findFundamentalMat(points1, points2, FM_RANSAC, 1.0, 0.99): RANSAC / LMedS / 7-point branches of calib3d/fundam.cpp as restated by oracle/fundamental.cpp;
// an empty Mat when no model was found (the reference then reads F12.at<double> of an empty matrix: quirk Q11 -- the pin's scenarios stay away from it)
inline Mat findFundamentalMat(const std::vector<Point2f>& a, const std::vector<Point2f>& b, int method, double thresh, double conf) {
    assert(method == FM_RANSAC);
    const int n = (int)a.size();
    std::vector<float> m1(2 * (size_t)n + 2), m2(2 * (size_t)n + 2);
    for (int i = 0; i < n; ++i) { m1[2 * i] = a[i].x; m1[2 * i + 1] = a[i].y; m2[2 * i] = b[i].x; m2[2 * i + 1] = b[i].y; }
    double F[9];
    if (!sgo_find_fundamental_ransac(m1.data(), m2.data(), n, thresh, conf, 1000, F, nullptr, nullptr)) return Mat();
    Mat m(3, 3, CV_64F);
    for (int i = 0; i < 9; ++i) m.at<double>(i / 3, i % 3) = F[i];
    return m;
}
// undistortPoints(src, dst, K, dist, R = empty, P = K) on N x 2 floats, in place
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat&, const Mat&) {
    const int n = src.rows;
    std::vector<float> in(2 * (size_t)n + 2), out(2 * (size_t)n + 2);
    for (int i = 0; i < n; ++i) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
    float d5[5] = {0, 0, 0, 0, 0};
    const int nd = dist.rows * dist.cols;
    for (int i = 0; i < 5 && i < nd; ++i) d5[i] = dist.at<float>(i);
    sgo_undistort_points(in.data(), n, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d5, out.data());
    for (int i = 0; i < n; ++i) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
}
}  // namespace cv

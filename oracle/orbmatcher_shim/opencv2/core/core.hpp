// Stand-in for <opencv2/core/core.hpp>, just enough to compile the reference's own src/ORBmatcher.cc WITHOUT OpenCV as the parity pin of the matcher
// control flow (oracle/_ref/liborbmatcher_ref.so, recipe in oracle/Makefile).  TEST INFRASTRUCTURE: nothing under sg-slam_b200/ or include/ includes it.
//
// cv::Mat here is a reference-counted 2-D array of CV_32F or CV_8U with row / column views.  The float arithmetic the matchers use goes through a small
// MatExpr so that it is evaluated the way OpenCV evaluates it -- the rules were probed with cv2 and are pinned by tests/golden/frustum.npz
// (tests/golden/make_golden_frustum.py evaluates the reference's expressions with the real cv2.gemm / cv2.norm):
//   * A*B (+C) with untransposed small operands = gemm(A, B, alpha, C, beta) on OpenCV's small-matrix path: float products summed in float, left to
//     right, then one rounding of (double)sum * alpha + (double)c * beta;
//   * a transposed operand (A.t()*B, as in -Rcw.t()*tcw) takes the general path: double accumulator, one cast;
//   * cv::norm and Mat::dot accumulate in double.
// Scalar scaling (s*A, A/s: Sim3 helpers of the loop-closing matchers) is plain float arithmetic here and NOT part of the pin.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6
typedef unsigned char uchar;

namespace cv {

struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
    Point2f& operator*=(float s) { x = x * s; y = y * s; return *this; }      // Point_<float> *= float: one float multiply per coordinate
};
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };      // include/Converter.h:49
// geometry PODs of the detector's post-processing and of its (never executed) drawing helper
struct Point { int x = 0, y = 0; Point() {} Point(int x_, int y_) : x(x_), y(y_) {} };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
template <class T> struct Rect_ {
    T x = 0, y = 0, width = 0, height = 0;
    Rect_() {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point& p, const Size& s) : x((T)p.x), y((T)p.y), width((T)s.width), height((T)s.height) {}
    template <class U> Rect_(const Rect_<U>& r) : x((T)r.x), y((T)r.y), width((T)r.width), height((T)r.height) {}
};
typedef Rect_<int> Rect;
typedef Point Point2i;

class MatExpr;

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;                                       // bytes between rows
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = (size_t)c * esz();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step, (uchar)0);
        data = buf_->data();
    }
    int type() const { return type_; }
    bool empty() const { return !data || rows * cols == 0; }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz());
        return m;
    }
    // Mat::zeros is a MatExpr in OpenCV: ASSIGNING it to a matrix of the same size and type clears that matrix IN PLACE (ORBextractor's computeDescriptors
    // relies on it: its `descriptors = Mat::zeros(...)` must keep writing into the row range of the caller's output)
    struct ZerosExpr { int r, c, type; };
    static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
    Mat(const ZerosExpr& z) { create(z.r, z.c, z.type); }
    Mat& operator=(const ZerosExpr& z) {
        if (data && rows == z.r && cols == z.c && type_ == z.type) { for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols * esz()); }
        else create(z.r, z.c, z.type);
        return *this;
    }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < std::min(r, c); ++i) m.at<float>(i, i) = 1.f; return m; }
    Mat view(int r0, int r1, int c0, int c1) const {
        Mat m; m.type_ = type_; m.buf_ = buf_; m.rows = r1 - r0; m.cols = c1 - c0; m.step = step; m.data = data + (size_t)r0 * step + (size_t)c0 * esz();
        return m;
    }
    Mat operator()(const Rect_<int>& r) const { return view(r.y, r.y + r.height, r.x, r.x + r.width); }
    size_t step1() const { return step / esz(); }
    Size size() const { return Size(cols, rows); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    Mat row(int r) const { return view(r, r + 1, 0, cols); }
    Mat col(int c) const { return view(0, rows, c, c + 1); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }                       // vectors only
    template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    MatExpr t() const;
    double dot(const Mat& o) const {
        double s = 0;
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
        return s;
    }
    // OpenCV's copyTo writes INTO a destination that already has the right size and type (views included) and reallocates otherwise
    void copyTo(Mat& o) const {
        if (o.data && o.rows == rows && o.cols == cols && o.type_ == type_) { for (int r = 0; r < rows; ++r) std::memmove(o.data + (size_t)r * o.step, data + (size_t)r * step, (size_t)cols * esz()); }
        else o = clone();
    }
    void copyTo(Mat&& view) const { Mat& o = view; copyTo(o); }       // Rcw.copyTo(Tcw.rowRange(0,3).colRange(0,3))
    // ---- what src/Frame.cc needs on top (oracle/frame_shim, the Frame.cc pin)
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.f; return m; }      // CV_32F only
    void convertTo(Mat& o, int type) const {                 // CV_8U -> CV_32F (the stereo matcher's patches)
        Mat m(rows, cols, type);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(r, c) = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
        o = m;
    }
    // ---- what src/Tracking.cc needs on top (oracle/tracking_shim, the tracking front-end pin; none of it is on the pinned path)
    int channels() const { return 1; }
    Mat inv() const;                                         // declared only: mapping-thread sources that are compiled for the drop-in check, never linked or run
    void resize(size_t nrows) { Mat m((int)nrows, cols, type_); for (int r = 0; r < std::min(rows, (int)nrows); ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz()); *this = m; }
    void convertTo(Mat& o, int type, double scale) const {
        Mat m(rows, cols, type);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(r, c) = (float)((type_ == CV_32F ? (double)at<float>(r, c) : (double)at<uchar>(r, c)) * scale);
        o = m;
    }
    explicit Mat(const Point3f& p) { create(3, 1, CV_32F); at<float>(0) = p.x; at<float>(1) = p.y; at<float>(2) = p.z; }
    Mat reshape(int) const { return *this; }                // N x 2 one-channel <-> N x 1 two-channel: the same memory; cv::undistortPoints below reads N x 2 floats
    void push_back(const Mat& row) {                         // append one row (a fresh buffer: the destination of the reference's compaction loop is never a view)
        const int c = rows ? cols : row.cols, t = rows ? type_ : row.type_;
        Mat m(rows + 1, c, t);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)c * m.esz());
        std::memcpy(m.data + (size_t)rows * m.step, row.data, (size_t)c * m.esz());
        *this = m;
    }
private:
    int type_ = CV_8U;
    size_t esz() const { return type_ == CV_64F ? 8 : type_ == CV_32F ? 4 : 1; }
    std::shared_ptr<std::vector<uchar>> buf_;
};

// alpha * op(a) [* b] [+ beta * c]   or   a + sign * b
class MatExpr {
public:
    enum Kind { SCALE, GEMM, ADD };
    Kind kind = SCALE;
    Mat a, b, c;
    bool ta = false, has_c = false;
    double alpha = 1, beta = 0, sign = 1;
    Mat eval() const {
        if (kind == ADD) {
            Mat m(a.rows, a.cols, CV_32F);
            for (int r = 0; r < a.rows; ++r) for (int cc = 0; cc < a.cols; ++cc) m.at<float>(r, cc) = sign > 0 ? a.at<float>(r, cc) + b.at<float>(r, cc) : a.at<float>(r, cc) - b.at<float>(r, cc);
            return m;
        }
        if (kind == SCALE) {
            const int R = ta ? a.cols : a.rows, C = ta ? a.rows : a.cols;
            Mat m(R, C, CV_32F);
            for (int r = 0; r < R; ++r) for (int cc = 0; cc < C; ++cc) { const float v = ta ? a.at<float>(cc, r) : a.at<float>(r, cc); m.at<float>(r, cc) = alpha == 1 ? v : alpha == -1 ? -v : (float)alpha * v; }
            return m;
        }
        const int M = ta ? a.cols : a.rows, K = ta ? a.rows : a.cols, N = b.cols;
        Mat m(M, N, CV_32F);
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                if (!ta) {                                // small-matrix path: float sum, left to right
                    float acc = 0.f;
                    for (int k = 0; k < K; ++k) { const float p = a.at<float>(i, k) * b.at<float>(k, j); acc = k == 0 ? p : acc + p; }
                    m.at<float>(i, j) = (float)((double)acc * alpha + (has_c ? (double)c.at<float>(i, j) * beta : 0.0));
                } else {                                  // general path: double accumulator
                    double acc = 0;
                    for (int k = 0; k < K; ++k) acc += (double)a.at<float>(k, i) * (double)b.at<float>(k, j);
                    m.at<float>(i, j) = (float)(acc * alpha + (has_c ? (double)c.at<float>(i, j) * beta : 0.0));
                }
            }
        return m;
    }
    MatExpr t() const { MatExpr e = *this; if (e.kind == SCALE) e.ta = !e.ta; else { Mat m = eval(); e = MatExpr(); e.a = m; e.ta = true; } return e; }
    template <class T> T at(int i) const { return eval().at<T>(i); }
    Mat inv() const;                                         // declared only (see Mat::inv)
};
inline Mat::Mat(const MatExpr& e) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) { *this = e.eval(); return *this; }
inline MatExpr Mat::t() const { MatExpr e; e.a = *this; e.ta = true; return e; }

inline MatExpr operator*(const Mat& a, const Mat& b) { MatExpr e; e.kind = MatExpr::GEMM; e.a = a; e.b = b; return e; }
inline MatExpr operator*(const MatExpr& x, const Mat& b) {
    MatExpr e;
    if (x.kind == MatExpr::SCALE) { e.kind = MatExpr::GEMM; e.a = x.a; e.ta = x.ta; e.alpha = x.alpha; e.b = b; }
    else { e.kind = MatExpr::GEMM; e.a = x.eval(); e.b = b; }
    return e;
}
inline MatExpr operator+(const MatExpr& x, const Mat& c) {
    if (x.kind == MatExpr::GEMM && !x.has_c) { MatExpr e = x; e.c = c; e.has_c = true; e.beta = 1; return e; }
    MatExpr e; e.kind = MatExpr::ADD; e.a = x.eval(); e.b = c; return e;
}
inline MatExpr operator+(const Mat& a, const Mat& b) { MatExpr e; e.kind = MatExpr::ADD; e.a = a; e.b = b; return e; }
inline MatExpr operator-(const Mat& a, const Mat& b) { MatExpr e; e.kind = MatExpr::ADD; e.a = a; e.b = b; e.sign = -1; return e; }
inline MatExpr operator-(const Mat& a) { MatExpr e; e.a = a; e.alpha = -1; return e; }
inline MatExpr operator-(const MatExpr& x) { MatExpr e = x; if (e.kind == MatExpr::ADD) { Mat m = x.eval(); e = MatExpr(); e.a = m; e.alpha = -1; } else e.alpha = -e.alpha; return e; }
inline MatExpr operator*(double s, const Mat& a) { MatExpr e; e.a = a; e.alpha = s; return e; }
inline MatExpr operator*(const Mat& a, double s) { MatExpr e; e.a = a; e.alpha = s; return e; }
inline MatExpr operator*(double s, const MatExpr& x) { MatExpr e = x; if (e.kind == MatExpr::ADD) { Mat m = x.eval(); e = MatExpr(); e.a = m; e.alpha = s; } else e.alpha *= s; return e; }
inline MatExpr operator/(const Mat& a, double s) { MatExpr e; e.a = a; e.alpha = (double)(1.f / (float)s); return e; }

inline double norm(const Mat& m) {
    double s = 0;
    for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) { const double v = m.at<float>(r, c); s += v * v; }
    return std::sqrt(s);
}
inline double norm(const MatExpr& e) { return norm(e.eval()); }

// array proxies of the extractor's call operator
class _InputArray { const Mat* m_; public: _InputArray(const Mat& m) : m_(&m) {} bool empty() const { return m_->empty(); } Mat getMat() const { return *m_; } };
typedef const _InputArray& InputArray;
class _OutputArray {
    Mat* m_;
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void release() const { *m_ = Mat(); }
    void create(int r, int c, int type) const { if (m_->rows != r || m_->cols != c || m_->type() != type || !m_->data) m_->create(r, c, type); }
    Mat getMat() const { return *m_; }
};
typedef const _OutputArray& OutputArray;

}  // namespace cv

// cvRound = round half to even (SSE cvtsd2si / cvtss2si in OpenCV), cvFloor, cvCeil
inline int cvRound(double v) { return (int)std::lrint(v); }
inline int cvRound(float v) { return (int)std::lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

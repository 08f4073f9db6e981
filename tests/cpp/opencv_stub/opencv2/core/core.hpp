// Stand-in for the REAL <opencv2/core/core.hpp> with OpenCV's actual API shape where it differs from the repo's cv_compat.h: cv::InputArray /
// cv::OutputArray are proxy classes (_InputArray / _OutputArray) without data / cols / rows / step / ptr members -- only getMat(), empty(), type(),
// create(), release().  tests/test_cpp_headers.py compiles include/sgslam/*.h against this stub with -std=c++11 -DSGS_WITH_OPENCV (the
// reference's language level, CMakeLists.txt:20-30) to prove the mirror headers only use what real OpenCV offers.  Compile-only: bodies are trivial.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6

namespace cv {
template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} Point_& operator*=(T s) { x *= s; y *= s; return *this; } };
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;
template <class T> struct Rect_ { T x, y, width, height; Rect_() : x(0), y(0), width(0), height(0) {} Rect_(T a, T b, T c, T d) : x(a), y(b), width(c), height(d) {} };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} };
struct MatStep { size_t p; MatStep() : p(0) {} operator size_t() const { return p; } };
class Mat {
public:
    int flags, dims, rows, cols;
    unsigned char* data;
    MatStep step;
    Mat() : flags(0), dims(2), rows(0), cols(0), data(0) {}
    Mat(int r, int c, int) : flags(0), dims(2), rows(r), cols(c), data(0) {}
    Mat(int r, int c, int, void* d, size_t s = 0) : flags(0), dims(2), rows(r), cols(c), data((unsigned char*)d) { step.p = s; }
    void create(int r, int c, int) { rows = r; cols = c; }
    void release() { rows = cols = 0; data = 0; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    int type() const { return 0; }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.p); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.p); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    Mat row(int r) const { return Mat(1, cols, 0, data + (size_t)r * step.p, step.p); }
    Mat clone() const { return *this; }
};
class _InputArray {
public:
    _InputArray() : m_(0) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat(int = -1) const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    int type(int = -1) const { return 0; }
    int rows(int = -1) const { return m_ ? m_->rows : 0; }          // member FUNCTIONS, as in OpenCV
    int cols(int = -1) const { return m_ ? m_->cols : 0; }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int t, int = -1, bool = false, int = 0) const { m_->create(r, c, t); }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
}  // namespace cv

// cv_compat.h -- the handful of OpenCV value types the ORB-SLAM2 / SG-SLAM headers of the hot path use, for builds WITHOUT
// OpenCV (tests in this repo).  Inside the reference tree define SGS_WITH_OPENCV: the real <opencv2/core/core.hpp> types are
// used instead and this file adds nothing.  Layouts match OpenCV (cv::KeyPoint is 28 bytes == sgs_keypoint).
#pragma once
#ifdef SGS_WITH_OPENCV
#include <opencv2/core/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6

namespace cv {

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(T s) { x *= s; y *= s; return *this; }
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// minimal dense matrix: reference-counted buffer, row stride `step` in bytes
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_) : rows(r), cols(c), step(step_ ? step_ : (size_t)c * esz(type)), data((uint8_t*)ext), type_(type) {}
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && data && owner_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
        owner_.reset(new std::vector<uint8_t>((size_t)r * step));
        data = owner_->data();
    }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    Mat row(int r) const { Mat m(1, cols, type_, const_cast<uint8_t*>(data) + (size_t)r * step, step); m.owner_ = owner_; return m; }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; ++r) std::memcpy(m.ptr<uint8_t>(r), ptr<uint8_t>(r), (size_t)cols * esz(type_)); return m; }
private:
    static size_t esz(int type) { return type == CV_8U ? 1 : type == CV_32F ? 4 : 8; }
    int type_ = CV_8U;
    std::shared_ptr<std::vector<uint8_t>> owner_;
};

// cv::InputArray / cv::OutputArray: proxy classes with the member-FUNCTION surface of OpenCV's _InputArray / _OutputArray (getMat(), empty(), type(),
// create(), release()), so that code written against them compiles unchanged against the real headers (where `image.cols` or `image.data` do not exist).
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    int type() const { return m_ ? m_->type() : 0; }
    int rows() const { return m_ ? m_->rows : 0; }
    int cols() const { return m_ ? m_->cols : 0; }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline InputArray noArray() { static _InputArray none; return none; }

}  // namespace cv
#endif

// Minimal stand-in for <opencv2/core/core.hpp>, just enough to compile the reference's own DBoW2 sources
// (Thirdparty/DBoW2/DBoW2/*.cpp, TemplatedVocabulary.h) WITHOUT OpenCV, as the parity pin of the bag-of-words path (oracle/_ref/libdbow2_ref.so;
// recipe in oracle/Makefile).  TEST INFRASTRUCTURE: nothing under sg-slam_b200/ or include/ includes this file.
// DBoW2 uses cv::Mat only as a reference-counted byte row (create / zeros / clone / ptr<T>() / cols / release) and cv::FileStorage only in the
// YAML save/load members, which the pin never calls (it uses saveToTextFile / saveToBinaryFile / loadFrom*): those are declaration-only stubs that abort.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c;
        buf_ = std::make_shared<std::vector<uint8_t>>((size_t)r * c * esz(), (uint8_t)0);
        data = buf_->data();
    }
    void release() { buf_.reset(); rows = cols = 0; data = nullptr; }
    bool empty() const { return !buf_ || rows * cols == 0; }
    Mat clone() const {
        Mat m; m.rows = rows; m.cols = cols; m.type_ = type_;
        if (buf_) { m.buf_ = std::make_shared<std::vector<uint8_t>>(*buf_); m.data = m.buf_->data(); }
        return m;
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    template <class T> T* ptr(int row = 0) { return reinterpret_cast<T*>(buf_->data() + (size_t)row * cols * esz()); }
    template <class T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(buf_->data() + (size_t)row * cols * esz()); }
    Mat row(int r) const {                                  // deep copy of one row (the driver's way to cut descriptors out of a matrix)
        Mat m(1, cols, type_);
        std::memcpy(m.buf_->data(), buf_->data() + (size_t)r * cols * esz(), (size_t)cols * esz());
        return m;
    }
private:
    int type_ = CV_8U;
    size_t esz() const { return type_ == CV_32F ? 4 : 1; }
    std::shared_ptr<std::vector<uint8_t>> buf_;
};

class FileNode {
public:
    FileNode operator[](const char*) const { std::abort(); }
    FileNode operator[](const std::string&) const { std::abort(); }
    FileNode operator[](int) const { std::abort(); }
    size_t size() const { std::abort(); }
    operator int() const { std::abort(); }
    operator float() const { std::abort(); }
    operator double() const { std::abort(); }
    operator std::string() const { std::abort(); }
};

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const char*) const { std::abort(); }
    FileNode operator[](const std::string&) const { std::abort(); }
    void release() {}
};
template <class T> FileStorage& operator<<(FileStorage& fs, const T&) { std::abort(); return fs; }

}  // namespace cv

// What src/Tracking.cc uses of OpenCV beyond the shims of orbmatcher_shim / frame_shim: the settings reader of its constructor (values planted by the driver),
// colour conversion (never executed: the driver builds the frames itself).  TEST INFRASTRUCTURE.
#pragma once
#include <map>
#include <string>
#include <opencv2/opencv.hpp>
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6
#define CV_RGBA2GRAY 11
#define CV_BGRA2GRAY 10
namespace cv {
inline void cvtColor(const Mat&, Mat&, int) {}
struct SVD { enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 }; static void compute(const Mat&, Mat&, Mat&, Mat&, int = 0); };      // declared only (LocalMapping's triangulation: compiled for the drop-in check, never linked)
struct FileNode {
    double v; bool ok;
    operator float() const { return (float)v; }
    operator double() const { return v; }
    operator int() const { return (int)v; }
    bool empty() const { return !ok; }
};
class FileStorage {
public:
    enum { READ = 0 };
    FileStorage(const std::string&, int) {}
    static std::map<std::string, double>& values() { static std::map<std::string, double> m; return m; }
    FileNode operator[](const char* key) const { auto it = values().find(key); return it == values().end() ? FileNode{0.0, false} : FileNode{it->second, true}; }
    bool isOpened() const { return true; }
};
}

// Stand-in for ncnn's net.h: Detector2D.cc compiles against it and `extract("detection_out", out)` hands back the rows the driver planted -- the pin covers the
// reference's handling of the DetectionOutput rows (src/Detector2D.cc:52-88: thresholds, clamping, scaling, the person split and the two flags), NOT ncnn's
// network execution (ncnn is not installable here).  TEST INFRASTRUCTURE.
#pragma once
#include <vector>
namespace ncnn {
struct Option { bool use_vulkan_compute = false; };
class Mat {
public:
    enum { PIXEL_RGB = 1 };
    int w = 0, h = 0, c = 1;
    std::vector<float> d;
    static Mat from_pixels_resize(const unsigned char*, int, int, int, int, int) { return Mat(); }
    void substract_mean_normalize(const float*, const float*) {}
    const float* row(int i) const { return d.data() + (size_t)i * w; }
};
inline Mat*& planted_detection_out() { static Mat* m = nullptr; return m; }
class Extractor {
public:
    void set_light_mode(bool) {}
    int input(const char*, const Mat&) { return 0; }
    int extract(const char*, Mat& out) { if (planted_detection_out()) out = *planted_detection_out(); return 0; }
};
class Net {
public:
    Option opt;
    int load_param(const char*) { return 0; }
    int load_model(const char*) { return 0; }
    Extractor create_extractor() { return Extractor(); }
};
}  // namespace ncnn

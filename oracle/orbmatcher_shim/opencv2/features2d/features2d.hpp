// Stand-in for <opencv2/features2d/features2d.hpp>: cv::KeyPoint only (see ../core/core.hpp).  TEST INFRASTRUCTURE.
#pragma once
#include <cstdint>
#include <vector>
#include "../core/core.hpp"
namespace cv {
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
// cv::FAST = the oracle's restatement (sgo_fast_view: 9-16 segment test + 3x3 non-maximum suppression, pinned against cv2.FastFeatureDetector)
extern "C" int sgo_fast_view(const uint8_t* base, int pitch, int w, int h, int threshold, int nms, int32_t* out, int cap);
inline void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    const Mat im = image_.getMat();
    std::vector<int32_t> out((size_t)std::max(1, im.cols * im.rows) * 3);
    const int n = sgo_fast_view(im.data, (int)im.step, im.cols, im.rows, threshold, nonmaxSuppression ? 1 : 0, out.data(), (int)(out.size() / 3));
    keypoints.clear();
    for (int i = 0; i < n; ++i) { KeyPoint k; k.pt = Point2f((float)out[3 * i], (float)out[3 * i + 1]); k.size = 7.f; k.angle = -1.f; k.response = (float)out[3 * i + 2]; keypoints.push_back(k); }
}
struct KeyPointsFilter {       // only the never-called ComputeKeyPointsOld uses it
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if ((int)k.size() > n) { std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; }); k.resize(n); }
    }
};
}  // namespace cv

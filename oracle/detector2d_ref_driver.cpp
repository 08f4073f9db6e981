// C entry point around the reference's OWN Detector2D::detect (src/Detector2D.cc compiled from the reference tree against the stand-ins of orbmatcher_shim/):
// DetectionOutput rows in ([label, score, x1, y1, x2, y2], normalised corners as ncnn produces them), the reference's accepted objects, person boxes and flags out.
// TEST INFRASTRUCTURE (oracle/_ref/libdetector2d_ref.so).
#include <cstdint>
#include <cstring>

#include "Detector2D.h"

using namespace ORB_SLAM2;

extern "C" __attribute__((visibility("default")))
int ref_detector2d_postprocess(int nrows, const float* rows6, int img_w, int img_h, float det_thr, float dyn_thr, int cap, float* objects_to_view6, int* n_to_view,
                               float* objects6, int* n_objects, float* dyn_map4, int* n_dyn_map, float* dyn_rm4, int* n_dyn_rm, int* have_map, int* have_rm) {
    ncnn::Mat out; out.w = 6; out.h = nrows; out.d.assign(rows6, rows6 + (size_t)nrows * 6);
    ncnn::planted_detection_out() = &out;
    Detector2D det(det_thr, dyn_thr);
    cv::Mat bgr(img_h, img_w, CV_8U);
    det.detect(bgr);
    ncnn::planted_detection_out() = nullptr;
    auto put = [&](const std::vector<Object2D>& v, float* o, int* n) {
        *n = (int)v.size();
        for (int i = 0; i < (int)v.size() && i < cap; ++i) { o[6 * i] = (float)v[i].id; o[6 * i + 1] = v[i].prob; o[6 * i + 2] = v[i].rect.x; o[6 * i + 3] = v[i].rect.y; o[6 * i + 4] = v[i].rect.width; o[6 * i + 5] = v[i].rect.height; }
    };
    auto putr = [&](const std::vector<cv::Rect_<float> >& v, float* o, int* n) {
        *n = (int)v.size();
        for (int i = 0; i < (int)v.size() && i < cap; ++i) { o[4 * i] = v[i].x; o[4 * i + 1] = v[i].y; o[4 * i + 2] = v[i].width; o[4 * i + 3] = v[i].height; }
    };
    put(det.mvObjects2D_to_View, objects_to_view6, n_to_view); put(det.mvObjects2D, objects6, n_objects);
    putr(det.mvPotentialDynamicBorderForMapping, dyn_map4, n_dyn_map); putr(det.mvPotentialDynamicBorderForRmDynamicFeature, dyn_rm4, n_dyn_rm);
    *have_map = det.mbHaveDynamicObjectForMapping; *have_rm = det.mbHaveDynamicObjectForRmDynamicFeature;
    return 0;
}

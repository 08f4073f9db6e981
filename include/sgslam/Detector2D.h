// Detector2D.h -- the detection half of ORB_SLAM2::Detector2D (include/Detector2D.h:40-80, src/Detector2D.cc:16-89) on the GPU.
//
// DetectorGPU keeps the reference's constructor arguments, the `detect(const cv::Mat&)` call and the public result members the rest of the
// pipeline reads (Frame.cc:482-500 copies mbHaveDynamicObjectForRmDynamicFeature / mvPotentialDynamicBorderForRmDynamicFeature, the mapping
// thread reads mvObjects2D / mvPotentialDynamicBorderForMapping, the viewer mvObjects2D_to_View).  In the reference tree, Detector2D keeps its
// thread / handshake code (Run, SetTracker, isNewImageArrived, ImageDetectFinished, draw_objects: src/Detector2D.cc:91-156) and replaces the
// ncnn members by one DetectorGPU: `detect(img)` forwards to it and the result vectors are swapped in.
//
// The model files are the ones the reference loads (./Thirdparty/ncnn_model/mobilenetv3_ssdlite_voc.param/.bin, Detector2D.cc:25-26); the
// library reads the ncnn text graph and weight blob itself, ncnn is not needed.  The image is taken as given (the reference passes the BGR frame
// as PIXEL_RGB, i.e. without a channel swap: Detector2D.cc:39).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../sgs_abi.h"
#include "cv_compat.h"

#ifndef SGSLAM_OBJECT2D_DEFINED          // the reference's own Detector2D.h defines the same struct; define this macro when both are included
typedef struct Object2D {
    cv::Rect_<float> rect;
    float prob;
    std::string name;
    int id;
} Object2D;
#endif

namespace ORB_SLAM2 {

class DetectorGPU {
public:
    DetectorGPU(float detection_confidence_threshold_, float dynamic_detection_confidence_threshold_,
                const std::string& param_path = "./Thirdparty/ncnn_model/mobilenetv3_ssdlite_voc.param",
                const std::string& bin_path = "./Thirdparty/ncnn_model/mobilenetv3_ssdlite_voc.bin", int device = 0)
        : detection_confidence_threshold(detection_confidence_threshold_) {
        if (sgs_detector_create(param_path.c_str(), bin_path.c_str(), 1, detection_confidence_threshold_, dynamic_detection_confidence_threshold_, 0, device, &h_) != SGS_OK)
            throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
        int cap = 0;
        sgs_detector_info(h_, &cap, nullptr, nullptr, nullptr);
        buf_.resize((size_t)cap);
    }
    ~DetectorGPU() { sgs_detector_destroy(h_); }
    DetectorGPU(const DetectorGPU&) = delete;
    DetectorGPU& operator=(const DetectorGPU&) = delete;

    // Detector2D::detect (src/Detector2D.cc:34-89).  Like the reference, mvObjects2D_to_View is appended to (draw_objects clears it, :120).
    template <class MatT>
    void detect(const MatT& bgr) {
        int n = 0;
        if (sgs_detect(h_, bgr.data, bgr.cols, bgr.rows, (int)bgr.step, buf_.data(), (int)buf_.size(), &n) != SGS_OK)
            throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
        mvObjects2D.clear();
        mbHaveDynamicObjectForMapping = false;
        mbHaveDynamicObjectForRmDynamicFeature = false;
        mvPotentialDynamicBorderForRmDynamicFeature.clear();
        mvPotentialDynamicBorderForMapping.clear();
        for (int i = 0; i < n; ++i) {
            Object2D o;
            o.id = buf_[i].id;
            o.name = class_name(o.id);
            o.prob = buf_[i].prob;
            o.rect = cv::Rect_<float>(buf_[i].rect.x, buf_[i].rect.y, buf_[i].rect.w, buf_[i].rect.h);
            mvObjects2D_to_View.push_back(o);
            if (o.id == 15) {                                      // person (:69-80)
                mbHaveDynamicObjectForMapping = true;
                mvPotentialDynamicBorderForMapping.emplace_back(o.rect);
                if (o.prob > 0.2) {
                    mbHaveDynamicObjectForRmDynamicFeature = true;
                    mvPotentialDynamicBorderForRmDynamicFeature.emplace_back(o.rect);
                }
            } else
                mvObjects2D.emplace_back(o);
        }
    }

    static const char* class_name(int id) {                        // Detector2D::class_names (src/Detector2D.cc:8-14): the VOC labels
        static const char* names[] = {"background", "aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
                                      "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor"};
        return id >= 0 && id < 21 ? names[id] : "?";
    }

    std::vector<Object2D> mvObjects2D;
    std::vector<Object2D> mvObjects2D_to_View;
    bool mbHaveDynamicObjectForMapping = false;
    bool mbHaveDynamicObjectForRmDynamicFeature = false;
    std::vector<cv::Rect_<float> > mvPotentialDynamicBorderForMapping;
    std::vector<cv::Rect_<float> > mvPotentialDynamicBorderForRmDynamicFeature;
    float detection_confidence_threshold;

private:
    sgs_detector* h_ = nullptr;
    std::vector<sgs_object2d> buf_;
};

}  // namespace ORB_SLAM2

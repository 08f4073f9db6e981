// ORBextractor.h -- header-compatible mirror of ORB_SLAM2::ORBextractor (/root/reference/src/sg-slam/include/ORBextractor.h:45-105)
// implemented on top of the C ABI of libsgs_cuda.so (include/sgs_abi.h).  Same constructor, operator(), getters and the public
// mvImagePyramid member, so Frame::ExtractORB (src/Frame.cc:274-280) and Tracking (src/Tracking.cc:119-125) compile unchanged.
//
// Differences a maintainer should know:
//   * the extractor is bound to the first image size it sees (device buffers are sized once); a different size re-creates it;
//   * mvImagePyramid is filled only when keepPyramidOnHost(true) was called (it is read by stereo matching only,
//     src/Frame.cc:813,830); the 19-px reflect-101 border of the reference is not materialised (never read for RGB-D/mono);
//   * failures of the GPU path (no device, CUDA error) throw std::runtime_error -- there is no CPU fallback.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../sgs_abi.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_, int device = 0)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_), device_(device) {
        params_.nfeatures = nfeatures_; params_.scale_factor = scaleFactor_; params_.nlevels = nlevels_;
        params_.ini_th_fast = iniThFAST_; params_.min_th_fast = minThFAST_;
        mvImagePyramid.resize(nlevels_);
        // the scale tables do not depend on the image size: compute them exactly as the constructor of the reference does
        // (src/ORBextractor.cc:416-430) so that the getters work before the first frame
        mvScaleFactor.assign(nlevels_, 1.0f); mvLevelSigma2.assign(nlevels_, 1.0f);
        for (int i = 1; i < nlevels_; ++i) {
            mvScaleFactor[i] = (float)((double)mvScaleFactor[i - 1] * scaleFactor);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels_); mvInvLevelSigma2.resize(nlevels_);
        for (int i = 0; i < nlevels_; ++i) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    }
    ~ORBextractor() { if (h_) sgs_extractor_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  The mask is ignored, as in the reference (src/ORBextractor.cc:1045).
    void operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors) {
        if (image.empty()) return;                                                  // :1048
        const cv::Mat im = image.getMat();                                          // :1050 (only member functions of _InputArray / _OutputArray are used: real OpenCV compiles this)
        if (im.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");   // assert at :1051
        ensure(im.cols, im.rows);
        kps_.resize(cap_);
        desc_.resize((size_t)cap_ * 32);
        int n = 0;
        check(sgs_extract(h_, im.data, im.cols, im.rows, (int)im.step, reinterpret_cast<sgs_keypoint*>(kps_.data()), desc_.data(), cap_, &n));
        keypoints.assign(kps_.begin(), kps_.begin() + n);
        if (n == 0) { descriptors.release(); }                                      // :1066-1067
        else {
            descriptors.create(n, 32, CV_8U);                                       // :1069-1070
            cv::Mat d = descriptors.getMat();
            for (int i = 0; i < n; ++i) std::memcpy(d.ptr<uint8_t>(i), desc_.data() + (size_t)i * 32, 32);
        }
        if (keep_pyramid_)
            for (int l = 0; l < nlevels; ++l) {
                int w, h, p;
                check(sgs_extractor_level_info(h_, l, &w, &h, &p));
                mvImagePyramid[l].create(h, w, CV_8U);
                check(sgs_extractor_read_level(h_, 0, l, 0, mvImagePyramid[l].data, (int)mvImagePyramid[l].step));
            }
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
    int inline GetnFeatures() { return nfeatures; }

    void keepPyramidOnHost(bool on) { keep_pyramid_ = on; }
    sgs_extractor* handle() { return h_; }

    std::vector<cv::Mat> mvImagePyramid;

protected:
    void ensure(int w, int h) {
        if (h_ && w == w_ && h == h_img_) return;
        if (h_) { sgs_extractor_destroy(h_); h_ = nullptr; }
        check(sgs_extractor_create(&params_, w, h, 1, device_, &h_));
        check(sgs_extractor_max_keypoints(h_, &cap_));
        w_ = w; h_img_ = h;
    }
    static void check(int status) {
        if (status != SGS_OK) throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    }

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    sgs_orb_params params_{};
    sgs_extractor* h_ = nullptr;
    int device_ = 0, w_ = 0, h_img_ = 0, cap_ = 0;
    bool keep_pyramid_ = false;
    std::vector<cv::KeyPoint> kps_;
    std::vector<uint8_t> desc_;
};

}  // namespace ORB_SLAM2

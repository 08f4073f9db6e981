// FrameDynamic.h -- the geometry half of Frame::RmDynamicPointWithSemanticAndGeometry (src/Frame.cc:430-612) on the GPU.
//
// In the reference tree the body of the "version3" loop (src/Frame.cc:560-604) is replaced by one call:
//
//     Cur_keypoint_sum = ORB_SLAM2::RmDynamicPointsGeometry(mvKeys, mDescriptors, Prepoint, FundMat,
//                            mvPotentialDynamicBorderForRmDynamicFeature, mbHaveDynamicObjectForRmDynamicFeature,
//                            mpORBextractorLeft->GetnFeatures());
//
// and the two OpenCV calls in front of it (src/Frame.cc:445 and :469-472) by
//
//     ORB_SLAM2::CalcOpticalFlowPyrLK(imGray, imGrayPre, Curpoint, Prepoint);                 // cv::calcOpticalFlowPyrLK(..., Size(21,21), 3, {30, 0.01})
//     FundMat = ORB_SLAM2::FindFundamentalMatRansac(CurpointRmDynamic, PrepointRmDynamic);     // cv::findFundamentalMat(..., FM_RANSAC, 1.0, 0.99)
//
// (status / err of LK are not produced: the reference never reads them, quirk Q4.)
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../sgs_abi.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

// keys/descriptors are compacted IN PLACE exactly as the erase loop does (ordered), unless the restore-all guard fires
// (src/Frame.cc:599-602), in which case both stay untouched.  F12: 3x3 CV_64F (empty => keep everything, quirk Q11).
inline int RmDynamicPointsGeometry(std::vector<cv::KeyPoint>& keys, cv::Mat& descriptors, const std::vector<cv::Point2f>& prepoints, const cv::Mat& F12,
                                   const std::vector<cv::Rect_<float> >& dynamicBoxes, bool haveDynamic, int nfeatures, int device = 0) {
    const int n = (int)keys.size();
    if (n == 0) return 0;
    std::vector<float> cur(2 * (size_t)n), prev(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { cur[2 * i] = keys[i].pt.x; cur[2 * i + 1] = keys[i].pt.y; prev[2 * i] = prepoints[i].x; prev[2 * i + 1] = prepoints[i].y; }
    double Fm[9];
    const bool haveF = !F12.empty();
    if (haveF) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Fm[3 * r + c] = F12.at<double>(r, c);
    std::vector<sgs_rect> boxes(dynamicBoxes.size());
    for (size_t b = 0; b < boxes.size(); ++b) { boxes[b].x = dynamicBoxes[b].x; boxes[b].y = dynamicBoxes[b].y; boxes[b].w = dynamicBoxes[b].width; boxes[b].h = dynamicBoxes[b].height; }
    std::vector<uint8_t> keep(n);
    int nkeep = 0, restored = 0;
    const int st = sgs_dynreject(cur.data(), prev.data(), n, haveF ? Fm : nullptr, boxes.empty() ? nullptr : boxes.data(), (int)boxes.size(), haveDynamic ? 1 : 0,
                                 nfeatures, keep.data(), nullptr, &nkeep, &restored, device);
    if (st != SGS_OK) throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    if (!restored) {
        cv::Mat out(nkeep, 32, CV_8U);
        int w = 0;
        for (int i = 0; i < n; ++i)
            if (keep[i]) { keys[w] = keys[i]; std::memcpy(out.ptr<uint8_t>(w), descriptors.ptr<uint8_t>(i), 32); ++w; }
        keys.resize(w);
        descriptors = out;
    }
    return nkeep;
}

// cv::calcOpticalFlowPyrLK(imGray, imGrayPre, Curpoint, Prepoint, State, Err, cv::Size(21, 21), 3, {COUNT|EPS, 30, 0.01}) (src/Frame.cc:445).
// One sgs_lk handle per image size is kept for the life of the process (the reference keeps imGrayPre in a file-scope global the same way).
inline void CalcOpticalFlowPyrLK(const cv::Mat& imGray, const cv::Mat& imGrayPre, const std::vector<cv::Point2f>& curpoints, std::vector<cv::Point2f>& prepoints,
                                 int device = 0) {
    static sgs_lk* lk = nullptr;
    static int lw = 0, lh = 0;
    if (imGray.rows != imGrayPre.rows || imGray.cols != imGrayPre.cols) throw std::runtime_error("sgs: LK images differ in size");
    if (!lk || lw != imGray.cols || lh != imGray.rows) {
        if (lk) sgs_lk_destroy(lk);
        lk = nullptr;
        if (sgs_lk_create(imGray.cols, imGray.rows, 1, device, &lk) != SGS_OK) throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
        lw = imGray.cols; lh = imGray.rows;
    }
    prepoints.resize(curpoints.size());
    if (curpoints.empty()) return;
    static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f layout");
    if (imGray.step != imGrayPre.step) throw std::runtime_error("sgs: LK images differ in row stride");
    if (sgs_lk_track(lk, imGray.ptr<uint8_t>(), imGrayPre.ptr<uint8_t>(), (int)imGray.step, &curpoints[0].x, (int)curpoints.size(), &prepoints[0].x) != SGS_OK)
        throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
}

// cv::findFundamentalMat(points1, points2, cv::FM_RANSAC, 1.0, 0.99) (src/Frame.cc:470,472): 3x3 CV_64F, or an empty Mat when OpenCV would
// return one (no model) -- and also for 7..14 pairs, where OpenCV switches to the plain 7-point / LMedS estimators that are not provided.
inline cv::Mat FindFundamentalMatRansac(const std::vector<cv::Point2f>& points1, const std::vector<cv::Point2f>& points2, double ransacThresh = 1.0,
                                        double confidence = 0.99, int device = 0) {
    cv::Mat F;
    const int n = (int)points1.size();
    if (n < 7 || points2.size() != points1.size()) return F;
    double Fm[9];
    if (sgs_fundamental_ransac(&points1[0].x, &points2[0].x, n, ransacThresh, confidence, 1000, Fm, nullptr, nullptr, device) != SGS_OK)
        throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    if (Fm[0] != Fm[0]) return F;            // NaN == empty matrix
    F.create(3, 3, CV_64F);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F.at<double>(r, c) = Fm[3 * r + c];
    return F;
}

}  // namespace ORB_SLAM2

// FrameDynamic.h -- the geometry half of Frame::RmDynamicPointWithSemanticAndGeometry (src/Frame.cc:430-612) on the GPU.
//
// In the reference tree the body of the "version3" loop (src/Frame.cc:560-604) is replaced by one call:
//
//     Cur_keypoint_sum = ORB_SLAM2::RmDynamicPointsGeometry(mvKeys, mDescriptors, Prepoint, FundMat,
//                            mvPotentialDynamicBorderForRmDynamicFeature, mbHaveDynamicObjectForRmDynamicFeature,
//                            mpORBextractorLeft->GetnFeatures());
//
// calcOpticalFlowPyrLK and findFundamentalMat (src/Frame.cc:445-472) stay where they are (host OpenCV) in this round.
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

}  // namespace ORB_SLAM2

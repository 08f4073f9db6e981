// Optimizer.h -- Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:239-451) on the GPU.
//
// In the reference tree the three call sites (src/Tracking.cc:880, :933, :1314) keep their form:
//
//     Optimizer::PoseOptimization(&mCurrentFrame);      ->      ORB_SLAM2::PoseOptimizationGPU(&mCurrentFrame);
//
// The function reads what the reference reads (mvpMapPoints + GetWorldPos, mvKeysUn, mvuRight, mvInvLevelSigma2, fx..cy, mbf, mTcw), writes
// mvbOutlier and the pose (SetPose) and returns nInitialCorrespondences - nBad.  The optimiser is a restatement of the g2o algorithm the
// reference configures (see sg-slam_b200/csrc/pose_opt.cu); the restatement it is checked against is pinned against the reference's own
// Optimizer.cc + g2o compiled unmodified (tests/test_optimizer_ref.py).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../sgs_abi.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

template <class FrameT>
int PoseOptimizationGPU(FrameT* pFrame, int device = 0) {
    const int n = pFrame->N;
    std::vector<uint8_t> has(n, 0), outlier(n, 0);
    std::vector<float> xyz(3 * (size_t)n, 0.f);
    for (int i = 0; i < n; ++i) {
        auto* pMP = pFrame->mvpMapPoints[i];
        if (!pMP) continue;
        has[i] = 1;
        const cv::Mat Xw = pMP->GetWorldPos();
        for (int k = 0; k < 3; ++k) xyz[3 * (size_t)i + k] = Xw.template at<float>(k, 0);
    }
    sgs_camera cam;
    cam.min_x = FrameT::mnMinX; cam.min_y = FrameT::mnMinY; cam.max_x = FrameT::mnMaxX; cam.max_y = FrameT::mnMaxY;
    cam.fx = FrameT::fx; cam.fy = FrameT::fy; cam.cx = FrameT::cx; cam.cy = FrameT::cy; cam.bf = pFrame->mbf;
    cam.nlevels = (int32_t)pFrame->mvScaleFactors.size();
    float inv_s2[16];
    for (int l = 0; l < 16; ++l) { cam.scale_factors[l] = l < cam.nlevels ? pFrame->mvScaleFactors[l] : 0.f; inv_s2[l] = l < (int)pFrame->mvInvLevelSigma2.size() ? pFrame->mvInvLevelSigma2[l] : 0.f; }
    float Tin[16], Tout[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Tin[4 * r + c] = pFrame->mTcw.template at<float>(r, c);
    int ninliers = 0;
    if (sgs_pose_optimization(&cam, Tin, n, reinterpret_cast<const sgs_keypoint*>(pFrame->mvKeysUn.data()), pFrame->mvuRight.data(), has.data(), xyz.data(), inv_s2, Tout,
                              outlier.data(), &ninliers, device) != SGS_OK)
        throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    for (int i = 0; i < n; ++i) if (has[i]) pFrame->mvbOutlier[i] = outlier[i] != 0;
    cv::Mat pose(4, 4, CV_32F);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pose.template at<float>(r, c) = Tout[4 * r + c];
    pFrame->SetPose(pose);
    return ninliers;
}

}  // namespace ORB_SLAM2

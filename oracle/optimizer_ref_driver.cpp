// C entry point around the reference's OWN Optimizer::PoseOptimization (src/Optimizer.cc:239-451) and the g2o it drives (Thirdparty/g2o: SparseOptimizer,
// OptimizationAlgorithmLevenberg, BlockSolver_6_3, LinearSolverDense, the OnlyPose edges, SE3Quat, the Huber kernel), all compiled unmodified from where they lie
// against the Eigen stand-in of g2o_shim/Eigen and the collaborator stand-ins of g2o_shim/optimizer_standins.h.  TEST INFRASTRUCTURE (oracle/_ref/liboptimizer_ref.so).
#include <cstdint>
#include <cstring>
#include <vector>

#include "Optimizer.h"

using namespace ORB_SLAM2;
#define REF_API extern "C" __attribute__((visibility("default")))

std::mutex MapPoint::mGlobalMutex;

// same argument list as the oracle's sgo_pose_optimization (oracle/pose_opt.cpp); nlevels = length of inv_level_sigma2
REF_API int ref_pose_optimization(const float* Tcw, int n, const uint8_t* has_mp, const float* xyz, const float* kp_xy, const int32_t* octave, const float* uright,
                                  const float* inv_level_sigma2, int nlevels, float fx, float fy, float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier) {
    Frame f;
    f.N = n;
    f.mTcw = cv::Mat(4, 4, CV_32F);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) f.mTcw.at<float>(i, j) = Tcw[4 * i + j];
    f.fx = fx; f.fy = fy; f.cx = cx; f.cy = cy; f.mbf = bf;
    f.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
    f.mvKeysUn.resize(n); f.mvuRight.assign(uright, uright + n); f.mvbOutlier.assign(n, false); f.mvpMapPoints.assign(n, nullptr);
    std::vector<MapPoint> pts(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) {
        f.mvKeysUn[i].pt.x = kp_xy[2 * i]; f.mvKeysUn[i].pt.y = kp_xy[2 * i + 1]; f.mvKeysUn[i].octave = octave[i];
        if (has_mp[i]) {
            pts[i].mWorldPos = cv::Mat(3, 1, CV_32F);
            for (int k = 0; k < 3; ++k) pts[i].mWorldPos.at<float>(k) = xyz[3 * i + k];
            f.mvpMapPoints[i] = &pts[i];
        }
    }
    const int r = Optimizer::PoseOptimization(&f);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Tcw_out[4 * i + j] = f.mTcw.at<float>(i, j);
    for (int i = 0; i < n; ++i) outlier[i] = f.mvbOutlier[i] ? 1 : 0;
    return r;
}

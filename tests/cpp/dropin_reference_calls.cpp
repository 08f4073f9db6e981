// Compiled (not run) by tests/test_dropin_reference.py with the reference's REAL include/Frame.h and include/MapPoint.h: the function mirrors of
// include/sgslam/Optimizer.h, FrameGeometry.h and FrameDynamic.h instantiated on the reference's own classes, written as the call sites would be
// (src/Tracking.cc:880 / :933 / :1314 PoseOptimization, :1262-1312 SearchLocalPoints, src/Frame.cc:445 / :469-472 / :560-604 dynamic-point rejection).
#include "Frame.h"
#include "MapPoint.h"
#include "sgslam/Optimizer.h"
#include "sgslam/FrameGeometry.h"
#include "sgslam/FrameDynamic.h"

using namespace ORB_SLAM2;

int dropin_calls(Frame& cur, std::vector<MapPoint*>& local, const cv::Mat& gray, const cv::Mat& gray_prev, std::vector<cv::Rect_<float> >& boxes) {
    int n = PoseOptimizationGPU(&cur);                                        // Optimizer::PoseOptimization(&mCurrentFrame)
    n += UpdateTrackInView(cur, local, 0.5f);                                 // the isInFrustum loop of SearchLocalPoints
    std::vector<cv::Point2f> curpts, prevpts;
    for (size_t i = 0; i < cur.mvKeys.size(); ++i) curpts.push_back(cur.mvKeys[i].pt);
    CalcOpticalFlowPyrLK(gray, gray_prev, curpts, prevpts);                   // cv::calcOpticalFlowPyrLK
    cv::Mat F12 = FindFundamentalMatRansac(curpts, prevpts);                  // cv::findFundamentalMat(FM_RANSAC, 1.0, 0.99)
    n += RmDynamicPointsGeometry(cur.mvKeys, cur.mDescriptors, prevpts, F12, boxes, cur.mbHaveDynamicObjectForRmDynamicFeature, 1000);
    return n;
}

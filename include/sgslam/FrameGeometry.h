// FrameGeometry.h -- Frame::isInFrustum (src/Frame.cc:296-352) for the whole local map at once, on the GPU.
//
// In the reference tree the loop of Tracking::SearchLocalPoints (src/Tracking.cc:1262-1290)
//
//     for (MapPoint* pMP : mvpLocalMapPoints) { ... if (mCurrentFrame.isInFrustum(pMP, 0.5)) { pMP->IncreaseVisible(); nToMatch++; } }
//
// becomes, for the points that survive the loop's own skips (already matched / bad),
//
//     nToMatch = ORB_SLAM2::UpdateTrackInView(mCurrentFrame, candidates, 0.5f);        // -> sgs_frustum
//
// which fills mbTrackInView / mTrackProjX / mTrackProjY / mTrackProjXR / mnTrackScaleLevel / mTrackViewCos exactly as isInFrustum does
// (the caller keeps the IncreaseVisible() side effect for the points flagged in view).  Templates over the reference's own member
// names (include/Frame.h, include/MapPoint.h); MapPoint must expose GetWorldPos(), GetNormal(), mfMinDistance / mfMaxDistance through
// GetMinDistanceInvariance() / GetMaxDistanceInvariance() (0.8 / 1.2 times the raw values, src/MapPoint.cc:372-382).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../sgs_abi.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

template <class FrameT, class MapPointT>
int UpdateTrackInView(FrameT& F, const std::vector<MapPointT*>& points, float viewingCosLimit, int device = 0) {
    const int n = (int)points.size();
    if (n == 0) return 0;
    std::vector<float> xyz(3 * (size_t)n), nrm(3 * (size_t)n), mn(n), mx(n), px(n), py(n), pxr(n), vc(n);
    std::vector<int32_t> lvl(n);
    std::vector<uint8_t> in(n);
    for (int i = 0; i < n; ++i) {
        const cv::Mat P = points[i]->GetWorldPos(), N = points[i]->GetNormal();
        for (int k = 0; k < 3; ++k) { xyz[3 * i + k] = P.template at<float>(k, 0); nrm[3 * i + k] = N.template at<float>(k, 0); }
        mn[i] = points[i]->GetMinDistanceInvariance() / 0.8f;      // the ABI takes the raw mfMinDistance / mfMaxDistance
        mx[i] = points[i]->GetMaxDistanceInvariance() / 1.2f;
    }
    sgs_camera cam;
    cam.min_x = FrameT::mnMinX; cam.min_y = FrameT::mnMinY; cam.max_x = FrameT::mnMaxX; cam.max_y = FrameT::mnMaxY;
    cam.fx = FrameT::fx; cam.fy = FrameT::fy; cam.cx = FrameT::cx; cam.cy = FrameT::cy; cam.bf = F.mbf;
    cam.nlevels = (int32_t)F.mvScaleFactors.size();
    for (int l = 0; l < 16; ++l) cam.scale_factors[l] = l < cam.nlevels ? F.mvScaleFactors[l] : 0.f;
    float T[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[4 * r + c] = F.mTcw.template at<float>(r, c);
    if (sgs_frustum(&cam, T, n, xyz.data(), nrm.data(), mn.data(), mx.data(), viewingCosLimit, in.data(), px.data(), py.data(), pxr.data(), lvl.data(),
                    vc.data(), device) != SGS_OK)
        throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        MapPointT* p = points[i];
        p->mbTrackInView = in[i] != 0;
        if (in[i]) { p->mTrackProjX = px[i]; p->mTrackProjXR = pxr[i]; p->mTrackProjY = py[i]; p->mnTrackScaleLevel = lvl[i]; p->mTrackViewCos = vc[i]; ++cnt; }
    }
    return cnt;
}

}  // namespace ORB_SLAM2

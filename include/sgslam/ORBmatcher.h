// ORBmatcher.h -- mirror of the tracking-side part of ORB_SLAM2::ORBmatcher
// (/root/reference/src/sg-slam/include/ORBmatcher.h:36-101) on top of the C ABI of libsgs_cuda.so.
//
// The reference methods take object graphs (Frame&, MapPoint*).  This header keeps the same method names, argument order and
// return values, as templates over the Frame / MapPoint types: inside the reference tree they instantiate with
// ORB_SLAM2::Frame / ORB_SLAM2::MapPoint unchanged (the member names used below are the reference's own,
// include/Frame.h and include/MapPoint.h); in this repository's tests they instantiate with small structs of the same shape.
// Each method FLATTENS the graph to the arrays the C ABI wants, calls the GPU, and writes mvpMapPoints back.
//
// Covered here (tracking thread, every frame): SearchByProjection(Frame&, const Frame&, th, bMono)   src/ORBmatcher.cc:1332-1472
//                                              SearchByProjection(Frame&, vector<MapPoint*>&, th)     src/ORBmatcher.cc:45-129
//                                              DescriptorDistance                                      src/ORBmatcher.cc:1649-1665
//                                              SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:1474-1601
//                                              SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)      src/ORBmatcher.cc:159-290
// The mapping / loop-closing variants (SearchByBoW, Fuse, SearchBySim3, ...) are SURVEY section 8(f) "next" rows.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../sgs_abi.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), device_(device) {}

    // Hamming distance of two 256-bit descriptors: a pure scalar helper the rest of the reference calls on single pairs
    // (e.g. MapPoint::ComputeDistinctiveDescriptors); batches go through sgs_hamming_bf / sgs_hamming_pairs.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        const uint32_t* pa = a.ptr<uint32_t>(); const uint32_t* pb = b.ptr<uint32_t>();
        int dist = 0;
        for (int i = 0; i < 8; ++i) dist += __builtin_popcount(pa[i] ^ pb[i]);
        return dist;
    }

    // Project MapPoints tracked in last frame into the current frame and search matches (Tracking::TrackWithMotionModel).
    template <class FrameT>
    int SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th, const bool bMono) {
        const int n = CurrentFrame.N, nl = LastFrame.N;
        std::vector<float> scale(CurrentFrame.mvScaleFactors.begin(), CurrentFrame.mvScaleFactors.end());
        sgs_frame_view cur = view(CurrentFrame, scale);
        std::vector<uint8_t> has(nl, 0), obs(nl, 0), ldesc((size_t)nl * 32, 0);
        std::vector<float> xyz((size_t)nl * 3, 0.f), ang(nl, 0.f);
        std::vector<int32_t> oct(nl, 0), mp(n, -1);
        std::vector<uint8_t> mp_obs(n, 0);
        for (int i = 0; i < nl; ++i) {
            auto* pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            has[i] = 1; obs[i] = pMP->Observations() > 0 ? 1 : 0;
            const cv::Mat x = pMP->GetWorldPos(); const cv::Mat d = pMP->GetDescriptor();
            for (int k = 0; k < 3; ++k) xyz[3 * (size_t)i + k] = x.template at<float>(k, 0);
            std::memcpy(&ldesc[(size_t)i * 32], d.template ptr<uint8_t>(), 32);
            oct[i] = LastFrame.mvKeys[i].octave; ang[i] = LastFrame.mvKeysUn[i].angle;
        }
        // entries the caller left in CurrentFrame.mvpMapPoints (Tracking.cc:916 clears them before this call) keep their pointer;
        // they are encoded as nl + j so that new matches (0..nl-1) can be told apart when writing back
        for (int j = 0; j < n; ++j)
            if (CurrentFrame.mvpMapPoints[j]) { mp[j] = nl + j; mp_obs[j] = CurrentFrame.mvpMapPoints[j]->Observations() > 0 ? 1 : 0; }
        float tc[16], tl[16];
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { tc[4 * r + c] = CurrentFrame.mTcw.template at<float>(r, c); tl[4 * r + c] = LastFrame.mTcw.template at<float>(r, c); }
        int nmatches = 0;
        check(sgs_match_project_lastframe(&cur, tc, tl, nl, has.data(), xyz.data(), ldesc.data(), obs.data(), oct.data(), ang.data(), th, bMono ? 1 : 0,
                                          mbCheckOrientation ? 1 : 0, mp.data(), mp_obs.data(), &nmatches, device_));
        for (int j = 0; j < n; ++j) {
            if (mp[j] < 0) CurrentFrame.mvpMapPoints[j] = nullptr;
            else if (mp[j] < nl) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[mp[j]];
        }
        return nmatches;
    }

    // Search matches between Frame keypoints and projected MapPoints (Tracking::SearchLocalPoints).  The per-point fields are the
    // ones Frame::isInFrustum stored in the MapPoints (src/Frame.cc:296-352).
    template <class FrameT, class MapPointT>
    int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th = 3) {
        const int n = F.N, nmp = (int)vpMapPoints.size();
        std::vector<float> scale(F.mvScaleFactors.begin(), F.mvScaleFactors.end());
        sgs_frame_view fv = view(F, scale);
        std::vector<uint8_t> inview(nmp, 0), obs(nmp, 0), desc((size_t)nmp * 32, 0), fobs(n, 0);
        std::vector<float> px(nmp, 0.f), py(nmp, 0.f), pxr(nmp, 0.f), vc(nmp, 0.f);
        std::vector<int32_t> lvl(nmp, 0), fmp(n, -1);
        for (int i = 0; i < nmp; ++i) {
            MapPointT* pMP = vpMapPoints[i];
            if (!pMP->mbTrackInView || pMP->isBad()) continue;
            inview[i] = 1; obs[i] = pMP->Observations() > 0 ? 1 : 0;
            px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; pxr[i] = pMP->mTrackProjXR; lvl[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos;
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(&desc[(size_t)i * 32], d.template ptr<uint8_t>(), 32);
        }
        for (int j = 0; j < n; ++j)
            if (F.mvpMapPoints[j]) { fmp[j] = nmp + j; fobs[j] = F.mvpMapPoints[j]->Observations() > 0 ? 1 : 0; }
        int nmatches = 0;
        check(sgs_match_project_localmap(&fv, nmp, inview.data(), px.data(), py.data(), pxr.data(), lvl.data(), vc.data(), desc.data(), obs.data(), th,
                                         mfNNratio, 0, fmp.data(), fobs.data(), &nmatches, device_));
        for (int j = 0; j < n; ++j)
            if (fmp[j] >= 0 && fmp[j] < nmp) F.mvpMapPoints[j] = vpMapPoints[fmp[j]];
        return nmatches;
    }

    // Project MapPoints seen in KeyFrame into the Frame and search matches (Tracking::Relocalization, src/Tracking.cc:1494,1508).
    template <class FrameT, class KeyFrameT, class SetT>
    int SearchByProjection(FrameT& CurrentFrame, KeyFrameT* pKF, const SetT& sAlreadyFound, const float th, const int ORBdist) {
        const int n = CurrentFrame.N;
        const auto vpMPs = pKF->GetMapPointMatches();
        const int nkf = (int)vpMPs.size();
        std::vector<float> scale(CurrentFrame.mvScaleFactors.begin(), CurrentFrame.mvScaleFactors.end());
        sgs_frame_view fv = view(CurrentFrame, scale);
        std::vector<uint8_t> valid(nkf, 0), desc((size_t)nkf * 32, 0);
        std::vector<float> xyz((size_t)nkf * 3, 0.f), ang(nkf, 0.f), mn(nkf, 0.f), mx(nkf, 0.f);
        for (int i = 0; i < nkf; ++i) {
            auto* pMP = vpMPs[i];
            if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
            valid[i] = 1;
            const cv::Mat P = pMP->GetWorldPos(), d = pMP->GetDescriptor();
            for (int k = 0; k < 3; ++k) xyz[3 * (size_t)i + k] = P.template at<float>(k, 0);
            std::memcpy(&desc[(size_t)i * 32], d.template ptr<uint8_t>(), 32);
            ang[i] = pKF->mvKeysUn[i].angle;
            mn[i] = pMP->GetMinDistanceInvariance() / 0.8f; mx[i] = pMP->GetMaxDistanceInvariance() / 1.2f;      // raw mfMinDistance / mfMaxDistance
        }
        std::vector<int32_t> cmp(n, -1);
        for (int j = 0; j < n; ++j) if (CurrentFrame.mvpMapPoints[j]) cmp[j] = nkf + j;
        float T[16];
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[4 * r + c] = CurrentFrame.mTcw.template at<float>(r, c);
        int nmatches = 0;
        check(sgs_match_project_keyframe(&fv, T, nkf, valid.data(), xyz.data(), desc.data(), ang.data(), mn.data(), mx.data(), th, ORBdist,
                                         mbCheckOrientation ? 1 : 0, cmp.data(), &nmatches, device_));
        for (int j = 0; j < n; ++j)
            if (cmp[j] >= 0 && cmp[j] < nkf) CurrentFrame.mvpMapPoints[j] = vpMPs[cmp[j]];
        return nmatches;
    }

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame, brute force constrained to ORB that belong to the same vocabulary
    // node (Tracking::TrackReferenceKeyFrame src/Tracking.cc:865, Relocalization :1455).  mFeatVec: DBoW2::FeatureVector of both sides.
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches) {
        const auto vpMapPointsKF = pKF->GetMapPointMatches();
        const int nkf = (int)vpMapPointsKF.size(), nf = F.N;
        vpMapPointMatches = std::vector<MapPointT*>(nf, static_cast<MapPointT*>(NULL));
        if (nkf == 0 || nf == 0) return 0;
        std::vector<int32_t> kn(nkf, 0), fn(nf, 0), m(nf, -1);
        std::vector<double> kw(nkf, 0.0), fw(nf, 0.0);             // weight > 0 <=> the feature is listed in the FeatureVector
        std::vector<uint8_t> kv(nkf, 0);
        std::vector<float> ka(nkf, 0.f), fa(nf, 0.f);
        for (const auto& kvp : pKF->mFeatVec) for (unsigned idx : kvp.second) { kn[idx] = (int32_t)kvp.first; kw[idx] = 1.0; }
        for (const auto& kvp : F.mFeatVec) for (unsigned idx : kvp.second) { fn[idx] = (int32_t)kvp.first; fw[idx] = 1.0; }
        for (int i = 0; i < nkf; ++i) { kv[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) ? 1 : 0; ka[i] = pKF->mvKeysUn[i].angle; }
        for (int j = 0; j < nf; ++j) fa[j] = F.mvKeysUn[j].angle;
        int nmatches = 0;
        check(sgs_match_bow(nkf, kn.data(), kw.data(), kv.data(), pKF->mDescriptors.template ptr<uint8_t>(), ka.data(), nf, fn.data(), fw.data(),
                            F.mDescriptors.template ptr<uint8_t>(), fa.data(), mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches, device_));
        for (int j = 0; j < nf; ++j) if (m[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[m[j]];
        return nmatches;
    }

    static const int TH_LOW = 50;      // src/ORBmatcher.cc:37-39
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    template <class FrameT>
    static sgs_frame_view view(const FrameT& F, const std::vector<float>& scale) {
        sgs_frame_view v;
        v.n = F.N;
        v.keys_un = reinterpret_cast<const sgs_keypoint*>(F.mvKeysUn.data());
        v.u_right = F.mvuRight.data();
        v.desc = F.mDescriptors.template ptr<uint8_t>();     // N x 32, continuous (cv::Mat::create)
        v.min_x = FrameT::mnMinX; v.min_y = FrameT::mnMinY; v.max_x = FrameT::mnMaxX; v.max_y = FrameT::mnMaxY;
        v.fx = FrameT::fx; v.fy = FrameT::fy; v.cx = FrameT::cx; v.cy = FrameT::cy; v.bf = F.mbf;
        v.nlevels = (int)scale.size(); v.scale_factors = scale.data();
        return v;
    }
    static void check(int status) {
        if (status != SGS_OK) throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    }

    float mfNNratio;
    bool mbCheckOrientation;
    int device_;
};

}  // namespace ORB_SLAM2

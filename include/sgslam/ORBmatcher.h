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
// Mapping / loop closing:                      Fuse(KeyFrame*, vpMapPoints, th)                       src/ORBmatcher.cc:829-980
//                                              Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)     src/ORBmatcher.cc:982-1104
//                                              SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  src/ORBmatcher.cc:292-405
//                                              SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)  src/ORBmatcher.cc:1106-1330
//                                              SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)         src/ORBmatcher.cc:524-657
//                                              SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  src/ORBmatcher.cc:659-827
//                                              SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  src/ORBmatcher.cc:407-522
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
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
            mn[i] = raw_distance(pMP->GetMinDistanceInvariance(), 0.8f); mx[i] = raw_distance(pMP->GetMaxDistanceInvariance(), 1.2f);      // mfMinDistance / mfMaxDistance
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

    // Matching for the Map Initialization (only used in the monocular case, Tracking::MonocularInitialization src/Tracking.cc:660-661).  src/ORBmatcher.cc:407-522
    template <class FrameT, class PointT>
    int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<PointT>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) {
        const int n1 = (int)F1.mvKeysUn.size();
        vnMatches12 = std::vector<int>(n1, -1);
        std::vector<float> scale(F2.mvScaleFactors.begin(), F2.mvScaleFactors.end());
        sgs_frame_view v1 = view(F1, scale), v2 = view(F2, scale);
        v1.n = n1; v2.n = (int)F2.mvKeysUn.size();
        std::vector<float> prev(2 * (size_t)n1);
        for (int i = 0; i < n1; ++i) { prev[2 * (size_t)i] = vbPrevMatched[i].x; prev[2 * (size_t)i + 1] = vbPrevMatched[i].y; }
        std::vector<int32_t> m(n1, -1);
        int nmatches = 0;
        check(sgs_search_for_initialization(&v1, &v2, prev.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches, device_));
        for (int i = 0; i < n1; ++i) {
            vnMatches12[i] = m[i];
            if (m[i] >= 0) { vbPrevMatched[i].x = prev[2 * (size_t)i]; vbPrevMatched[i].y = prev[2 * (size_t)i + 1]; }
        }
        return nmatches;
    }

    // Matching between two key frames' map points through the vocabulary tree (LoopClosing::ComputeSim3, src/LoopClosing.cc:272).  src/ORBmatcher.cc:524-657
    template <class KeyFrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12) {
        const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
        vpMatches12 = std::vector<MapPointT*>(n1, static_cast<MapPointT*>(NULL));
        if (n1 == 0 || n2 == 0) return 0;
        KfBow a = flatten_bow(pKF1, n1), b = flatten_bow(pKF2, n2);
        for (int i = 0; i < n1; ++i) a.valid[i] = (vpMapPoints1[i] && !vpMapPoints1[i]->isBad()) ? 1 : 0;
        for (int i = 0; i < n2; ++i) b.valid[i] = (vpMapPoints2[i] && !vpMapPoints2[i]->isBad()) ? 1 : 0;
        std::vector<int32_t> m(n1, -1);
        int nmatches = 0;
        check(sgs_match_bow_keyframes(1, n1, a.node.data(), a.weight.data(), a.valid.data(), pKF1->mDescriptors.template ptr<uint8_t>(), a.angle.data(), n2, b.node.data(),
                                      b.weight.data(), b.valid.data(), pKF2->mDescriptors.template ptr<uint8_t>(), b.angle.data(), mfNNratio, mbCheckOrientation ? 1 : 0, nullptr,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, m.data(), &nmatches, device_));
        for (int i = 0; i < n1; ++i) if (m[i] >= 0) vpMatches12[i] = vpMapPoints2[m[i]];
        return nmatches;
    }

    // Matching to triangulate new MapPoints, checking the epipolar constraint (LocalMapping::CreateNewMapPoints, src/LocalMapping.cc:298).  src/ORBmatcher.cc:659-827
    template <class KeyFrameT>
    int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, const cv::Mat& F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
        const int n1 = pKF1->N, n2 = pKF2->N;
        vMatchedPairs.clear();
        if (n1 == 0 || n2 == 0) return 0;
        // epipole of camera 1 in image 2 (:667-672): C2 = R2w*Cw + t2w (small-matrix product: float sum, the addition in double), invz = 1.0f / z
        const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
        float C2[3];
        for (int r = 0; r < 3; ++r) {
            volatile float acc = R2w.template at<float>(r, 0) * Cw.template at<float>(0, 0);
            acc = acc + R2w.template at<float>(r, 1) * Cw.template at<float>(1, 0);
            acc = acc + R2w.template at<float>(r, 2) * Cw.template at<float>(2, 0);
            C2[r] = (float)((double)acc + (double)t2w.template at<float>(r, 0));
        }
        const float invz = 1.0f / C2[2];
        volatile float exv = pKF2->fx * C2[0]; exv = exv * invz; exv = exv + pKF2->cx;
        volatile float eyv = pKF2->fy * C2[1]; eyv = eyv * invz; eyv = eyv + pKF2->cy;
        const float epi[2] = {exv, eyv};
        KfBow a = flatten_bow(pKF1, n1), b = flatten_bow(pKF2, n2);
        std::vector<uint8_t> st1(n1), st2(n2);
        std::vector<float> xy1(2 * (size_t)n1), xy2(2 * (size_t)n2), F(9);
        std::vector<int32_t> oct2(n2);
        for (int i = 0; i < n1; ++i) { a.valid[i] = pKF1->GetMapPoint(i) ? 0 : 1; st1[i] = pKF1->mvuRight[i] >= 0 ? 1 : 0; xy1[2 * (size_t)i] = pKF1->mvKeysUn[i].pt.x; xy1[2 * (size_t)i + 1] = pKF1->mvKeysUn[i].pt.y; }
        for (int i = 0; i < n2; ++i) { b.valid[i] = pKF2->GetMapPoint(i) ? 0 : 1; st2[i] = pKF2->mvuRight[i] >= 0 ? 1 : 0; xy2[2 * (size_t)i] = pKF2->mvKeysUn[i].pt.x; xy2[2 * (size_t)i + 1] = pKF2->mvKeysUn[i].pt.y; oct2[i] = pKF2->mvKeysUn[i].octave; }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F[3 * r + c] = F12.template at<float>(r, c);
        std::vector<float> s2(pKF2->mvLevelSigma2.begin(), pKF2->mvLevelSigma2.end()), sf(pKF2->mvScaleFactors.begin(), pKF2->mvScaleFactors.end());
        std::vector<int32_t> m(n1, -1);
        int nmatches = 0;
        check(sgs_match_bow_keyframes(2, n1, a.node.data(), a.weight.data(), a.valid.data(), pKF1->mDescriptors.template ptr<uint8_t>(), a.angle.data(), n2, b.node.data(),
                                      b.weight.data(), b.valid.data(), pKF2->mDescriptors.template ptr<uint8_t>(), b.angle.data(), mfNNratio, mbCheckOrientation ? 1 : 0, st1.data(),
                                      st2.data(), xy1.data(), xy2.data(), oct2.data(), F.data(), epi, s2.data(), sf.data(), (int)sf.size(), bOnlyStereo ? 1 : 0, m.data(), &nmatches,
                                      device_));
        vMatchedPairs.reserve(nmatches);
        for (int i = 0; i < n1; ++i) if (m[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m[i]));
        return nmatches;
    }

    static const int TH_LOW = 50;      // src/ORBmatcher.cc:37-39
    // ---- mapping / loop-closing variants.  The search runs on the GPU; the reference's map side effects (Replace, AddObservation, AddMapPoint,
    // vpReplacePoint / vpMatched writes) are applied afterwards on the host, in the reference's order and with its own checks re-evaluated at that
    // moment (the search of one point reads only immutable key-frame data and that point's own fields, so it does not depend on them).
    // MapPoint keeps mfMinDistance / mfMaxDistance private: they are recovered from Get{Min,Max}DistanceInvariance() (0.8 f / 1.2 f factors).

    // Project MapPoints into KeyFrame and search for duplicated MapPoints (LocalMapping::SearchInNeighbors, src/LocalMapping.cc:215).  src/ORBmatcher.cc:829-980
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, const float th = 3.0) {
        const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
        float T[16] = {0}, O[3];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = Rcw.template at<float>(r, c); T[4 * r + 3] = tcw.template at<float>(r, 0); O[r] = Ow.template at<float>(r, 0); }
        T[15] = 1.f;
        std::vector<int32_t> bi, bd;
        search_kf(pKF, T, O, vpMapPoints, [&](MapPointT* p) { return p && !p->isBad() && !p->IsInKeyFrame(pKF); }, th, 0, bi, bd, nullptr);
        int nFused = 0;
        for (size_t i = 0; i < vpMapPoints.size(); ++i) {
            MapPointT* pMP = vpMapPoints[i];
            if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF) || bd[i] > TH_LOW) continue;        // :848-853 re-evaluated in order, :963
            MapPointT* pMPinKF = pKF->GetMapPoint(bi[i]);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, bi[i]);
                pKF->AddMapPoint(pMP, bi[i]);
            }
            nFused++;
        }
        return nFused;
    }

    // Project MapPoints into KeyFrame using a given Sim3 and search for duplicated MapPoints (LoopClosing::SearchAndFuse, src/LoopClosing.cc:589).  :982-1104
    // Rcw / tcw / Ow are decomposed from Scw with the expressions of :988-992 evaluated by the caller's own cv::Mat (decompose_scw below).
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, const cv::Mat& Scw, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint) {
        float T[16], O[3];
        decompose_scw(Scw, T, O);
        return Fuse(pKF, T, O, vpPoints, th, vpReplacePoint);
    }
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, const float T[16], const float O[3], const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint) {
        const auto spAlreadyFound = pKF->GetMapPoints();
        std::vector<int32_t> bi, bd;
        search_kf(pKF, T, O, vpPoints, [&](MapPointT* p) { return !p->isBad() && !spAlreadyFound.count(p); }, th, 1, bi, bd, nullptr);
        int nFused = 0;
        for (size_t i = 0; i < vpPoints.size(); ++i) {
            MapPointT* pMP = vpPoints[i];
            if (pMP->isBad() || spAlreadyFound.count(pMP) || bd[i] > TH_LOW) continue;
            MapPointT* pMPinKF = pKF->GetMapPoint(bi[i]);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
            } else {
                pMP->AddObservation(pKF, bi[i]);
                pKF->AddMapPoint(pMP, bi[i]);
            }
            nFused++;
        }
        return nFused;
    }

    // Project MapPoints using a Similarity Transformation and search matches (LoopClosing::ComputeSim3 / DetectLoop, src/LoopClosing.cc:239,:589).  :292-405
    // The points are matched in order: a key-frame feature taken by an earlier point is not offered to later ones.
    template <class KeyFrameT, class MapPointT>
    int SearchByProjection(KeyFrameT* pKF, const cv::Mat& Scw, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th) {
        float T[16], O[3];
        decompose_scw(Scw, T, O);
        return SearchByProjection(pKF, T, O, vpPoints, vpMatched, th);
    }
    template <class KeyFrameT, class MapPointT>
    int SearchByProjection(KeyFrameT* pKF, const float T[16], const float O[3], const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th) {
        std::vector<int32_t> occupied(vpMatched.size(), -1);
        for (size_t j = 0; j < vpMatched.size(); ++j) if (vpMatched[j]) occupied[j] = (int32_t)vpPoints.size();       // any id >= 0: taken on entry
        auto found = [&](MapPointT* p) { for (MapPointT* q : vpMatched) if (q == p) return true; return false; };     // spAlreadyFound (:308-309)
        std::vector<int32_t> bi, bd;
        const int nmatches = search_kf(pKF, T, O, vpPoints, [&](MapPointT* p) { return !p->isBad() && !found(p); }, (float)th, 3, bi, bd, &occupied);
        for (size_t j = 0; j < vpMatched.size(); ++j)
            if (occupied[j] >= 0 && occupied[j] < (int32_t)vpPoints.size()) vpMatched[j] = vpPoints[occupied[j]];
        return nmatches;
    }

    // Search matches between MapPoints seen in KF1 and KF2 transforming by a Sim3 [s12*R12|t12] (LoopClosing::ComputeSim3, src/LoopClosing.cc:313).  :1106-1330
    // Both projection searches run on the GPU; vbAlreadyMatched, the <= TH_HIGH decision and the agreement check are the reference's, on the host.
    // sR12 = s12*R12, sR21 = (1.0/s12)*R12.t(), t21 = -sR21*t12 (:1122-1124) in scalar float arithmetic (Mat*scalar = multiply by the float value of the
    // factor -- restated, see decompose_scw; the small-matrix product sums float products left to right).
    template <class KeyFrameT, class MapPointT>
    int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, const float th) {
        float x21[12], x12[12];                                   // [sR21 | t21], [sR12 | t12]
        const float inv = (float)(1.0 / (double)s12);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) { x12[3 * r + c] = R12.at<float>(r, c) * s12; x21[3 * r + c] = R12.at<float>(c, r) * inv; }
            x12[9 + r] = t12.at<float>(r, 0);
        }
        for (int r = 0; r < 3; ++r) {
            volatile float acc = x21[3 * r] * x12[9];
            acc = acc + x21[3 * r + 1] * x12[10];
            acc = acc + x21[3 * r + 2] * x12[11];
            x21[9 + r] = -acc;
        }
        const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
        std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
        for (int i = 0; i < N1; ++i) {
            MapPointT* pMP = vpMatches12[i];
            if (pMP) {
                vbAlreadyMatched1[i] = true;
                const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
                if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
            }
        }
        auto pose = [](KeyFrameT* kf, float T[16]) {
            const cv::Mat R = kf->GetRotation(), t = kf->GetTranslation();
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = R.template at<float>(r, c); T[4 * r + 3] = t.template at<float>(r, 0); }
            T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
        };
        float T1[16], T2[16];
        pose(pKF1, T1); pose(pKF2, T2);
        const float zero[3] = {0.f, 0.f, 0.f};
        std::vector<int32_t> bi1, bd1, bi2, bd2;
        // the skip test depends on the feature index, not only on the point: flatten with explicit masks
        auto masked = [&](const std::vector<MapPointT*>& pts, const std::vector<bool>& already) {
            std::vector<MapPointT*> v(pts.size(), nullptr);
            for (size_t i = 0; i < pts.size(); ++i) if (pts[i] && !already[i] && !pts[i]->isBad()) v[i] = pts[i];
            return v;
        };
        const std::vector<MapPointT*> c1 = masked(vpMapPoints1, vbAlreadyMatched1), c2 = masked(vpMapPoints2, vbAlreadyMatched2);
        search_kf(pKF2, T1, zero, c1, [](MapPointT* p) { return p != nullptr; }, th, 2, bi1, bd1, nullptr, x21);      // KF1's points into KF2 (:1148-1225)
        search_kf(pKF1, T2, zero, c2, [](MapPointT* p) { return p != nullptr; }, th, 2, bi2, bd2, nullptr, x12);      // KF2's points into KF1 (:1227-1307)
        int nFound = 0;
        for (int i1 = 0; i1 < N1; ++i1) {
            const int idx2 = bd1[i1] <= TH_HIGH ? bi1[i1] : -1;
            if (idx2 >= 0) {
                const int idx1 = bd2[idx2] <= TH_HIGH ? bi2[idx2] : -1;
                if (idx1 == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
            }
        }
        return nFound;
    }

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
    struct KfBow { std::vector<int32_t> node; std::vector<double> weight; std::vector<uint8_t> valid; std::vector<float> angle; };
    // per-feature view of a key frame's DBoW2::FeatureVector: node id of the feature, weight 1 = the feature is listed (0 = not in mFeatVec)
    template <class KeyFrameT>
    static KfBow flatten_bow(KeyFrameT* pKF, int n) {
        KfBow b;
        b.node.assign(n, 0); b.weight.assign(n, 0.0); b.valid.assign(n, 0); b.angle.assign(n, 0.f);
        for (const auto& kvp : pKF->mFeatVec) for (unsigned idx : kvp.second) { b.node[idx] = (int32_t)kvp.first; b.weight[idx] = 1.0; }
        for (int i = 0; i < n; ++i) b.angle[i] = pKF->mvKeysUn[i].angle;
        return b;
    }

    // sRcw = Scw(0:3,0:3); scw = sqrt(sRcw.row(0).dot(sRcw.row(0))); Rcw = sRcw/scw; tcw = Scw(0:3,3)/scw; Ow = -Rcw.t()*tcw   (src/ORBmatcher.cc:301-305, :988-992)
    // with scalar float operations in OpenCV's evaluation order: Mat::dot accumulates in double; Mat/scalar is convertTo(alpha = 1/s), which for CV_32F
    // multiplies by the FLOAT value of the reciprocal; the transposed product accumulates in double (pinned with cv2.gemm).  The
    // Mat/scalar rule is restated from OpenCV's sources and could not be checked here (cv2 has no MatExpr): inside the reference tree prefer the
    // overloads below that take Rcw|tcw (4x4 row major) and Ow computed by the reference's own three lines.
    static void decompose_scw(const cv::Mat& Scw, float T[16], float O[3]) {
        double dot = 0;
        for (int c = 0; c < 3; ++c) dot += (double)Scw.at<float>(0, c) * (double)Scw.at<float>(0, c);
        const float scw = std::sqrt((float)dot);
        const float inv = (float)(1.0 / (double)scw);
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = Scw.at<float>(r, c) * inv; T[4 * r + 3] = Scw.at<float>(r, 3) * inv; }
        T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
        for (int r = 0; r < 3; ++r) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += (double)T[4 * k + r] * (double)T[4 * k + 3];
            O[r] = (float)(-acc);
        }
    }

    // Flattens the key frame and the candidate points and runs sgs_fuse_search.  valid(p) = the reference's per-point skip test.
    template <class KeyFrameT, class MapPointT, class ValidF>
    int search_kf(KeyFrameT* pKF, const float T[16], const float O[3], const std::vector<MapPointT*>& pts, ValidF valid, float th, int variant,
                  std::vector<int32_t>& bi, std::vector<int32_t>& bd, std::vector<int32_t>* kf_matched, const float* xform2 = nullptr) {
        const int nmp = (int)pts.size();
        std::vector<float> scale(pKF->mvScaleFactors.begin(), pKF->mvScaleFactors.end()), inv_s2(pKF->mvInvLevelSigma2.begin(), pKF->mvInvLevelSigma2.end());
        sgs_frame_view v;
        v.n = pKF->N;
        v.keys_un = reinterpret_cast<const sgs_keypoint*>(pKF->mvKeysUn.data());
        v.u_right = pKF->mvuRight.data();
        v.desc = pKF->mDescriptors.template ptr<uint8_t>();
        v.min_x = pKF->mnMinX; v.min_y = pKF->mnMinY; v.max_x = pKF->mnMaxX; v.max_y = pKF->mnMaxY;
        v.fx = pKF->fx; v.fy = pKF->fy; v.cx = pKF->cx; v.cy = pKF->cy; v.bf = pKF->mbf;
        v.nlevels = (int)scale.size(); v.scale_factors = scale.data();
        std::vector<uint8_t> ok(nmp, 0), desc((size_t)nmp * 32, 0);
        std::vector<float> xyz((size_t)nmp * 3, 0.f), nrm((size_t)nmp * 3, 0.f), mn(nmp, 0.f), mx(nmp, 0.f);
        for (int i = 0; i < nmp; ++i) {
            MapPointT* p = pts[i];
            if (!valid(p)) continue;
            ok[i] = 1;
            const cv::Mat P = p->GetWorldPos(), d = p->GetDescriptor();
            for (int k = 0; k < 3; ++k) xyz[3 * (size_t)i + k] = P.template at<float>(k, 0);
            if (variant != 2) { const cv::Mat Nn = p->GetNormal(); for (int k = 0; k < 3; ++k) nrm[3 * (size_t)i + k] = Nn.template at<float>(k, 0); }
            std::memcpy(&desc[(size_t)i * 32], d.template ptr<uint8_t>(), 32);
            mn[i] = raw_distance(p->GetMinDistanceInvariance(), 0.8f); mx[i] = raw_distance(p->GetMaxDistanceInvariance(), 1.2f);
        }
        bi.assign(nmp, -1); bd.assign(nmp, 256);
        int nmatches = 0;
        check(sgs_fuse_search(&v, T, O, nmp, ok.data(), xyz.data(), nrm.data(), mn.data(), mx.data(), desc.data(), th, inv_s2.data(), variant, xform2, bi.data(),
                              bd.data(), kf_matched ? kf_matched->data() : nullptr, &nmatches, device_));
        return nmatches;
    }

    static void check(int status) {
        if (status != SGS_OK) throw std::runtime_error(std::string("sgs: ") + sgs_last_error());
    }

    // MapPoint only exposes k * mfMinDistance / k * mfMaxDistance (GetMin/MaxDistanceInvariance, k = 0.8f / 1.2f); the C ABI takes the raw distances and
    // applies k itself.  inv / k is not always a float whose product with k gives inv back, so the neighbours are tried: the value returned satisfies
    // k * x == inv exactly, i.e. the kernels' distance gates are bit-identical to the reference's.  (PredictScale divides the raw mfMaxDistance; where several
    // floats satisfy the equation the one chosen may be 1 ulp off the member -- add `float GetMaxDistance()` to MapPoint to remove even that.)
    static float raw_distance(float inv, float k) {
        float x = inv / k;
        if (k * x != inv) {
            const float up = std::nextafter(x, INFINITY), dn = std::nextafter(x, -INFINITY);
            if (k * up == inv) x = up; else if (k * dn == inv) x = dn;
        }
        return x;
    }

    float mfNNratio;
    bool mbCheckOrientation;
    int device_;
};

}  // namespace ORB_SLAM2

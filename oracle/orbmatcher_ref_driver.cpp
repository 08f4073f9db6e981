// C entry points around the reference's OWN ORBmatcher (src/ORBmatcher.cc compiled from the reference tree against the stand-ins of orbmatcher_shim/):
// flat arrays in, the reference's Frame / MapPoint objects built from them, the reference's method called, its side effects read back.  Used by
// tests/test_orbmatcher_ref.py to pin the oracle's restatement of the matchers' control flow.  TEST INFRASTRUCTURE (oracle/_ref/liborbmatcher_ref.so).
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM2;

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };      // sgs_keypoint / cv::KeyPoint layout of the oracle

void fill_frame(Frame& F, int n, const Kp* kps, const float* uright, const uint8_t* desc, const float* cam /* fx fy cx cy bf minx miny maxx maxy */, int nlevels,
                const float* scale_factors, const float* tcw16) {
    F.N = n;
    F.mvKeysUn.resize(n); F.mvKeys.resize(n); F.mvuRight.assign(uright, uright + n);
    F.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint k; k.pt.x = kps[i].x; k.pt.y = kps[i].y; k.size = kps[i].size; k.angle = kps[i].angle; k.response = kps[i].response; k.octave = kps[i].octave; k.class_id = kps[i].class_id;
        F.mvKeysUn[i] = k; F.mvKeys[i] = k;
        std::memcpy(F.mDescriptors.ptr<uint8_t>(i), desc + 32 * (size_t)i, 32);
    }
    F.fx = cam[0]; F.fy = cam[1]; F.cx = cam[2]; F.cy = cam[3]; F.mbf = cam[4]; F.mb = F.mbf / F.fx;                     // src/Frame.cc:196
    F.mnMinX = cam[5]; F.mnMinY = cam[6]; F.mnMaxX = cam[7]; F.mnMaxY = cam[8];
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.mnMaxX - F.mnMinX);            // src/Frame.cc:180-181
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.mnMaxY - F.mnMinY);
    F.mnScaleLevels = nlevels; F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    F.mfScaleFactor = nlevels > 1 ? scale_factors[1] : 1.2f; F.mfLogScaleFactor = log(F.mfScaleFactor);                   // src/Frame.cc:139
    F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL)); F.mvbOutlier.assign(n, false);
    F.mTcw.create(4, 4, CV_32F);
    if (tcw16) for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) F.mTcw.at<float>(r, c) = tcw16[4 * r + c];
    F.AssignFeaturesToGrid();
}

void set_vec3(cv::Mat& m, const float* p) { m.create(3, 1, CV_32F); for (int k = 0; k < 3; ++k) m.at<float>(k) = p[k]; }
void set_desc(cv::Mat& m, const uint8_t* d) { m.create(1, 32, CV_8U); std::memcpy(m.ptr<uint8_t>(0), d, 32); }
}  // namespace

// ORBmatcher(0.9, check_ori).SearchByProjection(CurrentFrame, LastFrame, th, mono)  (src/ORBmatcher.cc:1332-1472).  cur_mp_inout: index of the last-frame point
// already assigned to current keypoint j (or -1) on entry, the assignment after the call on return; cur_mp_obs_in: Observations() > 0 of that point.
REF_API int ref_search_by_projection_last(int ncur, const Kp* cur_kps, const float* cur_uright, const uint8_t* cur_desc, const float* cam, int nlevels,
                                          const float* scale_factors, const float* tcw_cur, const float* tcw_last, int nlast, const uint8_t* last_has_mp,
                                          const float* last_xyz, const uint8_t* last_desc, const uint8_t* last_obs, const int32_t* last_octave, const float* last_angle,
                                          float th, int mono, int check_ori, int32_t* cur_mp_inout, const uint8_t* cur_mp_obs_in) {
    Frame cur, last;
    fill_frame(cur, ncur, cur_kps, cur_uright, cur_desc, cam, nlevels, scale_factors, tcw_cur);
    std::vector<Kp> lk(nlast);
    for (int i = 0; i < nlast; ++i) { lk[i] = Kp{0, 0, 31.f, last_angle[i], 0, last_octave[i], -1}; }
    std::vector<float> lur(nlast, -1.f); std::vector<uint8_t> ld((size_t)(nlast > 0 ? nlast : 1) * 32, 0);
    fill_frame(last, nlast, lk.data(), lur.data(), ld.data(), cam, nlevels, scale_factors, tcw_last);
    std::vector<MapPoint> pts(nlast), pre(ncur);
    for (int i = 0; i < nlast; ++i) {
        if (!last_has_mp[i]) continue;
        set_vec3(pts[i].mWorldPos, last_xyz + 3 * i); set_desc(pts[i].mDescriptor, last_desc + 32 * (size_t)i); pts[i].nObs = last_obs[i] ? 1 : 0;
        last.mvpMapPoints[i] = &pts[i];
    }
    for (int j = 0; j < ncur; ++j)
        if (cur_mp_inout[j] >= 0) { pre[j].nObs = cur_mp_obs_in ? cur_mp_obs_in[j] : 1; cur.mvpMapPoints[j] = &pre[j]; }      // a map point assigned before the call
    ORBmatcher matcher(0.9, check_ori != 0);
    const int nm = matcher.SearchByProjection(cur, last, th, mono != 0);
    for (int j = 0; j < ncur; ++j) {
        MapPoint* p = cur.mvpMapPoints[j];
        if (!p) cur_mp_inout[j] = -1;
        else if (p >= pts.data() && p < pts.data() + nlast) cur_mp_inout[j] = (int32_t)(p - pts.data());
        // else: the pre-assigned stand-in stays as it was
    }
    return nm;
}

// ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th)  (src/ORBmatcher.cc:45-129).  The per-point inputs are what Frame::isInFrustum leaves in the MapPoint.
REF_API int ref_search_by_projection_local(int n, const Kp* kps, const float* uright, const uint8_t* desc, const float* cam, int nlevels, const float* scale_factors,
                                           int nmp, const uint8_t* inview, const float* projx, const float* projy, const float* projxr, const int32_t* level,
                                           const float* viewcos, const uint8_t* mp_desc, const uint8_t* mp_obs, float th, float nnratio, int id_base,
                                           int32_t* f_mp_inout, uint8_t* f_mp_obs_inout) {
    Frame F;
    fill_frame(F, n, kps, uright, desc, cam, nlevels, scale_factors, nullptr);
    std::vector<MapPoint> pts(nmp), pre(n);
    std::vector<MapPoint*> vp(nmp);
    for (int i = 0; i < nmp; ++i) {
        MapPoint& p = pts[i];
        p.mbTrackInView = inview[i] != 0; p.mTrackProjX = projx[i]; p.mTrackProjY = projy[i]; p.mTrackProjXR = projxr[i]; p.mnTrackScaleLevel = level[i];
        p.mTrackViewCos = viewcos[i]; set_desc(p.mDescriptor, mp_desc + 32 * (size_t)i); p.nObs = mp_obs[i] ? 1 : 0;
        vp[i] = &p;
    }
    for (int j = 0; j < n; ++j) if (f_mp_inout[j] >= 0) { pre[j].nObs = f_mp_obs_inout[j] ? 1 : 0; F.mvpMapPoints[j] = &pre[j]; }
    ORBmatcher matcher(nnratio);
    const int nm = matcher.SearchByProjection(F, vp, th);
    for (int j = 0; j < n; ++j) {
        MapPoint* p = F.mvpMapPoints[j];
        if (p && p >= pts.data() && p < pts.data() + nmp) { f_mp_inout[j] = id_base + (int32_t)(p - pts.data()); f_mp_obs_inout[j] = p->nObs > 0; }
    }
    return nm;
}

REF_API int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma, mb; set_desc(ma, a); set_desc(mb, b);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

// ORBmatcher(0.75, check_ori).SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)  (src/ORBmatcher.cc:1474-1601, relocalisation).
// kf_valid[i] = the key frame's map point i exists, is not bad and is not in sAlreadyFound; cur_mp_inout[j] >= 0 = keypoint j already holds a map point.
REF_API int ref_search_by_projection_kf(int ncur, const Kp* cur_kps, const float* cur_uright, const uint8_t* cur_desc, const float* cam, int nlevels,
                                        const float* scale_factors, const float* tcw_cur, int nkf, const uint8_t* kf_valid, const float* kf_xyz, const uint8_t* kf_desc,
                                        const float* kf_angle, const float* min_dist, const float* max_dist, float th, int orb_dist, int check_ori, int32_t* cur_mp_inout) {
    Frame cur;
    fill_frame(cur, ncur, cur_kps, cur_uright, cur_desc, cam, nlevels, scale_factors, tcw_cur);
    KeyFrame kf;
    kf.N = nkf; kf.mvKeysUn.resize(nkf); kf.mvpMapPoints.assign(nkf, static_cast<MapPoint*>(NULL));
    std::vector<MapPoint> pts(nkf), pre(ncur), found(nkf);
    std::set<MapPoint*> sAlreadyFound;
    for (int i = 0; i < nkf; ++i) {
        kf.mvKeysUn[i].angle = kf_angle[i];
        MapPoint& p = pts[i];
        set_vec3(p.mWorldPos, kf_xyz + 3 * i); set_desc(p.mDescriptor, kf_desc + 32 * (size_t)i); p.mfMinDistance = min_dist[i]; p.mfMaxDistance = max_dist[i]; p.nObs = 1;
        if (kf_valid[i]) kf.mvpMapPoints[i] = &p;
    }
    for (int j = 0; j < ncur; ++j) if (cur_mp_inout[j] >= 0) cur.mvpMapPoints[j] = &pre[j];
    ORBmatcher matcher(0.75, check_ori != 0);
    const int nm = matcher.SearchByProjection(cur, &kf, sAlreadyFound, th, orb_dist);
    for (int j = 0; j < ncur; ++j) {
        MapPoint* p = cur.mvpMapPoints[j];
        if (!p) cur_mp_inout[j] = -1;
        else if (p >= pts.data() && p < pts.data() + nkf) cur_mp_inout[j] = (int32_t)(p - pts.data());
    }
    return nm;
}

// ORBmatcher(nnratio, check_ori).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  (src/ORBmatcher.cc:407-522)
REF_API int ref_search_for_initialization(int n1, const Kp* k1, const uint8_t* d1, int n2, const Kp* k2, const uint8_t* d2, const float* cam, int nlevels,
                                          const float* scale_factors, float* prev_xy_inout, int window_size, float nnratio, int check_ori, int32_t* match12) {
    Frame F1, F2;
    std::vector<float> u1(n1, -1.f), u2(n2, -1.f);
    fill_frame(F1, n1, k1, u1.data(), d1, cam, nlevels, scale_factors, nullptr);
    fill_frame(F2, n2, k2, u2.data(), d2, cam, nlevels, scale_factors, nullptr);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_xy_inout[2 * i], prev_xy_inout[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, window_size);
    for (int i = 0; i < n1; ++i) { match12[i] = m12[i]; prev_xy_inout[2 * i] = prev[i].x; prev_xy_inout[2 * i + 1] = prev[i].y; }
    return nm;
}

namespace {
void fill_featvec(DBoW2::FeatureVector& fv, int n, const int32_t* node, const double* weight) {
    for (int i = 0; i < n; ++i) if (weight[i] > 0) fv.addFeature((DBoW2::NodeId)node[i], (unsigned)i);      // TemplatedVocabulary::transform adds a feature iff its word weight > 0
}
}  // namespace

// ORBmatcher(nnratio, check_ori).SearchByBoW(pKF, F, vpMapPointMatches)  (src/ORBmatcher.cc:159-290).  match_f[j] = key-frame feature whose map point went to feature j.
REF_API int ref_search_by_bow(int nkf, const int32_t* kf_node, const double* kf_weight, const uint8_t* kf_valid, const uint8_t* kf_desc, const float* kf_angle, int nf,
                              const int32_t* f_node, const double* f_weight, const uint8_t* f_desc, const float* f_angle, float nnratio, int check_ori, int32_t* match_f) {
    KeyFrame kf; Frame F;
    kf.N = nkf; kf.mvKeysUn.resize(nkf); kf.mvpMapPoints.assign(nkf, static_cast<MapPoint*>(NULL)); kf.mDescriptors.create(nkf > 0 ? nkf : 1, 32, CV_8U);
    std::vector<MapPoint> pts(nkf);
    for (int i = 0; i < nkf; ++i) { kf.mvKeysUn[i].angle = kf_angle[i]; std::memcpy(kf.mDescriptors.ptr<uint8_t>(i), kf_desc + 32 * (size_t)i, 32); if (kf_valid[i]) kf.mvpMapPoints[i] = &pts[i]; }
    fill_featvec(kf.mFeatVec, nkf, kf_node, kf_weight);
    F.N = nf; F.mvKeys.resize(nf); F.mvKeysUn.resize(nf); F.mDescriptors.create(nf > 0 ? nf : 1, 32, CV_8U);
    for (int j = 0; j < nf; ++j) { F.mvKeys[j].angle = f_angle[j]; F.mvKeysUn[j].angle = f_angle[j]; std::memcpy(F.mDescriptors.ptr<uint8_t>(j), f_desc + 32 * (size_t)j, 32); }
    fill_featvec(F.mFeatVec, nf, f_node, f_weight);
    std::vector<MapPoint*> vp;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchByBoW(&kf, F, vp);
    for (int j = 0; j < nf; ++j) match_f[j] = vp[j] ? (int32_t)(vp[j] - pts.data()) : -1;
    return nm;
}

// ORBmatcher(nnratio, check_ori).SearchByBoW(pKF1, pKF2, vpMatches12)  (src/ORBmatcher.cc:524-657).  match_1[i1] = feature of KF2 whose map point matched, or -1.
REF_API int ref_search_by_bow_kfkf(int n1, const int32_t* node1, const double* weight1, const uint8_t* valid1, const uint8_t* desc1, const float* angle1, int n2,
                                   const int32_t* node2, const double* weight2, const uint8_t* valid2, const uint8_t* desc2, const float* angle2, float nnratio,
                                   int check_ori, int32_t* match_1) {
    KeyFrame k1, k2;
    std::vector<MapPoint> p1(n1), p2(n2);
    auto fill = [](KeyFrame& k, std::vector<MapPoint>& pts, int n, const int32_t* node, const double* w, const uint8_t* valid, const uint8_t* desc, const float* ang) {
        k.N = n; k.mvKeysUn.resize(n); k.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL)); k.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
        for (int i = 0; i < n; ++i) { k.mvKeysUn[i].angle = ang[i]; std::memcpy(k.mDescriptors.ptr<uint8_t>(i), desc + 32 * (size_t)i, 32); if (valid[i]) k.mvpMapPoints[i] = &pts[i]; }
        fill_featvec(k.mFeatVec, n, node, w);
    };
    fill(k1, p1, n1, node1, weight1, valid1, desc1, angle1); fill(k2, p2, n2, node2, weight2, valid2, desc2, angle2);
    std::vector<MapPoint*> vp;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchByBoW(&k1, &k2, vp);
    for (int i = 0; i < n1; ++i) match_1[i] = vp[i] ? (int32_t)(vp[i] - p2.data()) : -1;
    return nm;
}

// ORBmatcher(0.6, check_ori).SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  (src/ORBmatcher.cc:659-827).  free = the feature has NO map
// point; stereo = mvuRight >= 0.  The epipole comes from the poses as in the reference: KF2 at the identity, KF1's camera centre at `cw1`.  match_1[i1] = i2 or -1.
REF_API int ref_search_for_triangulation(int n1, const int32_t* node1, const double* weight1, const uint8_t* free1, const uint8_t* stereo1, const uint8_t* desc1,
                                         const float* xy1, const float* angle1, int n2, const int32_t* node2, const double* weight2, const uint8_t* free2,
                                         const uint8_t* stereo2, const uint8_t* desc2, const float* xy2, const int32_t* octave2, const float* angle2, const float* F12,
                                         const float* cw1, const float* cam, int nlevels, const float* sigma2, const float* scale, int only_stereo, int check_ori,
                                         int32_t* match_1) {
    KeyFrame k1, k2;
    std::vector<MapPoint> p1(n1), p2(n2);
    auto fill = [&](KeyFrame& k, std::vector<MapPoint>& pts, int n, const int32_t* node, const double* w, const uint8_t* fr, const uint8_t* st, const uint8_t* desc,
                    const float* xy, const int32_t* oct, const float* ang) {
        k.N = n; k.mvKeysUn.resize(n); k.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL)); k.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U); k.mvuRight.resize(n);
        for (int i = 0; i < n; ++i) {
            k.mvKeysUn[i].pt.x = xy[2 * i]; k.mvKeysUn[i].pt.y = xy[2 * i + 1]; k.mvKeysUn[i].angle = ang[i]; k.mvKeysUn[i].octave = oct ? oct[i] : 0;
            std::memcpy(k.mDescriptors.ptr<uint8_t>(i), desc + 32 * (size_t)i, 32);
            if (!fr[i]) k.mvpMapPoints[i] = &pts[i];
            k.mvuRight[i] = st[i] ? 1.f : -1.f;
        }
        fill_featvec(k.mFeatVec, n, node, w);
        k.fx = cam[0]; k.fy = cam[1]; k.cx = cam[2]; k.cy = cam[3];
        k.mvScaleFactors.assign(scale, scale + nlevels); k.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
        k.Tcw = cv::Mat::eye(4, 4, CV_32F); k.Ow = cv::Mat::zeros(3, 1, CV_32F);
    };
    fill(k1, p1, n1, node1, weight1, free1, stereo1, desc1, xy1, nullptr, angle1);
    fill(k2, p2, n2, node2, weight2, free2, stereo2, desc2, xy2, octave2, angle2);
    for (int k = 0; k < 3; ++k) k1.Ow.at<float>(k) = cw1[k];
    cv::Mat F(3, 3, CV_32F);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F.at<float>(r, c) = F12[3 * r + c];
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6, check_ori != 0);
    const int nm = matcher.SearchForTriangulation(&k1, &k2, F, pairs, only_stereo != 0);
    for (int i = 0; i < n1; ++i) match_1[i] = -1;
    for (const auto& pr : pairs) match_1[pr.first] = (int32_t)pr.second;
    return nm;
}

// ORBmatcher(0.75, true).SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)  (src/ORBmatcher.cc:292-405, loop closing).  kf_matched_inout[idx] >= 0 = the key
// frame's feature idx is already matched on entry; features claimed by the call hold the claiming point's index afterwards.
REF_API int ref_search_by_projection_sim3(int n, const Kp* kps, const float* uright, const uint8_t* desc, const float* cam, int nlevels, const float* scale_factors,
                                          const float* scw16, int nmp, const uint8_t* mp_valid, const float* mp_xyz, const float* mp_normal, const float* min_dist,
                                          const float* max_dist, const uint8_t* mp_desc, int th, int32_t* kf_matched_inout) {
    KeyFrame kf;
    {   // the grid owner part of a key frame is filled like a frame's
        Frame tmp;
        fill_frame(tmp, n, kps, uright, desc, cam, nlevels, scale_factors, nullptr);
        static_cast<GridOwner&>(kf) = static_cast<GridOwner&>(tmp);
    }
    std::vector<MapPoint> pts(nmp), taken(n);
    std::vector<MapPoint*> vp(nmp), vm(n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < nmp; ++i) {
        MapPoint& p = pts[i];
        set_vec3(p.mWorldPos, mp_xyz + 3 * i); set_vec3(p.mNormalVector, mp_normal + 3 * i); set_desc(p.mDescriptor, mp_desc + 32 * (size_t)i);
        p.mfMinDistance = min_dist[i]; p.mfMaxDistance = max_dist[i]; p.mbBad = !mp_valid[i]; p.nObs = 1;
        vp[i] = &p;
    }
    for (int j = 0; j < n; ++j) if (kf_matched_inout[j] >= 0) vm[j] = &taken[j];
    cv::Mat Scw(4, 4, CV_32F);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Scw.at<float>(r, c) = scw16[4 * r + c];
    ORBmatcher matcher(0.75, true);
    const int nm = matcher.SearchByProjection(&kf, Scw, vp, vm, th);
    for (int j = 0; j < n; ++j) if (vm[j] && vm[j] >= pts.data() && vm[j] < pts.data() + nmp) kf_matched_inout[j] = (int32_t)(vm[j] - pts.data());
    return nm;
}

// ORBmatcher().Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:829-980) and Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:982-1104, scw16 != NULL).  The key
// frame starts without map points; kf_obs[j] = -1, or the Observations() of a map point the feature j already holds.  Outputs per input point: best_idx =
// the feature it was fused onto (read back from the side effects: AddObservation's index, the resident of a Replace, vpReplacePoint) or -1 when the reference
// did not fuse it; returns nFused.
REF_API int ref_fuse(int n, const Kp* kps, const float* uright, const uint8_t* desc, const float* cam, int nlevels, const float* scale_factors, const float* tcw16,
                     const float* ow3, const float* scw16, int nmp, const uint8_t* mp_valid, const float* mp_xyz, const float* mp_normal, const float* min_dist,
                     const float* max_dist, const uint8_t* mp_desc, const int32_t* mp_nobs, const int32_t* kf_obs, float th, const float* inv_level_sigma2,
                     int32_t* best_idx) {
    KeyFrame kf;
    {
        Frame tmp;
        fill_frame(tmp, n, kps, uright, desc, cam, nlevels, scale_factors, nullptr);
        static_cast<GridOwner&>(kf) = static_cast<GridOwner&>(tmp);
    }
    kf.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
    kf.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL));
    kf.Tcw.create(4, 4, CV_32F);
    if (tcw16) for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) kf.Tcw.at<float>(r, c) = tcw16[4 * r + c];
    kf.Ow.create(3, 1, CV_32F);
    if (ow3) for (int k = 0; k < 3; ++k) kf.Ow.at<float>(k) = ow3[k];
    std::vector<MapPoint> pts(nmp), res(n);
    std::vector<MapPoint*> vp(nmp);
    for (int j = 0; j < n; ++j) if (kf_obs[j] >= 0) { res[j].nObs = kf_obs[j]; res[j].mObservations[&kf] = j; kf.mvpMapPoints[j] = &res[j]; }
    for (int i = 0; i < nmp; ++i) {
        MapPoint& p = pts[i];
        set_vec3(p.mWorldPos, mp_xyz + 3 * i); set_vec3(p.mNormalVector, mp_normal + 3 * i); set_desc(p.mDescriptor, mp_desc + 32 * (size_t)i);
        p.mfMinDistance = min_dist[i]; p.mfMaxDistance = max_dist[i]; p.mbBad = !mp_valid[i]; p.nObs = mp_nobs[i];
        vp[i] = &p; best_idx[i] = -1;
    }
    std::vector<RefEvent> log;
    ref_event_log() = &log;
    int nf;
    ORBmatcher matcher;
    if (scw16) {
        cv::Mat Scw(4, 4, CV_32F);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Scw.at<float>(r, c) = scw16[4 * r + c];
        std::vector<MapPoint*> rep(nmp, static_cast<MapPoint*>(NULL));
        nf = matcher.Fuse(&kf, Scw, vp, th, rep);
        for (int i = 0; i < nmp; ++i) if (rep[i]) best_idx[i] = (int32_t)rep[i]->GetIndexInKeyFrame(&kf);
    } else nf = matcher.Fuse(&kf, vp, th);
    ref_event_log() = nullptr;
    auto is_input = [&](MapPoint* p) { return p >= pts.data() && p < pts.data() + nmp; };
    for (const RefEvent& e : log) {
        if (e.kind == 0) { if (is_input(e.a)) best_idx[e.a - pts.data()] = (int32_t)e.idx; }
        else {                                                         // a->Replace(b): one of the two is the point being fused, the other sits in the key frame
            MapPoint* in = nullptr; MapPoint* resident = nullptr;
            if (e.a->IsInKeyFrame(&kf) && !e.b->IsInKeyFrame(&kf)) { resident = e.a; in = e.b; }
            else if (e.b->IsInKeyFrame(&kf) && !e.a->IsInKeyFrame(&kf)) { resident = e.b; in = e.a; }
            if (in && is_input(in)) best_idx[in - pts.data()] = (int32_t)resident->GetIndexInKeyFrame(&kf);
        }
    }
    return nf;
}

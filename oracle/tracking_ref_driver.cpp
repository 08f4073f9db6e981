// C entry point around the reference's OWN tracking front end: Tracking::TrackWithMotionModel (src/Tracking.cc:906-967) and Tracking::TrackLocalMap (:969-1014, with
// UpdateLocalMap / UpdateLocalPoints :1314-1348 and SearchLocalPoints :1262-1312) running on the reference's own ORBmatcher::SearchByProjection (both forms),
// Optimizer::PoseOptimization (+ vendored g2o), Frame::isInFrustum / GetFeaturesInArea and MapPoint::PredictScale -- every source compiled unmodified from where it
// lies (oracle/Makefile, target ref_tracking; stand-ins in tracking_shim/tracking_standins.h).  The driver only builds the object graph the flat arrays describe
// (frames, map points, one key frame that owns the local map) and reads the results back.  TEST INFRASTRUCTURE (oracle/_ref/libtracking_ref.so).
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "Tracking.h"
#include "Optimizer.h"

using namespace ORB_SLAM2;
#define REF_API extern "C" __attribute__((visibility("default")))

long unsigned int KeyFrame::nNextId = 0;
KeyFrame::KeyFrame(Frame& F, Map*, KeyFrameDatabase*) : mnFrameId(F.mnId), N(F.N), mvpMapPoints(F.mvpMapPoints) { mnId = nNextId++; }     // CreateNewKeyFrame: compiled, never run
cv::Mat KeyFrame::GetPoseInverse() { return Tcw.clone(); }
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float&, const float&, const float&) const { return std::vector<size_t>(); }       // the key-frame matchers: compiled, never run

namespace {

struct OpenPoint : MapPoint {      // the protected state a map point normally acquires through the mapping thread
    OpenPoint(const cv::Mat& p, KeyFrame* kf, Map* m) : MapPoint(p, kf, m) {}
    void set(const float* normal, float mn, float mx, const uint8_t* desc, int observations, bool bad) {
        mNormalVector = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) mNormalVector.at<float>(k) = normal ? normal[k] : 0.f;
        mfMinDistance = mn; mfMaxDistance = mx;
        mDescriptor = cv::Mat(1, 32, CV_8U); std::memcpy(mDescriptor.ptr(0), desc, 32);
        nObs = observations; mbBad = bad;
    }
};

struct OpenTracking : Tracking {
    using Tracking::Tracking;
    bool motion_model() { return TrackWithMotionModel(); }
    bool local_map() { return TrackLocalMap(); }
    bool reference_keyframe(KeyFrame* kf) { mpReferenceKF = kf; return TrackReferenceKeyFrame(); }
    void state(const Frame& cur, const Frame& last, const cv::Mat& velocity, const cv::Mat& Tlr, KeyFrame* local) {
        mCurrentFrame = Frame(cur); mLastFrame = Frame(last);
        mVelocity = velocity.clone();
        mlRelativeFramePoses.clear(); mlRelativeFramePoses.push_back(Tlr.clone());
        mvpLocalKeyFrames.assign(1, local); mvpLocalMapPoints.clear();
        mnLastKeyFrameId = 0; mnLastRelocFrameId = 0; mbOnlyTracking = false; mState = OK;
    }
    int inliers() const { return mnMatchesInliers; }
};

cv::Mat mat44(const float* T) { cv::Mat m(4, 4, CV_32F); for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m.at<float>(i, j) = T[4 * i + j]; return m; }
cv::Mat vec3(const float* p) { cv::Mat m(3, 1, CV_32F); for (int k = 0; k < 3; ++k) m.at<float>(k) = p[k]; return m; }

void fill_frame(Frame& F, int n, const float* xy, const int32_t* octave, const float* angle, const float* uright, const uint8_t* desc, const float* scale_factors,
                const float* inv_level_sigma2, int nlevels, float bf, long unsigned int id) {
    F.N = n; F.mnId = id;
    F.mvKeys.resize(n); F.mvKeysUn.resize(n); F.mvuRight.assign(n, -1.f); F.mvDepth.assign(n, -1.f);
    F.mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8U);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint k; k.pt.x = xy[2 * i]; k.pt.y = xy[2 * i + 1]; k.octave = octave[i]; k.angle = angle ? angle[i] : 0.f;
        F.mvKeys[i] = k; F.mvKeysUn[i] = k;                                   // no distortion: the undistorted keypoints are the keypoints (src/Frame.cc:657-660)
        if (uright) F.mvuRight[i] = uright[i];
        if (desc) std::memcpy(F.mDescriptors.ptr(i), desc + 32 * (size_t)i, 32);
    }
    F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL)); F.mvbOutlier.assign(n, false);
    F.mnScaleLevels = nlevels; F.mfScaleFactor = scale_factors[1]; F.mfLogScaleFactor = logf(F.mfScaleFactor);
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels); F.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
    F.mvInvScaleFactors.resize(nlevels); F.mvLevelSigma2.resize(nlevels);
    for (int l = 0; l < nlevels; ++l) { F.mvInvScaleFactors[l] = 1.0f / F.mvScaleFactors[l]; F.mvLevelSigma2[l] = F.mvScaleFactors[l] * F.mvScaleFactors[l]; }
    F.mbf = bf; F.mb = bf / Frame::fx; F.mThDepth = 40.f * F.mb;
    for (int i = 0; i < n; ++i) { int gx, gy; if (F.PosInGrid(F.mvKeysUn[i], gx, gy)) F.mGrid[gx][gy].push_back(i); }      // Frame::AssignFeaturesToGrid (private; :257-272), pinned by libframe_ref.so
}

}  // namespace

// cam9 = fx, fy, cx, cy, bf, min_x, min_y, max_x, max_y.  Current frame: n keypoints (undistorted).  Last frame: m keypoints with flags bit 0 = holds a map point,
// bit 1 = that point has observations, bit 2 = it is bad; last_local_id[j] >= 0 says that the point IS local-map point number last_local_id[j] (one object).
// Local map: L points (valid = not bad).  Outputs: after TrackWithMotionModel -- its return value, the pose, per keypoint the last-frame index it holds (-1 none);
// after TrackLocalMap -- its return value, the pose, per keypoint the point id (last-frame index, or point_cap + local index for points added by the local search),
// the outlier flags, mnMatchesInliers.
REF_API void ref_track_motion_and_local_map(const float* cam9, const float* scale_factors, const float* inv_level_sigma2, int nlevels,
                                            int n, const float* cur_xy, const int32_t* cur_octave, const float* cur_angle, const float* cur_uright, const uint8_t* cur_desc, const float* tcw_cur,
                                            int m, const float* last_xyz, const uint8_t* last_desc, const uint8_t* last_flags, const int32_t* last_octave, const float* last_angle,
                                            const float* tcw_last, const int32_t* last_local_id,
                                            int L, const float* mp_xyz, const float* mp_normal, const float* mp_min, const float* mp_max, const uint8_t* mp_desc, const uint8_t* mp_valid,
                                            const uint8_t* mp_obs, int point_cap,
                                            int32_t* ok_motion, float* tcw_motion, int32_t* mp_motion, int32_t* ok_local, float* tcw_final, int32_t* mp_final, uint8_t* outlier, int32_t* inliers) {
    Frame::fx = cam9[0]; Frame::fy = cam9[1]; Frame::cx = cam9[2]; Frame::cy = cam9[3]; Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
    Frame::mnMinX = cam9[5]; Frame::mnMinY = cam9[6]; Frame::mnMaxX = cam9[7]; Frame::mnMaxY = cam9[8];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);         // src/Frame.cc:175-176
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    Frame::mbInitialComputations = false;
    std::map<std::string, double>& S = cv::FileStorage::values();
    S.clear();
    S["Camera.fx"] = cam9[0]; S["Camera.fy"] = cam9[1]; S["Camera.cx"] = cam9[2]; S["Camera.cy"] = cam9[3]; S["Camera.bf"] = cam9[4]; S["Camera.fps"] = 30; S["Camera.RGB"] = 1;
    S["ORBextractor.nFeatures"] = 1000; S["ORBextractor.scaleFactor"] = scale_factors[1]; S["ORBextractor.nLevels"] = nlevels; S["ORBextractor.iniThFAST"] = 20; S["ORBextractor.minThFAST"] = 7;
    S["ThDepth"] = 40; S["DepthMapFactor"] = 1;
    Map map; KeyFrameDatabase db; ORBVocabulary voc; FrameDrawer fd; MapDrawer md; System sys;
    OpenTracking trk(&sys, &voc, &fd, &md, &map, &db, std::string("planted"), (int)System::RGBD, boost::shared_ptr<PointCloudMapping>());

    KeyFrame ref_kf; ref_kf.Tcw = cv::Mat::eye(4, 4, CV_32F);                  // reference key frame of the last frame: identity, so that UpdateLastFrame leaves the pose as given
    KeyFrame local_kf;                                                         // the one local key frame: it owns the local map points (UpdateLocalPoints collects them)
    std::vector<std::unique_ptr<OpenPoint> > pool;
    std::vector<MapPoint*> local(L, static_cast<MapPoint*>(NULL));
    for (int l = 0; l < L; ++l) {
        pool.emplace_back(new OpenPoint(vec3(mp_xyz + 3 * l), &ref_kf, &map));
        pool.back()->set(mp_normal + 3 * l, mp_min[l], mp_max[l], mp_desc + 32 * (size_t)l, mp_obs[l] ? 1 : 0, !mp_valid[l]);
        local[l] = pool.back().get();
    }
    local_kf.mvpMapPoints = local;

    Frame cur, last;
    fill_frame(cur, n, cur_xy, cur_octave, cur_angle, cur_uright, cur_desc, scale_factors, inv_level_sigma2, nlevels, cam9[4], 100);
    std::vector<float> zero2(2 * (size_t)(m > 0 ? m : 1), 0.f);
    fill_frame(last, m, zero2.data(), last_octave, last_angle, NULL, NULL, scale_factors, inv_level_sigma2, nlevels, cam9[4], 99);
    last.mTcw = mat44(tcw_last); last.UpdatePoseMatrices(); last.mpReferenceKF = &ref_kf;
    std::vector<MapPoint*> last_pts(m, static_cast<MapPoint*>(NULL));
    for (int j = 0; j < m; ++j) {
        if (!(last_flags[j] & 1)) continue;
        if (last_local_id[j] >= 0) last_pts[j] = local[last_local_id[j]];
        else {
            pool.emplace_back(new OpenPoint(vec3(last_xyz + 3 * j), &ref_kf, &map));
            pool.back()->set(NULL, 0.f, 0.f, last_desc + 32 * (size_t)j, (last_flags[j] & 2) ? 1 : 0, (last_flags[j] & 4) != 0);
            last_pts[j] = pool.back().get();
        }
        last.mvpMapPoints[j] = last_pts[j];
    }
    // mVelocity * mLastFrame.mTcw must be the given current pose: the velocity is that pose and the test's last pose is the identity (checked by the caller)
    trk.state(cur, last, mat44(tcw_cur), mat44(tcw_last), &local_kf);

    *ok_motion = trk.motion_model() ? 1 : 0;
    Frame& F = trk.mCurrentFrame;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) tcw_motion[4 * i + j] = F.mTcw.at<float>(i, j);
    std::vector<MapPoint*> held(F.mvpMapPoints);
    for (int i = 0; i < n; ++i) {
        mp_motion[i] = -1;
        if (held[i]) for (int j = 0; j < m; ++j) if (last_pts[j] == held[i]) { mp_motion[i] = j; break; }
    }
    *ok_local = trk.local_map() ? 1 : 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) tcw_final[4 * i + j] = F.mTcw.at<float>(i, j);
    for (int i = 0; i < n; ++i) {
        MapPoint* p = F.mvpMapPoints[i];
        mp_final[i] = -1; outlier[i] = (p && F.mvbOutlier[i]) ? 1 : 0;
        if (!p) continue;
        if (p == held[i]) { mp_final[i] = mp_motion[i]; continue; }            // kept from the motion-model search
        for (int l = 0; l < L; ++l) if (local[l] == p) { mp_final[i] = point_cap + l; break; }
    }
    *inliers = trk.inliers();
}

// Tracking::TrackReferenceKeyFrame (src/Tracking.cc:796-838): SearchByBoW(reference key frame, current frame) with nnratio 0.7, the pose of the last frame as the start,
// PoseOptimization, outlier discard.  *_node: the vocabulary node of every feature (the FeatureVector both sides hold; -1 = not listed); kf_flags bit 0 = the feature
// holds a map point, bit 1 = it has observations, bit 2 = it is bad.  Outputs: the return value, the pose, per keypoint the key-frame feature whose point it holds, the
// number of keypoints holding a point before the optimisation.
REF_API void ref_track_reference_keyframe(const float* cam9, const float* scale_factors, const float* inv_level_sigma2, int nlevels,
                                          int n, const float* cur_xy, const int32_t* cur_octave, const float* cur_angle, const float* cur_uright, const uint8_t* cur_desc, const int32_t* cur_node,
                                          int m, const float* kf_xyz, const uint8_t* kf_desc, const uint8_t* kf_flags, const float* kf_angle, const int32_t* kf_node, const float* tcw_last,
                                          int32_t* ok, float* tcw_out, int32_t* mp_out, int32_t* nheld) {
    Frame::fx = cam9[0]; Frame::fy = cam9[1]; Frame::cx = cam9[2]; Frame::cy = cam9[3]; Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
    Frame::mnMinX = cam9[5]; Frame::mnMinY = cam9[6]; Frame::mnMaxX = cam9[7]; Frame::mnMaxY = cam9[8];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    Frame::mbInitialComputations = false;
    std::map<std::string, double>& S = cv::FileStorage::values();
    S.clear();
    S["Camera.fx"] = cam9[0]; S["Camera.fy"] = cam9[1]; S["Camera.cx"] = cam9[2]; S["Camera.cy"] = cam9[3]; S["Camera.bf"] = cam9[4]; S["Camera.fps"] = 30; S["Camera.RGB"] = 1;
    S["ORBextractor.nFeatures"] = 1000; S["ORBextractor.scaleFactor"] = scale_factors[1]; S["ORBextractor.nLevels"] = nlevels; S["ORBextractor.iniThFAST"] = 20; S["ORBextractor.minThFAST"] = 7;
    S["ThDepth"] = 40; S["DepthMapFactor"] = 1;
    Map map; KeyFrameDatabase db; ORBVocabulary voc; FrameDrawer fd; MapDrawer md; System sys;
    OpenTracking trk(&sys, &voc, &fd, &md, &map, &db, std::string("planted"), (int)System::RGBD, boost::shared_ptr<PointCloudMapping>());

    KeyFrame kf; kf.Tcw = cv::Mat::eye(4, 4, CV_32F);
    kf.N = m; kf.mvKeysUn.resize(m); kf.mvKeys.resize(m); kf.mDescriptors = cv::Mat(m > 0 ? m : 1, 32, CV_8U); kf.mvpMapPoints.assign(m, static_cast<MapPoint*>(NULL));
    std::vector<std::unique_ptr<OpenPoint> > pool;
    for (int j = 0; j < m; ++j) {
        kf.mvKeysUn[j].angle = kf_angle[j]; kf.mvKeys[j].angle = kf_angle[j];
        std::memcpy(kf.mDescriptors.ptr(j), kf_desc + 32 * (size_t)j, 32);
        if (kf_node[j] >= 0) kf.mFeatVec[kf_node[j]].push_back(j);
        if (kf_flags[j] & 1) {
            pool.emplace_back(new OpenPoint(vec3(kf_xyz + 3 * j), &kf, &map));
            pool.back()->set(NULL, 0.f, 0.f, kf_desc + 32 * (size_t)j, (kf_flags[j] & 2) ? 1 : 0, (kf_flags[j] & 4) != 0);
            kf.mvpMapPoints[j] = pool.back().get();
        }
    }
    Frame cur, last;
    fill_frame(cur, n, cur_xy, cur_octave, cur_angle, cur_uright, cur_desc, scale_factors, inv_level_sigma2, nlevels, cam9[4], 100);
    for (int i = 0; i < n; ++i) if (cur_node[i] >= 0) cur.mFeatVec[cur_node[i]].push_back(i);
    cur.mBowVec[0] = 1.0;                                                     // not empty: Frame::ComputeBoW (src/Frame.cc:422-429) keeps the planted vectors
    cur.mpORBvocabulary = &voc;
    std::vector<float> zero2(2, 0.f); std::vector<int32_t> zo(1, 0);
    fill_frame(last, 0, zero2.data(), zo.data(), NULL, NULL, NULL, scale_factors, inv_level_sigma2, nlevels, cam9[4], 99);
    last.mTcw = mat44(tcw_last); last.UpdatePoseMatrices(); last.mpReferenceKF = &kf;
    trk.state(cur, last, cv::Mat::eye(4, 4, CV_32F), mat44(tcw_last), &kf);
    *ok = trk.reference_keyframe(&kf) ? 1 : 0;
    Frame& F = trk.mCurrentFrame;
    if (F.mTcw.empty()) F.mTcw = cv::Mat::eye(4, 4, CV_32F);                  // bailed out before SetPose
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) tcw_out[4 * i + j] = F.mTcw.at<float>(i, j);
    *nheld = 0;
    for (int i = 0; i < n; ++i) {
        mp_out[i] = -1;
        MapPoint* p = F.mvpMapPoints[i];
        if (!p) continue;
        ++*nheld;
        for (int j = 0; j < m; ++j) if (kf.mvpMapPoints[j] == p) { mp_out[i] = j; break; }
    }
}

// Stand-ins for the reference's Frame / KeyFrame / MapPoint (include/Frame.h, KeyFrame.h, MapPoint.h) carrying exactly the members src/ORBmatcher.cc touches,
// so that the reference's OWN ORBmatcher.cc compiles unmodified from where it lies (oracle/Makefile force-includes this header; it defines the include
// guards FRAME_H / KEYFRAME_H / MAPPOINT_H, which turns the real headers -- unusable here: they pull in DBoW2-on-OpenCV, g2o, ncnn, PCL -- into no-ops).
// What is restated here from the reference and therefore NOT pinned by this library: Frame::GetFeaturesInArea / PosInGrid / AssignFeaturesToGrid
// (src/Frame.cc:257-272, :354-419), KeyFrame::GetFeaturesInArea / IsInImage (src/KeyFrame.cc), MapPoint::PredictScale and the 0.8 / 1.2 invariance
// getters (src/MapPoint.cc:373-418).  Pinned: every line of src/ORBmatcher.cc.  TEST INFRASTRUCTURE.
#pragma once
#define FRAME_H
#define KEYFRAME_H
#define MAPPOINT_H
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;      // the reference's headers do this at namespace scope and ORBmatcher.h relies on it (`pair`, `vector`)

namespace ORB_SLAM2 {

class KeyFrame;
class Frame;
class MapPoint;

// side effects of Fuse, in call order, for the driver to read back (kind 0: a->AddObservation(kf, idx); kind 1: a->Replace(b))
struct RefEvent { int kind; MapPoint* a; MapPoint* b; long idx; };
inline std::vector<RefEvent>*& ref_event_log() { static std::vector<RefEvent>* log = nullptr; return log; }

class MapPoint {
public:
    cv::Mat mWorldPos, mNormalVector, mDescriptor;        // 3x1 CV_32F, 3x1 CV_32F, 1x32 CV_8U
    float mfMinDistance = 0, mfMaxDistance = 0;
    int nObs = 0;
    bool mbBad = false;
    // tracking scratch (written by Frame::isInFrustum in the reference, by the driver here)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    std::map<KeyFrame*, size_t> mObservations;
    MapPoint* mpReplaced = nullptr;
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    inline int PredictScale(const float& currentDist, KeyFrame* pKF);
    inline int PredictScale(const float& currentDist, Frame* pF);
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx) {
        if (ref_event_log()) ref_event_log()->push_back(RefEvent{0, this, nullptr, (long)idx});
        if (!mObservations.count(pKF)) { mObservations[pKF] = idx; ++nObs; }
    }
    // recorded only: the point stays usable, so that every later fusion onto the same feature leaves a trace too (the real Replace rewires the map)
    void Replace(MapPoint* pMP) { if (ref_event_log()) ref_event_log()->push_back(RefEvent{1, this, pMP, -1}); mpReplaced = pMP; }
};

struct GridOwner {       // the members Frame and KeyFrame share for the feature grid
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnScaleLevels = 0; float mfScaleFactor = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    DBoW2::BowVector mBowVec; DBoW2::FeatureVector mFeatVec;
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY) {          // src/Frame.cc:409-419
        posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
        posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
        return !(posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS);
    }
    void AssignFeaturesToGrid() {                                            // src/Frame.cc:257-272
        for (int i = 0; i < N; i++) { int gx, gy; if (PosInGrid(mvKeysUn[i], gx, gy)) mGrid[gx][gy].push_back(i); }
    }
    std::vector<size_t> FeaturesInArea(const float& x, const float& y, const float& r, const int minLevel, const int maxLevel) const {   // src/Frame.cc:354-407
        std::vector<size_t> vIndices;
        const int nMinCellX = max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
        const int nMaxCellX = min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
        const int nMaxCellY = min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (size_t j : mGrid[ix][iy]) {
                    const cv::KeyPoint& kpUn = mvKeysUn[j];
                    if (bCheckLevels) { if (kpUn.octave < minLevel) continue; if (maxLevel >= 0 && kpUn.octave > maxLevel) continue; }
                    if (fabs(kpUn.pt.x - x) < r && fabs(kpUn.pt.y - y) < r) vIndices.push_back(j);
                }
        return vIndices;
    }
};

class Frame : public GridOwner {
public:
    cv::Mat mTcw;                                          // 4x4 CV_32F
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        return FeaturesInArea(x, y, r, minLevel, maxLevel);
    }
};

class KeyFrame : public GridOwner {
public:
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Tcw, Ow;                                       // 4x4, 3x1
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return FeaturesInArea(x, y, r, -1, -1); }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }      // src/KeyFrame.cc
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
};

inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {      // src/MapPoint.cc:385-400
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0; else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}
inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) {          // src/MapPoint.cc:402-418
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0; else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

}  // namespace ORB_SLAM2

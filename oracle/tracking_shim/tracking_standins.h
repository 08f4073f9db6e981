// Force-included when compiling the reference's OWN tracking front end in one library (oracle/_ref/libtracking_ref.so, recipe in oracle/Makefile):
// src/Tracking.cc, src/Frame.cc, src/MapPoint.cc, src/ORBmatcher.cc, src/Optimizer.cc, src/Converter.cc, src/ORBextractor.cc and the vendored g2o, all unmodified,
// against the REAL include/Tracking.h, Frame.h, MapPoint.h, ORBmatcher.h, Optimizer.h, Converter.h, ORBextractor.h.  Stand-ins (this header defines the real
// headers' include guards) for what lies behind the tracking thread -- KeyFrame, Map, KeyFrameDatabase, ORBVocabulary, LocalMapping, LoopClosing, Viewer, FrameDrawer,
// MapDrawer, System, Initializer, PnPsolver, Detector2D, PointCloudMapping -- carrying exactly the members those sources touch; g2o's sparse-Cholesky wrapper is
// replaced as in g2o_shim/optimizer_standins.h.  Pinned through this library: Tracking::TrackWithMotionModel (src/Tracking.cc:906-967), SearchLocalPoints
// (:1262-1312) and TrackLocalMap (:969-1014) running on the reference's own SearchByProjection x2, PoseOptimization, isInFrustum, PredictScale.  TEST INFRASTRUCTURE.
#pragma once
#define KEYFRAME_H
#define MAP_H
#define KEYFRAMEDATABASE_H
#define ORBVOCABULARY_H
#ifndef SGS_REAL_LOCALMAPPING
#define LOCALMAPPING_H
#endif
#ifndef SGS_REAL_LOOPCLOSING
#define LOOPCLOSING_H
#endif
#define VIEWER_H
#define FRAMEDRAWER_H
#define MAPDRAWER_H
#define SYSTEM_H
#define INITIALIZER_H
#define PNPSOLVER_H
#define DETECTOR2D_H
#define G2O_LINEAR_SOLVER_EIGEN_H
#include <climits>
#include <cmath>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>
#include <opencv2/opencv.hpp>
#include "cv_tracking_extras.h"
#include <Eigen/StdVector>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "Thirdparty/g2o/g2o/core/linear_solver.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"

using namespace std;      // the reference's headers do this at namespace scope and rely on it

namespace g2o {
template <typename MatrixType> class LinearSolverEigen : public LinearSolver<MatrixType> {
public:
    virtual bool init() { return true; }
    virtual bool solve(const SparseBlockMatrix<MatrixType>&, double*, double*) { return false; }
    void setBlockOrdering(bool) {}
};
}

typedef struct Object2D {          // include/Detector2D.h:30-36
    cv::Rect_<float> rect;
    float prob;
    std::string name;
    int id;
} Object2D;

class PointCloudMapping {
public:
    template <class... A> void insertKeyFrame(A...) {}
};

namespace ORB_SLAM2 {

class Frame;
class MapPoint;
class Map;
class KeyFrame;
class KeyFrameDatabase;
class Tracking;

class ORBVocabulary {
public:
    void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) {}      // pinned separately (libdbow2_ref.so)
    double score(const DBoW2::BowVector&, const DBoW2::BowVector&) const { return 0.0; }
};

class Detector2D {                 // include/Detector2D.h:42-66
public:
    std::vector<Object2D> mvObjects2D;
    bool mbHaveDynamicObjectForMapping = false;
    bool mbHaveDynamicObjectForRmDynamicFeature = false;
    std::vector<cv::Rect_<float> > mvPotentialDynamicBorderForMapping;
    std::vector<cv::Rect_<float> > mvPotentialDynamicBorderForRmDynamicFeature;
    cv::Mat mImageToDetect;
    std::mutex mMutexGetNewImage, mMutexImageDetectFinished;
    bool mbNewImageFlag = false;
};

// the members ORBmatcher.cc / Optimizer.cc / MapPoint.cc / Tracking.cc read of a key frame; filled by the driver
class KeyFrame {
public:
    KeyFrame() {}
    KeyFrame(Frame& F, Map* pMap, KeyFrameDatabase* pKFDB);            // defined in the driver (needs the complete Frame)
    static long unsigned int nNextId;
    long unsigned int mnId = 0, mnFrameId = 0, mnTrackReferenceForFrame = 0, mnFuseTargetForKF = 0, mnBALocalForKF = 0, mnBAFixedForKF = 0, mnBAGlobalForKF = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mK, mTcwGBA, mTcwBefGBA, Tcw, Ow;
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnScaleLevels = 0; float mfScaleFactor = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    DBoW2::BowVector mBowVec; DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    bool mbBad = false;
    KeyFrame* mpParent = nullptr;
    bool isBad() { return mbBad; }
    void ComputeBoW() {}
    void SetPose(const cv::Mat& T) { T.copyTo(Tcw); }
    cv::Mat GetPose() { return Tcw.clone(); }
    cv::Mat GetPoseInverse();                                          // driver
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = nullptr; }
    void EraseMapPointMatch(MapPoint*) {}
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p) s.insert(p); return s; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    int TrackedMapPoints(const int&) { return 0; }
    void UpdateConnections() {}
    float ComputeSceneMedianDepth(const int) { return 1.f; }
    std::vector<KeyFrame*> GetBestCovisibilityKeyFrames(const int&) { return std::vector<KeyFrame*>(); }
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return std::vector<KeyFrame*>(); }
    std::vector<KeyFrame*> GetCovisiblesByWeight(const int&) { return std::vector<KeyFrame*>(); }
    std::set<KeyFrame*> GetChilds() { return std::set<KeyFrame*>(); }
    std::set<KeyFrame*> GetLoopEdges() { return std::set<KeyFrame*>(); }
    KeyFrame* GetParent() { return mpParent; }
    bool hasChild(KeyFrame*) { return false; }
    int GetWeight(KeyFrame*) { return 0; }
    void SetNotErase() {}
    void SetErase() {}
    void SetBadFlag() { mbBad = true; }
    std::set<KeyFrame*> GetConnectedKeyFrames() { return std::set<KeyFrame*>(); }
    void AddLoopEdge(KeyFrame*) {}
    cv::Mat UnprojectStereo(int) { return cv::Mat(); }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }      // src/KeyFrame.cc:611-614
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;                                     // driver (src/KeyFrame.cc:570-609)
};

class Map {
public:
    std::mutex mMutexMapUpdate, mMutexPointCreation;
    std::vector<KeyFrame*> mvpKeyFrameOrigins;
    std::vector<MapPoint*> mvpReferenceMapPoints;
    void AddKeyFrame(KeyFrame*) {}
    void AddMapPoint(MapPoint*) {}
    void EraseMapPoint(MapPoint*) {}
    void SetReferenceMapPoints(const std::vector<MapPoint*>& v) { mvpReferenceMapPoints = v; }
    std::vector<KeyFrame*> GetAllKeyFrames() { return std::vector<KeyFrame*>(); }
    std::vector<MapPoint*> GetAllMapPoints() { return std::vector<MapPoint*>(); }
    long unsigned int MapPointsInMap() { return 0; }
    long unsigned int KeyFramesInMap() { return 0; }
    long unsigned int GetMaxKFid() { return 0; }
    void InformNewBigChange() {}
    void clear() {}
};

class KeyFrameDatabase {
public:
    std::vector<KeyFrame*> DetectRelocalizationCandidates(Frame*) { return std::vector<KeyFrame*>(); }
    std::vector<KeyFrame*> DetectLoopCandidates(KeyFrame*, float) { return std::vector<KeyFrame*>(); }
    void add(KeyFrame*) {}
    void clear() {}
};

#ifndef SGS_REAL_LOCALMAPPING
class LocalMapping {
public:
    bool AcceptKeyFrames() { return true; }
    void InsertKeyFrame(KeyFrame*) {}
    void InterruptBA() {}
    int KeyframesInQueue() { return 0; }
    void RequestReset() {}
    bool SetNotStop(bool) { return true; }
    bool isStopped() { return false; }
    bool stopRequested() { return false; }
    void RequestStop() {}
    void Release() {}
    bool isFinished() { return true; }
};
#endif

#ifndef SGS_REAL_LOOPCLOSING
class LoopClosing {
public:
    typedef map<KeyFrame*, g2o::Sim3, std::less<KeyFrame*>, Eigen::aligned_allocator<std::pair<KeyFrame* const, g2o::Sim3> > > KeyFrameAndPose;    // include/LoopClosing.h:50-51
    void RequestReset() {}
    void InsertKeyFrame(KeyFrame*) {}
};
#endif

class Viewer { public: void Release() {} void RequestStop() {} bool isStopped() { return true; } };
class FrameDrawer { public: void Update(Tracking*) {} };
class MapDrawer { public: void SetCurrentCameraPose(const cv::Mat&) {} };
class System { public: enum eSensor { MONOCULAR = 0, STEREO = 1, RGBD = 2 }; void Reset() {} };

class Initializer {
public:
    Initializer(const Frame&, float = 1.0, int = 200) {}
    bool Initialize(const Frame&, const std::vector<int>&, cv::Mat&, cv::Mat&, std::vector<cv::Point3f>&, std::vector<bool>&) { return false; }
};

class PnPsolver {
public:
    PnPsolver(const Frame&, const std::vector<MapPoint*>&) {}
    void SetRansacParameters(double = 0.99, int = 8, int = 300, int = 4, float = 0.4, float = 5.991) {}
    cv::Mat iterate(int, bool& bNoMore, std::vector<bool>&, int&) { bNoMore = true; return cv::Mat(); }
};

}  // namespace ORB_SLAM2

// Stand-ins for the collaborators of the reference's src/Optimizer.cc (include/Frame.h, KeyFrame.h, MapPoint.h, Map.h, LoopClosing.h: unusable here -- they pull in
// DBoW2-on-OpenCV, ncnn, PCL, threads) carrying exactly the members Optimizer.cc touches, so that the reference's OWN Optimizer.cc and Converter.cc compile unmodified
// from where they lie (oracle/Makefile force-includes this header; it defines the real headers' include guards, which turns them into no-ops).  g2o's sparse
// Cholesky wrapper (solvers/linear_solver_eigen.h: Eigen's SimplicialLDLT internals) is replaced by a solver that refuses to solve: only the bundle adjustments and
// the essential-graph / Sim3 optimisations use it, and none of them is on the pinned path (PoseOptimization uses LinearSolverDense).
// Pinned through this library: every line of Optimizer::PoseOptimization (src/Optimizer.cc:239-451), Converter::toSE3Quat / toCvMat, and the g2o it drives.
// Not pinned: the other Optimizer functions (they compile, nothing calls them).  TEST INFRASTRUCTURE.
#pragma once
#define FRAME_H
#define KEYFRAME_H
#define MAPPOINT_H
#define MAP_H
#define LOOPCLOSING_H
#define G2O_LINEAR_SOLVER_EIGEN_H
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <Eigen/StdVector>
#include "Thirdparty/g2o/g2o/core/linear_solver.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"

using namespace std;      // the reference's headers do this at namespace scope and Optimizer.h relies on it (`map`, `set`)

namespace g2o {
template <typename MatrixType> class LinearSolverEigen : public LinearSolver<MatrixType> {
public:
    virtual bool init() { return true; }
    virtual bool solve(const SparseBlockMatrix<MatrixType>&, double*, double*) { return false; }
    void setBlockOrdering(bool) {}
};
}

namespace ORB_SLAM2 {

class KeyFrame;
class Map;

class MapPoint {
public:
    cv::Mat mWorldPos;                                     // 3x1 CV_32F
    bool mbBad = false;
    long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAGlobalForKF = 0, mnCorrectedByKF = 0, mnCorrectedReference = 0;
    cv::Mat mPosGBA;
    static std::mutex mGlobalMutex;
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    void SetWorldPos(const cv::Mat& Pos) { Pos.copyTo(mWorldPos); }
    bool isBad() { return mbBad; }
    std::map<KeyFrame*, size_t> GetObservations() { return std::map<KeyFrame*, size_t>(); }
    void EraseObservation(KeyFrame*) {}
    int GetIndexInKeyFrame(KeyFrame*) { return -1; }
    KeyFrame* GetReferenceKeyFrame() { return nullptr; }
    void UpdateNormalAndDepth() {}
};

class Frame {
public:
    int N = 0;
    cv::Mat mTcw;                                          // 4x4 CV_32F
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvInvLevelSigma2;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); }      // src/Frame.cc:274-278 (UpdatePoseMatrices only derives Rcw / tcw / Ow from it)
};

class KeyFrame {
public:
    long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAFixedForKF = 0, mnBAGlobalForKF = 0;
    cv::Mat mTcwGBA, Tcw, mK;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvInvLevelSigma2, mvLevelSigma2;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    std::vector<MapPoint*> mvpMapPoints;
    bool mbBad = false;
    bool isBad() { return mbBad; }
    cv::Mat GetPose() { return Tcw.clone(); }
    cv::Mat GetPoseInverse() { return Tcw.clone(); }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    void SetPose(const cv::Mat& T) { T.copyTo(Tcw); }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return std::vector<KeyFrame*>(); }
    std::vector<KeyFrame*> GetCovisiblesByWeight(const int&) { return std::vector<KeyFrame*>(); }
    std::set<KeyFrame*> GetLoopEdges() { return std::set<KeyFrame*>(); }
    KeyFrame* GetParent() { return nullptr; }
    bool hasChild(KeyFrame*) { return false; }
    int GetWeight(KeyFrame*) { return 0; }
    void EraseMapPointMatch(const size_t&) {}
    void EraseMapPointMatch(MapPoint*) {}
};

class Map {
public:
    std::mutex mMutexMapUpdate;
    std::vector<KeyFrame*> GetAllKeyFrames() { return std::vector<KeyFrame*>(); }
    std::vector<MapPoint*> GetAllMapPoints() { return std::vector<MapPoint*>(); }
    long unsigned int GetMaxKFid() { return 0; }
};

class LoopClosing {
public:
    typedef map<KeyFrame*, g2o::Sim3, std::less<KeyFrame*>, Eigen::aligned_allocator<std::pair<KeyFrame* const, g2o::Sim3> > > KeyFrameAndPose;    // include/LoopClosing.h:50-51 (whose pair<const KeyFrame*, ...> today's libstdc++ rejects)
};

}  // namespace ORB_SLAM2

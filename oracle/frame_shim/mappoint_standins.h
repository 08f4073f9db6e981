// Force-included when compiling the reference's OWN src/MapPoint.cc (oracle/_ref/libmappoint_ref.so): include/MapPoint.h is the REAL header; the headers it
// pulls in for its collaborators (KeyFrame.h, Frame.h, Map.h, ORBmatcher.h) become no-ops and the members MapPoint.cc touches are declared here.
// Pinned by this library: MapPoint::ComputeDistinctiveDescriptors (:242-307), PredictScale x2 (:385-418), the 0.8 / 1.2 invariance getters (:373-383),
// AddObservation.  Not pinned: UpdateNormalAndDepth and the Frame-based constructor (Mat / scalar arithmetic; SLAM state, not on the path).  TEST INFRASTRUCTURE.
#pragma once
#define KEYFRAME_H
#define FRAME_H
#define MAP_H
#define ORBMATCHER_H
#include <climits>
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include <opencv2/opencv.hpp>

using namespace std;

namespace cv {
inline MatExpr operator+(const Mat& a, const MatExpr& b) { MatExpr e; e.kind = MatExpr::ADD; e.a = a; e.b = b.eval(); return e; }      // UpdateNormalAndDepth (compiled, not pinned)
}

namespace ORB_SLAM2 {

class MapPoint;
class Map {
public:
    std::mutex mMutexPointCreation;
    void EraseMapPoint(MapPoint*) {}
};
struct ScaleOwner {      // what MapPoint.cc reads of a KeyFrame / Frame
    long unsigned int mnId = 0, mnFrameId = 0;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvuRight;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors, Ow;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
};
class KeyFrame : public ScaleOwner {
public:
    bool bad = false;
    bool isBad() { return bad; }
    void EraseMapPointMatch(const size_t&) {}
    void ReplaceMapPointMatch(const size_t&, MapPoint*) {}
};
class Frame : public ScaleOwner {};
class ORBmatcher {
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {      // src/ORBmatcher.cc:1649-1667 (pinned by liborbmatcher_ref.so): the population count of a ^ b
        int d = 0;
        for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a.ptr()[i] ^ b.ptr()[i]));
        return d;
    }
};

}  // namespace ORB_SLAM2

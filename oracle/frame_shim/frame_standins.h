// Force-included when compiling the reference's OWN src/Frame.cc (oracle/_ref/libframe_ref.so, recipe in oracle/Makefile): turns the headers Frame.h pulls in
// for its collaborators -- MapPoint.h, KeyFrame.h, ORBVocabulary.h, Tracking.h, Converter.h, ORBmatcher.h, Detector2D.h (DBoW2-on-OpenCV, g2o, ncnn, PCL: unusable
// here) -- into no-ops and declares, for each, exactly the members Frame.cc touches.  include/Frame.h and include/ORBextractor.h are the REAL headers.
// Restated here and therefore not pinned by this library: MapPoint::PredictScale and the 0.8 / 1.2 invariance getters (src/MapPoint.cc:373-418),
// Converter::toDescriptorVector, ORBmatcher::DescriptorDistance (only the stereo matcher, never run, uses it).  TEST INFRASTRUCTURE.
#pragma once
#define MAPPOINT_H
#define KEYFRAME_H
#define ORBVOCABULARY_H
#define TRACKING_H
#define CONVERTER_H
#define ORBMATCHER_H
#define DETECTOR2D_H
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
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;      // the reference's headers do this at namespace scope and Frame.h relies on it (`vector`)

typedef struct Object2D {          // include/Detector2D.h:30-36
    cv::Rect_<float> rect;
    float prob;
    std::string name;
    int id;
} Object2D;

namespace ORB_SLAM2 {

class Frame;
class KeyFrame {};

class MapPoint {
public:
    cv::Mat mWorldPos, mNormalVector;                       // 3x1 CV_32F
    float mfMinDistance = 0, mfMaxDistance = 0;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, Frame* pF);  // defined in the driver, after Frame is complete (src/MapPoint.cc:402-418)
};

class ORBVocabulary {
public:
    void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) {}      // pinned separately (libdbow2_ref.so)
};

class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& D) { std::vector<cv::Mat> v; for (int j = 0; j < D.rows; j++) v.push_back(D.row(j)); return v; }
};

class ORBmatcher {                 // what the (never executed) stereo matcher of Frame.cc refers to
public:
    static const int TH_LOW = 50, TH_HIGH = 100;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        int d = 0;
        for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a.ptr()[i] ^ b.ptr()[i]));
        return d;
    }
};

class Tracking;
class Detector2D {                 // include/Detector2D.h:42-66: the result members the Frame reads
public:
    std::vector<Object2D> mvObjects2D;
    bool mbHaveDynamicObjectForMapping = false;
    bool mbHaveDynamicObjectForRmDynamicFeature = false;
    std::vector<cv::Rect_<float> > mvPotentialDynamicBorderForMapping;
    std::vector<cv::Rect_<float> > mvPotentialDynamicBorderForRmDynamicFeature;
};

class Tracking {                   // include/Tracking.h:77-79
public:
    bool isDetectImageFinished() { return true; }           // the detector's results are planted before the Frame is built
    Detector2D* mpDetector2d = nullptr;
};

}  // namespace ORB_SLAM2

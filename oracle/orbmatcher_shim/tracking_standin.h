// Force-included when compiling the reference's src/Detector2D.cc: turns include/Tracking.h (which pulls in the whole system) into a no-op and provides the one
// member Detector2D.cc touches.  TEST INFRASTRUCTURE.
#pragma once
#define TRACKING_H
#include <cassert>
#include <string>
namespace ORB_SLAM2 {
class Tracking { public: bool mbDetectImageFinishedFlag = false; };
}

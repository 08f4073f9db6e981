// Force-included by tests/test_mirror_on_reference.py: the reference's src/Tracking.cc is compiled with the mirror classes of include/sgslam/ORBmatcher.h in place of
// ORBmatcher.h / ORBmatcher.cc and with Optimizer::PoseOptimization forwarded to include/sgslam/Optimizer.h's PoseOptimizationGPU -- the one-line change INTEGRATION.md
// describes for the call sites src/Tracking.cc:880 / :933 / :1314.  Everything else (Frame, MapPoint, the extractor) is the reference's own code.
#include "tracking_standins.h"
#define ORBMATCHER_H
#define OPTIMIZER_H
#include "sgslam/ORBmatcher.h"
#include "sgslam/Optimizer.h"
namespace ORB_SLAM2 {
class Frame;
class Optimizer {               // include/Optimizer.h:36-57: what src/Tracking.cc calls
public:
    template <class FrameT> static int PoseOptimization(FrameT* pFrame) { return PoseOptimizationGPU(pFrame); }
    static void GlobalBundleAdjustemnt(Map*, int = 5, bool* = NULL, const unsigned long = 0, const bool = true) {}
};
}

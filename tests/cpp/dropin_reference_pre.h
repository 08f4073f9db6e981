// Force-included by tests/test_dropin_reference.py when it compiles the REFERENCE'S OWN src/Tracking.cc and src/Frame.cc against this repository's mirror headers:
// the real include/ORBextractor.h and include/ORBmatcher.h are switched off by their include guards and include/sgslam/ORBextractor.h / ORBmatcher.h take their
// place, so every ORBextractor / ORBmatcher call site of the tracking thread (constructor, operator(), the getters, both SearchByProjection forms, SearchByBoW,
// the relocalisation search, SearchForInitialization, DescriptorDistance, TH_LOW / TH_HIGH) must resolve against the mirror classes as written in the reference.
// The rest of the reference (KeyFrame, Map, viewers ...) comes from the oracle's stand-ins; OpenCV from the oracle's cv shim (real-OpenCV API shape).
#include "tracking_standins.h"
#define ORBEXTRACTOR_H
#define ORBMATCHER_H
#include "sgslam/ORBextractor.h"
#include "sgslam/ORBmatcher.h"

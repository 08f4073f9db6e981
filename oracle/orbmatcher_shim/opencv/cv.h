// Stand-in for <opencv/cv.h> (include/ORBextractor.h includes it).  TEST INFRASTRUCTURE.
#pragma once
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/imgproc/imgproc.hpp>

// Stand-in for <opencv2/highgui/highgui.hpp> (nothing of it is used on the pinned path).  TEST INFRASTRUCTURE.
#pragma once
#include "../core/core.hpp"

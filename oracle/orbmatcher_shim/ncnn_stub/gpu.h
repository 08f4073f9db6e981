// Stand-in for ncnn's gpu.h (Vulkan instance management: nothing of it is called by Detector2D.cc).  TEST INFRASTRUCTURE.
#pragma once

// extract_dev.cuh -- device-side view of the extractor plan and small exact-arithmetic helpers.
#pragma once
#include <cuda_runtime.h>

#include "quadtree_core.h"
#include "sgs_common.h"

namespace sgs {

struct DevLevel {
    const uint8_t* img;        // frame 0 of this level (level 0 may alias the caller's device buffer)
    uint8_t* img_w;            // writable alias (levels >= 1 and the own level-0 staging buffer)
    uint8_t* blur;             // frame 0 of the blurred level
    int32_t w, h, pitch;       // pitch of img
    int32_t bpitch;            // pitch of blur
    int64_t fstride;           // bytes between frames of img
    int64_t bfstride;          // bytes between frames of blur
    uint32_t* cand;            // frame 0 candidate list of this level
    int32_t cand_cap;
    int32_t kp_cap;            // capacity of the per-level keypoint staging list
    int32_t kp_off;            // offset of this level inside a frame's staging block
    int32_t max_bx, max_by;
    QtGeom qt;
    float scale, patch_size;
    const short4* xtab;        // bilinear tables (levels >= 1)
    const short4* ytab;
};

struct DevPlan {
    DevLevel lv[kMaxLevels];
    int32_t nlevels;
    int32_t nframes;
    int32_t ini_th, min_th;
    int64_t cand_fstride;      // candidate words per frame (all levels)
    int32_t kp_stage_per_frame;// staging words per frame (sum of kp_cap)
    int32_t out_cap;           // keypoints per frame in the result arrays
    int32_t* cand_count;       // [nframes][nlevels]
    uint32_t* kp_stage;        // [nframes][kp_stage_per_frame] packed selected candidates
    int32_t* kp_stage_n;       // [nframes][nlevels]
    sgs_keypoint* out_kps;     // [nframes][out_cap]
    uint8_t* out_desc;         // [nframes][out_cap][32]
    int32_t* out_count;        // [nframes]
    int32_t* error_flag;       // device int: non-zero = capacity overflow somewhere (never expected)
    int32_t umax[kHalfPatch + 1];
};

// cvRound for float on the device: round-half-to-even
__device__ __forceinline__ int dev_cv_round(float v) { return __float2int_rn(v); }

// cv::fastAtan2 (degrees, [0,360)) with every product/sum individually rounded (the reference x86 build has no FMA).
__device__ __forceinline__ float dev_fast_atan2(float y, float x) {
    const float kRad2Deg = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, kRad2Deg);
    const float p3 = __fmul_rn(-0.3258083974640975f, kRad2Deg);
    const float p5 = __fmul_rn(0.1555786518463281f, kRad2Deg);
    const float p7 = __fmul_rn(-0.04432655554792128f, kRad2Deg);
    const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

}  // namespace sgs

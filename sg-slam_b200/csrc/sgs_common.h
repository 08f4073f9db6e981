// sgs_common.h -- shared host/device declarations of libsgs_cuda.so (product code; never includes oracle/).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sgs_abi.h"

#if defined(__CUDACC__)
#define SGS_HD __host__ __device__ __forceinline__
#else
#define SGS_HD inline
#endif

namespace sgs {

constexpr int kPatchSize = 31;       // PATCH_SIZE        src/ORBextractor.cc:73
constexpr int kHalfPatch = 15;       // HALF_PATCH_SIZE   :74
constexpr int kEdge = 19;            // EDGE_THRESHOLD    :75
constexpr int kMinBorder = kEdge - 3;  // minBorderX/Y    :774-775
constexpr int kMaxLevels = 16;
constexpr int kQtDepth = 12;         // quadtree path depth: supports level dimensions up to 4096 px

void set_error(const char* fmt, ...);

#define SGS_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::sgs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return SGS_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// One FAST cell = one cv::FAST call of the reference (src/ORBextractor.cc:790-830): view [x0,x1) x [y0,y1) in
// level coordinates; detection happens on its interior (3-px frame excluded).
struct FastCell {
    uint16_t x0, y0, x1, y1;
    uint8_t level;
    uint8_t ci, cj;   // cell row / column (candidate ordering key)
    uint8_t pad;
};

struct LevelGeom {
    int32_t w, h, pitch;        // level image size and device pitch (bytes)
    int32_t n_cols, n_rows, w_cell, h_cell;  // FAST cell grid (:785-788)
    int32_t max_bx, max_by;     // maxBorderX/Y (:776-777); minBorder is kMinBorder
    int32_t n_target;           // mnFeaturesPerLevel[level]
    int32_t n_ini;              // quadtree root count (:544)
    float h_x;                  // root strip width (:546)
    int32_t cand_cap;           // capacity of the candidate list of this level (exact upper bound)
    int32_t kp_cap;             // n_target + 3 (+ slack for n_ini), SURVEY Appendix E.5
    int64_t img_off;            // byte offset of frame 0 of this level inside the pyramid allocation
    int64_t frame_stride;       // bytes between consecutive frames of this level
    int64_t cand_off;           // element offset of this level's list inside one frame's candidate block
    int32_t cell_begin, cell_end;  // range in the cell table
    float scale;                // mvScaleFactor[level]
    float patch_size;           // (float)(int)(31 * scale)  (:838)
};

// Host-side plan: the tables ORBextractor's constructor computes plus everything derived from the image geometry.
struct OrbPlan {
    sgs_orb_params p;
    int width, height, nlevels;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> n_per_level;
    int umax[kHalfPatch + 1];
    std::vector<LevelGeom> lv;
    std::vector<FastCell> cells;
    // bilinear tables per level l >= 1 (source = level l-1): x then y entries, each (s, a0, a1, 0) as int16x4
    std::vector<std::vector<int16_t>> xtab, ytab;
    int max_kp_per_frame;       // sum of kp_cap
    int64_t cand_per_frame;     // sum of cand_cap
    int64_t pyr_bytes_per_frame;  // sum of h * pitch
};

// Builds the plan.  Returns SGS_OK or SGS_ERR_INVALID / SGS_ERR_UNSUPPORTED (with set_error).
int make_plan(const sgs_orb_params& p, int width, int height, OrbPlan* plan);

// cvRound on the host (round half to even), used for table construction only
int cv_round_f(float v);

}  // namespace sgs

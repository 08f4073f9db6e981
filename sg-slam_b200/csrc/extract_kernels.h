// extract_kernels.h -- launchers of the extractor kernels (extract_kernels.cu)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "extract_dev.cuh"

namespace sgs {
void launch_resize(const DevPlan& P, int level, cudaStream_t st);        // direct kernel (any scale factor)
bool resize_tile_supported(const DevPlan& P, int level);
void launch_resize_tile(const DevPlan& P, int level, cudaStream_t st);   // shared-memory tiled kernel (pyramid_kernel.cu)

// warp-per-cell FAST with TMA-staged tiles (fast_kernel.cu)
struct FastTmaMaps { CUtensorMap m[kMaxLevels]; };
struct FastLaunchPlan { int tp = 0, th = 0, list_cap = 0, warp_stride = 0; size_t smem_bytes = 0; };
FastLaunchPlan make_fast_launch_plan(const OrbPlan& PL);
cudaError_t configure_fast_smem(size_t smem_bytes);
void launch_fast_v2(const DevPlan& P, const FastTmaMaps& M, bool use_tma, const FastLaunchPlan& fp, const FastCell* d_cells, int ncells, int frame0, cudaStream_t st);
bool encode_level_map(CUtensorMap* out, const void* base, int w, int h, int pitch, int64_t fstride, int nframes, int box_w, int box_h);
void launch_quadtree(const DevPlan& P, int smem_key_cap, int node_cap, size_t smem_bytes, uint64_t* key_scratch, int64_t key_scratch_fstride,
                     const int64_t* d_key_scratch_off, cudaStream_t st);
cudaError_t configure_quadtree_smem(size_t smem_bytes);
size_t quadtree_node_bytes(int cap);
void launch_blur_all(const DevPlan& P, cudaStream_t st);          // every level in one launch
void launch_describe(const DevPlan& P, cudaStream_t st);
const char* last_error_cstr();
}  // namespace sgs

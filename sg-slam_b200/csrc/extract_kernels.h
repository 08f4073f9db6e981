// extract_kernels.h -- launchers of the extractor kernels (extract_kernels.cu)
#pragma once
#include <cuda_runtime.h>

#include "extract_dev.cuh"

namespace sgs {
void launch_resize(const DevPlan& P, int level, cudaStream_t st);
void launch_fast(const DevPlan& P, const FastCell* d_cells, int ncells, cudaStream_t st);
void launch_quadtree(const DevPlan& P, int smem_key_cap, int node_cap, size_t smem_bytes, uint64_t* key_scratch, int64_t key_scratch_fstride,
                     const int64_t* d_key_scratch_off, cudaStream_t st);
cudaError_t configure_quadtree_smem(size_t smem_bytes);
size_t quadtree_node_bytes(int cap);
void launch_blur(const DevPlan& P, int level, cudaStream_t st);
void launch_describe(const DevPlan& P, cudaStream_t st);
const char* last_error_cstr();
}  // namespace sgs

// device_ctx.h -- process-wide CUDA device bookkeeping shared by the C ABI files.
#pragma once
#include <cuda_runtime.h>
#include "../../include/speckv.h"
namespace speckv {
int device_count();                       // 0 when no CUDA device is usable
int current_sm_count();                   // SMs of the current device
speckv_status_t status_of(cudaError_t e); // cudaError_t -> speckv_status_t (clears the sticky error)
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st);   // stream-ordered, from a pool that keeps its memory; cudaFreeAsync releases
cudaError_t scratch_persistent(void** p, size_t bytes, cudaStream_t st, int tag = 0);   // per-(device, stream, tag) buffer kept across calls, zeroed when (re)allocated; never freed by the caller
void count_launch(unsigned n = 1);         // statistics: kernels launched by this library
void release_host_pipe();                 // frees the staging buffers of the *_host calls
}  // namespace speckv

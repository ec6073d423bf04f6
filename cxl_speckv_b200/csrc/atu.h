#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
namespace speckv {
cudaError_t launch_translate(const uint64_t* d_va, uint64_t* d_pa, size_t n, int sm_count, cudaStream_t st);
}

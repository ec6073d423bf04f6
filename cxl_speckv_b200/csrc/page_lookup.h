#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
namespace speckv {
struct KvPageDev {   // == KvPage (page_table.h) == speckv_page_t == the reference's KvPageHandle
    uint64_t virt_page_id;
    uint64_t phys_page_id;
    uint32_t page_size;
    uint32_t flags;
};
cudaError_t launch_page_lookup(const KvPageDev* d_pages, size_t num_pages, uint64_t va_base, const uint64_t* d_va,
                               uint64_t* d_pa, uint32_t* d_flags, size_t n, int sm_count, cudaStream_t st);
}  // namespace speckv

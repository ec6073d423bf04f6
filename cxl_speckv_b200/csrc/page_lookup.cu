// page_lookup.cu -- batched page-table lookup on the device.
//
// Replaces the per-call std::unordered_map probes behind one mutex of
//   SpeckvAllocator::access / is_in_l1_or_l2     host/src/speckv_allocator.cpp:54-74,105-113
//   CXLMemoryManager::translate_virtual_to_physical / is_in_cache
//                                                src/cxl_memory/cxl_memory_manager.cpp:106-128
// Both tables are dense in the virtual page number (pages of a handle are
// (h << 32) | (i << 12); the legacy manager bump-allocates from 0x1_0000_0000), so the
// device copy is a flat array of KvPage records indexed by (va - va_base) >> 12.
// Result: pa = entry.phys_page_id + (va & 0xFFF) and the entry's flag bits
// (bit0 L1, bit1 L2, bit2 compressed); an address outside the table, or whose
// entry is for a different page, gives pa = 0 / flags = 0 exactly like the
// reference's "not found" paths (:116 returns 0, access() returns nullptr).
// 8 B in, 12 B out and one 24 B gather per address: HBM/L2-bound integer work.
#include "device_ctx.h"
#include "page_lookup.h"

namespace speckv {

namespace {
__global__ void __launch_bounds__(256)
page_lookup_kernel(const KvPageDev* __restrict__ pages, size_t num_pages, uint64_t va_base,
                   const uint64_t* __restrict__ va, uint64_t* __restrict__ pa, uint32_t* __restrict__ flags, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t v = va[i];
        uint64_t out = 0;
        uint32_t fl = 0;
        if (v >= va_base) {
            const uint64_t idx = (v - va_base) >> 12;
            if (idx < num_pages) {
                const uint64_t virt = __ldg(&pages[idx].virt_page_id);
                if (virt == (v & ~0xFFFULL)) {
                    out = __ldg(&pages[idx].phys_page_id) + (v & 0xFFFULL);
                    fl = __ldg(&pages[idx].flags);
                }
            }
        }
        pa[i] = out;
        if (flags) flags[i] = fl;
    }
}
}  // namespace

cudaError_t launch_page_lookup(const KvPageDev* d_pages, size_t num_pages, uint64_t va_base, const uint64_t* d_va,
                               uint64_t* d_pa, uint32_t* d_flags, size_t n, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    page_lookup_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_pages, num_pages, va_base, d_va, d_pa, d_flags, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace speckv

// atu.cu -- batched address translation.
//
// FPGACacheEngine::translate_address (src/fpga_engine/cache_engine.cpp:118-140):
// on a TLB miss pa = 0x4000000000 + (va & 0xFFFFFFFFFFFF) (:132); on a hit the
// entry filled by that same formula is returned with the page offset re-attached
// (:126-128), i.e. the same value -- the TLB only changes latency on the FPGA.
// Pure integer streaming: 8 B in + 8 B out per address, two addresses per thread
// as one 128-bit load/store.
#include "atu.h"
#include "device_ctx.h"

namespace speckv {

namespace {
constexpr unsigned long long kPhysBase = 0x4000000000ULL;
constexpr unsigned long long kVaMask = 0xFFFFFFFFFFFFULL;

__global__ void __launch_bounds__(256)
translate_kernel(const uint64_t* __restrict__ va, uint64_t* __restrict__ pa, size_t n, bool vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        const size_t n2 = n >> 1;
        const ulonglong2* v2 = reinterpret_cast<const ulonglong2*>(va);
        ulonglong2* p2 = reinterpret_cast<ulonglong2*>(pa);
        for (size_t k = i; k < n2; k += stride) {
            ulonglong2 v = __ldg(v2 + k);
            v.x = kPhysBase + (v.x & kVaMask);
            v.y = kPhysBase + (v.y & kVaMask);
            p2[k] = v;
        }
        if ((n & 1) && i == 0) pa[n - 1] = kPhysBase + (va[n - 1] & kVaMask);
    } else {
        for (size_t k = i; k < n; k += stride) pa[k] = kPhysBase + (va[k] & kVaMask);
    }
}
}  // namespace

cudaError_t launch_translate(const uint64_t* d_va, uint64_t* d_pa, size_t n, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(d_va) | reinterpret_cast<uintptr_t>(d_pa)) & 15) == 0;
    const size_t work = vec_ok ? (n + 1) / 2 : n;
    size_t blocks = (work + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    translate_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_va, d_pa, n, vec_ok);
    count_launch();
    return cudaGetLastError();
}

}  // namespace speckv

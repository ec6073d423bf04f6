// atu_tlb.cu -- the stateful Address Translation Unit model, batched.
//
// AddressTranslationUnit (src/utils/address_translation.{h,cpp}) is a direct-mapped TLB of
// `tlb_size` entries in front of page_walk():
//   translate(va), address_translation.cpp:19-46
//     idx = (va >> 12) % tlb_size
//     hit  (valid && entry.vpage == va & ~0xFFF): ++hits,   return entry.ppage + (va & 0xFFF)
//     miss: ++misses, pp = page_walk(va) = 0x4000000000 + (va & 2^48-1)   (:85-90, takes the FULL va)
//           entry = {va & ~0xFFF, pp & ~0xFFF, valid}; return pp + (va & 0xFFF)   <- offset added twice
//   invalidate(va) :48-58 clears `valid` when entry.vpage matches (without testing valid);
//   invalidate_all :60-66.
// The result of a batch therefore depends on the ORDER of the addresses.  Sequential semantics
// are kept exactly: TLB set s is owned by one thread, which walks the batch in order and handles
// the addresses that map to its set (every thread reads the same va[i]: one broadcast load per
// step).  A stats/diagnostic path, not a bandwidth path; the stateless hot-path translate is atu.cu.
#include "../../include/speckv_ext.h"
#include "device_ctx.h"

namespace speckv {
namespace {

struct AtuState {
    uint64_t* vpage;
    uint64_t* ppage;
    uint8_t* valid;
    unsigned long long* counters;   // [0] hits, [1] misses
    uint32_t size;
};

__global__ void __launch_bounds__(128)
atu_translate_kernel(AtuState st, const uint64_t* __restrict__ va, uint64_t* __restrict__ pa, size_t n) {
    const uint32_t set = blockIdx.x * blockDim.x + threadIdx.x;
    if (set >= st.size) return;
    uint64_t vp = st.vpage[set], pp = st.ppage[set];
    bool valid = st.valid[set] != 0;
    unsigned long long hits = 0, misses = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t v = va[i];
        if ((uint32_t)((v >> 12) % st.size) != set) continue;
        const uint64_t page = v & ~0xFFFULL, off = v & 0xFFFULL;
        if (valid && vp == page) {
            ++hits;
            pa[i] = pp + off;
        } else {
            ++misses;
            const uint64_t walked = 0x4000000000ULL + (v & 0xFFFFFFFFFFFFULL);
            vp = page;
            pp = walked & ~0xFFFULL;
            valid = true;
            pa[i] = walked + off;
        }
    }
    st.vpage[set] = vp;
    st.ppage[set] = pp;
    st.valid[set] = valid ? 1 : 0;
    if (hits) atomicAdd(&st.counters[0], hits);
    if (misses) atomicAdd(&st.counters[1], misses);
}

__global__ void atu_invalidate_kernel(AtuState st, uint64_t va, int all) {
    const uint32_t set = blockIdx.x * blockDim.x + threadIdx.x;
    if (set >= st.size) return;
    if (all) {
        st.valid[set] = 0;
    } else {
        const uint64_t page = va & ~0xFFFULL;
        if ((uint32_t)((page >> 12) % st.size) == set && st.vpage[set] == page) st.valid[set] = 0;
    }
}

}  // namespace
}  // namespace speckv

using namespace speckv;

struct speckv_atu {
    AtuState st;
};

extern "C" {

speckv_status_t speckv_ext_atu_create(uint32_t tlb_size, speckv_atu_t** out_atu) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!out_atu || tlb_size == 0) return SPECKV_ERR_INVAL;
    speckv_atu* a = new speckv_atu();
    a->st.size = tlb_size;
    cudaError_t e = cudaMalloc((void**)&a->st.vpage, tlb_size * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&a->st.ppage, tlb_size * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&a->st.valid, tlb_size);
    if (e == cudaSuccess) e = cudaMalloc((void**)&a->st.counters, 16);
    if (e == cudaSuccess) e = cudaMemset(a->st.vpage, 0, tlb_size * 8);
    if (e == cudaSuccess) e = cudaMemset(a->st.ppage, 0, tlb_size * 8);
    if (e == cudaSuccess) e = cudaMemset(a->st.valid, 0, tlb_size);
    if (e == cudaSuccess) e = cudaMemset(a->st.counters, 0, 16);
    if (e != cudaSuccess) {
        speckv_ext_atu_destroy(a);
        return status_of(e);
    }
    *out_atu = a;
    return SPECKV_OK;
}

void speckv_ext_atu_destroy(speckv_atu_t* atu) {
    if (!atu) return;
    cudaFree(atu->st.vpage);
    cudaFree(atu->st.ppage);
    cudaFree(atu->st.valid);
    cudaFree(atu->st.counters);
    cudaGetLastError();
    delete atu;
}

speckv_status_t speckv_ext_atu_translate(speckv_atu_t* atu, const uint64_t* d_va, uint64_t* d_pa, size_t n,
                                         void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!atu || (n && (!d_va || !d_pa))) return SPECKV_ERR_INVAL;
    if (n == 0) return SPECKV_OK;
    atu_translate_kernel<<<(atu->st.size + 127) / 128, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(atu->st, d_va, d_pa, n);
    count_launch();
    return status_of(cudaGetLastError());
}

speckv_status_t speckv_ext_atu_invalidate(speckv_atu_t* atu, uint64_t va, int all, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!atu) return SPECKV_ERR_INVAL;
    atu_invalidate_kernel<<<(atu->st.size + 127) / 128, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(atu->st, va, all);
    count_launch();
    return status_of(cudaGetLastError());
}

speckv_status_t speckv_ext_atu_get_stats(speckv_atu_t* atu, uint64_t* hits, uint64_t* misses, int reset) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!atu || !hits || !misses) return SPECKV_ERR_INVAL;
    unsigned long long c[2];
    cudaError_t e = cudaMemcpy(c, atu->st.counters, 16, cudaMemcpyDeviceToHost);   // synchronises the device
    if (e == cudaSuccess && reset) e = cudaMemset(atu->st.counters, 0, 16);
    *hits = c[0];
    *misses = c[1];
    return status_of(e);
}

}  // extern "C"

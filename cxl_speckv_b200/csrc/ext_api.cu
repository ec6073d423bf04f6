// ext_api.cu -- the additive C ABI of libcxlspeckv.so (include/speckv_ext.h).
//
// Thin argument checking + dispatch onto the sm_100a kernels.  No torch types,
// no C++ types across the boundary.  There is no CPU fallback: without a usable
// CUDA device every entry point returns SPECKV_ERR_DRIVER.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/speckv_ext.h"
#include "atu.h"
#include "codec_math.cuh"
#include "device_ctx.h"
#include "kv_codec.h"

namespace speckv {

// ---- device context ---------------------------------------------------------------
static std::once_flag g_dev_once;
static int g_dev_count = 0;
static std::vector<int> g_sm_count;

static void probe_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    g_dev_count = n;
    g_sm_count.assign(n > 0 ? n : 0, 148);
    for (int d = 0; d < n; ++d) {
        int sm = 0;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, d) == cudaSuccess && sm > 0) g_sm_count[d] = sm;
    }
}

int device_count() {
    std::call_once(g_dev_once, probe_devices);
    return g_dev_count;
}

// Scratch for the codec calls (per-launch flag words) comes from a private stream-ordered pool that
// keeps its memory across synchronisations: the default pool releases everything at every sync, and
// the next cudaMallocAsync then pays a fresh physical allocation (measured: ~1 ms per first call).
static std::mutex g_pool_mu;
static std::vector<cudaMemPool_t> g_pools;

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    if (device_count() <= 0) return cudaErrorNoDevice;
    int d = 0;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) return e;
    if (d < 0 || d >= g_dev_count) return cudaErrorInvalidDevice;
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pools.empty()) g_pools.assign(g_dev_count, nullptr);
        if (!g_pools[d]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = d;
            e = cudaMemPoolCreate(&g_pools[d], &props);
            if (e != cudaSuccess) {
                g_pools[d] = nullptr;
                return e;
            }
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(g_pools[d], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = g_pools[d];
    }
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}

int current_sm_count() {
    if (device_count() <= 0) return 0;
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= g_dev_count) return 148;
    return g_sm_count[d];
}

speckv_status_t status_of(cudaError_t e) {
    if (e == cudaSuccess) return SPECKV_OK;
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) return SPECKV_ERR_NOMEM;
    if (e == cudaErrorInvalidValue) return SPECKV_ERR_INVAL;
    return SPECKV_ERR_DRIVER;
}

// ---- statistics ---------------------------------------------------------------------
static std::atomic<uint64_t> g_n_comp{0}, g_n_decomp{0}, g_n_xlate{0}, g_b_comp{0}, g_b_decomp{0};
std::atomic<uint64_t> g_kernel_launches{0};
void count_launch(unsigned n) { g_kernel_launches += n; }

static size_t elem_bytes(int dtype) { return dtype == DT_F32 ? 4 : 2; }

static bool valid_common(int dtype, size_t group_elems, size_t n_groups, size_t slot_bytes, int scheme) {
    if (dtype < 0 || dtype > 2 || scheme < 0 || scheme > 2) return false;
    if (group_elems >= (1ull << 31) || n_groups >= (1ull << 32)) return false;
    if (scheme == 0 && dtype == DT_F32) return false;  // FP16 scheme is a raw 16-bit passthrough
    if (slot_bytes % 16 != 0 || slot_bytes < speckv_ext_slot_bytes(group_elems, (speckv_comp_scheme_t)scheme)) return false;
    if (slot_bytes >= (1ull << 32)) return false;
    return true;
}

// ---- pipelined host-buffer path --------------------------------------------------------
struct HostPipe {
    static constexpr int kSlots = 3;
    std::mutex mu;
    int device = -1;
    cudaStream_t st[kSlots] = {};
    void* d_elems[kSlots] = {};
    void* d_payload[kSlots] = {};
    float* d_scales[kSlots] = {};
    uint32_t* d_comp[kSlots] = {};
    uint32_t* d_oel[kSlots] = {};
    size_t cap_elems = 0, cap_payload = 0, cap_groups = 0;

    void release() {
        for (int i = 0; i < kSlots; ++i) {
            if (d_elems[i]) cudaFree(d_elems[i]);
            if (d_payload[i]) cudaFree(d_payload[i]);
            if (d_scales[i]) cudaFree(d_scales[i]);
            if (d_comp[i]) cudaFree(d_comp[i]);
            if (d_oel[i]) cudaFree(d_oel[i]);
            d_elems[i] = d_payload[i] = nullptr;
            d_scales[i] = nullptr;
            d_comp[i] = d_oel[i] = nullptr;
        }
        cap_elems = cap_payload = cap_groups = 0;
    }

    cudaError_t ensure(size_t elems_bytes, size_t payload_bytes, size_t groups) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev != device) {
            release();
            for (int i = 0; i < kSlots; ++i) {
                if (st[i]) cudaStreamDestroy(st[i]);
                st[i] = nullptr;
            }
            device = dev;
        }
        for (int i = 0; i < kSlots; ++i)
            if (!st[i] && (e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (elems_bytes > cap_elems || payload_bytes > cap_payload || groups > cap_groups) {
            release();
            for (int i = 0; i < kSlots; ++i) {
                if ((e = cudaMalloc(&d_elems[i], elems_bytes)) != cudaSuccess) return e;
                if ((e = cudaMalloc(&d_payload[i], payload_bytes)) != cudaSuccess) return e;
                if ((e = cudaMalloc((void**)&d_scales[i], groups * sizeof(float))) != cudaSuccess) return e;
                if ((e = cudaMalloc((void**)&d_comp[i], groups * sizeof(uint32_t))) != cudaSuccess) return e;
                if ((e = cudaMalloc((void**)&d_oel[i], groups * sizeof(uint32_t))) != cudaSuccess) return e;
            }
            cap_elems = elems_bytes;
            cap_payload = payload_bytes;
            cap_groups = groups;
        }
        return cudaSuccess;
    }
};
static HostPipe g_pipe;

void release_host_pipe() {
    std::lock_guard<std::mutex> lk(g_pipe.mu);
    g_pipe.release();
}

// groups per chunk: aim at ~32 MiB of elements so H2D, kernel and D2H of neighbouring chunks overlap
static size_t chunk_groups(size_t group_bytes, size_t n_groups) {
    static const size_t target = [] {
        const char* e = std::getenv("SPECKV_HOST_CHUNK_MIB");
        const long v = e ? std::atol(e) : 0;
        return (size_t)(v > 0 ? v : 64) << 20;
    }();
    size_t c = group_bytes ? target / group_bytes : n_groups;
    if (c < 1) c = 1;
    if (c > n_groups) c = n_groups;
    return c ? c : 1;
}

// sum of payload sizes and sum of the per-group ratios original_size / compressed_size, where
// original_size = n * sizeof(float) exactly as the reference accounts it (cache_engine.cpp:49,72)
__global__ void __launch_bounds__(256)
ratio_kernel(const uint32_t* __restrict__ comp_bytes, size_t n_groups, double original_bytes, double* __restrict__ acc) {
    double c = 0.0, r = 0.0;
    for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < n_groups; g += (size_t)gridDim.x * 256) {
        const double cb = (double)comp_bytes[g];
        c += cb;
        if (cb > 0.0) r += original_bytes / cb;
    }
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], c);
        atomicAdd(&acc[1], r);
    }
}

}  // namespace speckv

using namespace speckv;

extern "C" {

speckv_status_t speckv_ext_ratio_stats(const uint32_t* d_comp_bytes, size_t n_groups, size_t group_elems,
                                       double* out_total_comp_bytes, double* out_mean_ratio, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!d_comp_bytes || !out_total_comp_bytes || !out_mean_ratio) return SPECKV_ERR_INVAL;
    *out_total_comp_bytes = 0.0;
    *out_mean_ratio = 0.0;
    if (n_groups == 0) return SPECKV_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    double* d_acc = nullptr;
    cudaError_t e = scratch_alloc((void**)&d_acc, 16, st);
    if (e != cudaSuccess) return status_of(e);
    cudaMemsetAsync(d_acc, 0, 16, st);
    size_t blocks = (n_groups + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    ratio_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_comp_bytes, n_groups, (double)group_elems * 4.0, d_acc);
    count_launch();
    double h[2] = {0.0, 0.0};
    e = cudaMemcpyAsync(h, d_acc, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(d_acc, st);
    *out_total_comp_bytes = h[0];
    *out_mean_ratio = h[1] / (double)n_groups;
    return status_of(e);
}

int speckv_ext_device_count(void) { return device_count(); }

const char* speckv_ext_version(void) { return "cxl-speckv-b200 0.1 (sm_100a)"; }

size_t speckv_ext_slot_bytes(size_t group_elems, speckv_comp_scheme_t scheme) {
    size_t b = (scheme == SPECKV_COMP_INT8) ? group_elems : 2 * group_elems;
    return (b + 15) / 16 * 16;
}

speckv_status_t speckv_ext_compress(const void* d_in, speckv_dtype_t dtype, size_t group_elems, size_t n_groups,
                                    void* d_payload, size_t slot_bytes, float* d_scales, uint32_t* d_comp_bytes,
                                    speckv_comp_scheme_t scheme, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_groups, slot_bytes, scheme)) return SPECKV_ERR_INVAL;
    if (n_groups == 0) return SPECKV_OK;
    if (!d_payload || !d_scales || !d_comp_bytes || (!d_in && group_elems)) return SPECKV_ERR_INVAL;
    if (reinterpret_cast<uintptr_t>(d_payload) & 15) return SPECKV_ERR_INVAL;
    CodecArgs a;
    a.in = d_in;
    a.payload = d_payload;
    a.scales = d_scales;
    a.comp_bytes = d_comp_bytes;
    a.slot_bytes = slot_bytes;
    a.group_elems = (uint32_t)group_elems;
    a.n_groups = (uint32_t)n_groups;
    a.dtype = dtype;
    a.scheme = scheme;
    a.sm_count = current_sm_count();
    cudaError_t e = launch_compress(a, static_cast<cudaStream_t>(cuda_stream));
    if (e == cudaSuccess) {
        g_n_comp += n_groups;
        g_b_comp += (uint64_t)n_groups * group_elems * elem_bytes(dtype);
    }
    return status_of(e);
}

speckv_status_t speckv_ext_decompress(const void* d_payload, size_t slot_bytes, const float* d_scales,
                                      const uint32_t* d_comp_bytes, size_t group_elems, size_t n_groups,
                                      speckv_dtype_t dtype, void* d_out, uint32_t* d_out_elems,
                                      speckv_comp_scheme_t scheme, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_groups, slot_bytes, scheme)) return SPECKV_ERR_INVAL;
    if (n_groups == 0) return SPECKV_OK;
    if (!d_payload || !d_scales || !d_comp_bytes || (!d_out && group_elems)) return SPECKV_ERR_INVAL;
    if (reinterpret_cast<uintptr_t>(d_payload) & 15) return SPECKV_ERR_INVAL;
    CodecArgs a;
    a.out = d_out;
    a.payload = const_cast<void*>(d_payload);
    a.scales = const_cast<float*>(d_scales);
    a.comp_bytes = const_cast<uint32_t*>(d_comp_bytes);
    a.out_elems = d_out_elems;
    a.slot_bytes = slot_bytes;
    a.group_elems = (uint32_t)group_elems;
    a.n_groups = (uint32_t)n_groups;
    a.dtype = dtype;
    a.scheme = scheme;
    a.sm_count = current_sm_count();
    cudaError_t e = launch_decompress(a, static_cast<cudaStream_t>(cuda_stream));
    if (e == cudaSuccess) {
        g_n_decomp += n_groups;
        g_b_decomp += (uint64_t)n_groups * group_elems * elem_bytes(dtype);
    }
    return status_of(e);
}

speckv_status_t speckv_ext_decompress_indexed(const void* d_payload, size_t slot_bytes, const float* d_scales,
                                              const uint32_t* d_comp_bytes, const uint32_t* d_block_index,
                                              size_t n_requests, size_t group_elems, speckv_dtype_t dtype, void* d_out,
                                              uint32_t* d_out_elems, speckv_comp_scheme_t scheme, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_requests, slot_bytes, scheme)) return SPECKV_ERR_INVAL;
    if (n_requests == 0) return SPECKV_OK;
    if (!d_payload || !d_scales || !d_comp_bytes || !d_block_index || (!d_out && group_elems)) return SPECKV_ERR_INVAL;
    if (reinterpret_cast<uintptr_t>(d_payload) & 15) return SPECKV_ERR_INVAL;
    CodecArgs a;
    a.out = d_out;
    a.payload = const_cast<void*>(d_payload);
    a.scales = const_cast<float*>(d_scales);
    a.comp_bytes = const_cast<uint32_t*>(d_comp_bytes);
    a.out_elems = d_out_elems;
    a.src_index = d_block_index;
    a.slot_bytes = slot_bytes;
    a.group_elems = (uint32_t)group_elems;
    a.n_groups = (uint32_t)n_requests;
    a.dtype = dtype;
    a.scheme = scheme;
    a.sm_count = current_sm_count();
    cudaError_t e = launch_decompress(a, static_cast<cudaStream_t>(cuda_stream));
    if (e == cudaSuccess) {
        g_n_decomp += n_requests;
        g_b_decomp += (uint64_t)n_requests * group_elems * elem_bytes(dtype);
    }
    return status_of(e);
}

speckv_status_t speckv_ext_compress_gather(const void* d_cache, const uint32_t* d_block_table, speckv_dtype_t dtype,
                                           size_t group_elems, size_t n_groups, void* d_payload, size_t slot_bytes,
                                           float* d_scales, uint32_t* d_comp_bytes, speckv_comp_scheme_t scheme,
                                           void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_groups, slot_bytes, scheme) || scheme == SPECKV_COMP_FP16) return SPECKV_ERR_INVAL;
    if (n_groups == 0) return SPECKV_OK;
    if (!d_payload || !d_scales || !d_comp_bytes || !d_block_table || (!d_cache && group_elems)) return SPECKV_ERR_INVAL;
    if (reinterpret_cast<uintptr_t>(d_payload) & 15) return SPECKV_ERR_INVAL;
    CodecArgs a;
    a.in = d_cache;
    a.elem_index = d_block_table;
    a.payload = d_payload;
    a.scales = d_scales;
    a.comp_bytes = d_comp_bytes;
    a.slot_bytes = slot_bytes;
    a.group_elems = (uint32_t)group_elems;
    a.n_groups = (uint32_t)n_groups;
    a.dtype = dtype;
    a.scheme = scheme;
    a.sm_count = current_sm_count();
    cudaError_t e = launch_compress(a, static_cast<cudaStream_t>(cuda_stream));
    if (e == cudaSuccess) {
        g_n_comp += n_groups;
        g_b_comp += (uint64_t)n_groups * group_elems * elem_bytes(dtype);
    }
    return status_of(e);
}

speckv_status_t speckv_ext_decompress_scatter(const void* d_payload, size_t slot_bytes, const float* d_scales,
                                              const uint32_t* d_comp_bytes, const uint32_t* d_src_index,
                                              const uint32_t* d_block_table, size_t n_requests, size_t group_elems,
                                              speckv_dtype_t dtype, void* d_cache, uint32_t* d_out_elems,
                                              speckv_comp_scheme_t scheme, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_requests, slot_bytes, scheme) || scheme == SPECKV_COMP_FP16) return SPECKV_ERR_INVAL;
    if (n_requests == 0) return SPECKV_OK;
    if (!d_payload || !d_scales || !d_comp_bytes || !d_block_table || (!d_cache && group_elems)) return SPECKV_ERR_INVAL;
    if (reinterpret_cast<uintptr_t>(d_payload) & 15) return SPECKV_ERR_INVAL;
    CodecArgs a;
    a.out = d_cache;
    a.elem_index = d_block_table;
    a.payload = const_cast<void*>(d_payload);
    a.scales = const_cast<float*>(d_scales);
    a.comp_bytes = const_cast<uint32_t*>(d_comp_bytes);
    a.out_elems = d_out_elems;
    a.src_index = d_src_index;
    a.slot_bytes = slot_bytes;
    a.group_elems = (uint32_t)group_elems;
    a.n_groups = (uint32_t)n_requests;
    a.dtype = dtype;
    a.scheme = scheme;
    a.sm_count = current_sm_count();
    cudaError_t e = launch_decompress(a, static_cast<cudaStream_t>(cuda_stream));
    if (e == cudaSuccess) {
        g_n_decomp += n_requests;
        g_b_decomp += (uint64_t)n_requests * group_elems * elem_bytes(dtype);
    }
    return status_of(e);
}

speckv_status_t speckv_ext_compress_host(const void* h_in, speckv_dtype_t dtype, size_t group_elems, size_t n_groups,
                                         void* h_payload, size_t slot_bytes, float* h_scales, uint32_t* h_comp_bytes,
                                         speckv_comp_scheme_t scheme) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_groups, slot_bytes, scheme)) return SPECKV_ERR_INVAL;
    if (n_groups == 0) return SPECKV_OK;
    if (!h_in || !h_payload || !h_scales || !h_comp_bytes) return SPECKV_ERR_INVAL;
    const size_t gbytes = group_elems * elem_bytes(dtype);
    const size_t cg = chunk_groups(gbytes > slot_bytes ? gbytes : slot_bytes, n_groups);
    std::lock_guard<std::mutex> lk(g_pipe.mu);
    cudaError_t e = g_pipe.ensure(cg * gbytes + 16, cg * slot_bytes, cg);
    if (e != cudaSuccess) return status_of(e);
    size_t chunk = 0;
    for (size_t g0 = 0; g0 < n_groups; g0 += cg, ++chunk) {
        const size_t ng = (n_groups - g0 < cg) ? n_groups - g0 : cg;
        const int k = (int)(chunk % HostPipe::kSlots);
        cudaStream_t st = g_pipe.st[k];
        if ((e = cudaMemcpyAsync(g_pipe.d_elems[k], (const char*)h_in + g0 * gbytes, ng * gbytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        CodecArgs a;
        a.in = g_pipe.d_elems[k];
        a.payload = g_pipe.d_payload[k];
        a.scales = g_pipe.d_scales[k];
        a.comp_bytes = g_pipe.d_comp[k];
        a.slot_bytes = slot_bytes;
        a.group_elems = (uint32_t)group_elems;
        a.n_groups = (uint32_t)ng;
        a.dtype = dtype;
        a.scheme = scheme;
        a.sm_count = current_sm_count();
        if ((e = launch_compress(a, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync((char*)h_payload + g0 * slot_bytes, g_pipe.d_payload[k], ng * slot_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(h_scales + g0, g_pipe.d_scales[k], ng * sizeof(float), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(h_comp_bytes + g0, g_pipe.d_comp[k], ng * sizeof(uint32_t), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    }
    for (int k = 0; k < HostPipe::kSlots; ++k) {
        cudaError_t e2 = cudaStreamSynchronize(g_pipe.st[k]);
        if (e == cudaSuccess) e = e2;
    }
    if (e == cudaSuccess) {
        g_n_comp += n_groups;
        g_b_comp += (uint64_t)n_groups * gbytes;
    }
    return status_of(e);
}

speckv_status_t speckv_ext_decompress_host(const void* h_payload, size_t slot_bytes, const float* h_scales,
                                           const uint32_t* h_comp_bytes, size_t group_elems, size_t n_groups,
                                           speckv_dtype_t dtype, void* h_out, uint32_t* h_out_elems,
                                           speckv_comp_scheme_t scheme) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!valid_common(dtype, group_elems, n_groups, slot_bytes, scheme)) return SPECKV_ERR_INVAL;
    if (n_groups == 0) return SPECKV_OK;
    if (!h_payload || !h_scales || !h_comp_bytes || !h_out) return SPECKV_ERR_INVAL;
    const size_t gbytes = group_elems * elem_bytes(dtype);
    const size_t cg = chunk_groups(gbytes > slot_bytes ? gbytes : slot_bytes, n_groups);
    std::lock_guard<std::mutex> lk(g_pipe.mu);
    cudaError_t e = g_pipe.ensure(cg * gbytes + 16, cg * slot_bytes, cg);
    if (e != cudaSuccess) return status_of(e);
    size_t chunk = 0;
    for (size_t g0 = 0; g0 < n_groups; g0 += cg, ++chunk) {
        const size_t ng = (n_groups - g0 < cg) ? n_groups - g0 : cg;
        const int k = (int)(chunk % HostPipe::kSlots);
        cudaStream_t st = g_pipe.st[k];
        if ((e = cudaMemcpyAsync(g_pipe.d_payload[k], (const char*)h_payload + g0 * slot_bytes, ng * slot_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(g_pipe.d_scales[k], h_scales + g0, ng * sizeof(float), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(g_pipe.d_comp[k], h_comp_bytes + g0, ng * sizeof(uint32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        CodecArgs a;
        a.out = g_pipe.d_elems[k];
        a.payload = g_pipe.d_payload[k];
        a.scales = g_pipe.d_scales[k];
        a.comp_bytes = g_pipe.d_comp[k];
        a.out_elems = g_pipe.d_oel[k];
        a.slot_bytes = slot_bytes;
        a.group_elems = (uint32_t)group_elems;
        a.n_groups = (uint32_t)ng;
        a.dtype = dtype;
        a.scheme = scheme;
        a.sm_count = current_sm_count();
        if ((e = launch_decompress(a, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync((char*)h_out + g0 * gbytes, g_pipe.d_elems[k], ng * gbytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if (h_out_elems && (e = cudaMemcpyAsync(h_out_elems + g0, g_pipe.d_oel[k], ng * sizeof(uint32_t), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    }
    for (int k = 0; k < HostPipe::kSlots; ++k) {
        cudaError_t e2 = cudaStreamSynchronize(g_pipe.st[k]);
        if (e == cudaSuccess) e = e2;
    }
    if (e == cudaSuccess) {
        g_n_decomp += n_groups;
        g_b_decomp += (uint64_t)n_groups * gbytes;
    }
    return status_of(e);
}

void* speckv_ext_host_alloc(size_t bytes) {
    if (device_count() <= 0) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void speckv_ext_host_free(void* p) {
    if (p && device_count() > 0) cudaFreeHost(p);
}

speckv_status_t speckv_ext_translate(const uint64_t* d_va, uint64_t* d_pa, size_t n, void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (n == 0) return SPECKV_OK;
    if (!d_va || !d_pa) return SPECKV_ERR_INVAL;
    cudaError_t e = launch_translate(d_va, d_pa, n, current_sm_count(), static_cast<cudaStream_t>(cuda_stream));
    if (e == cudaSuccess) g_n_xlate += n;
    return status_of(e);
}

void speckv_ext_get_stats(speckv_ext_stats_t* out) {
    if (!out) return;
    out->total_compressions = g_n_comp.load();
    out->total_decompressions = g_n_decomp.load();
    out->total_translations = g_n_xlate.load();
    out->bytes_in_compress = g_b_comp.load();
    out->bytes_out_decompress = g_b_decomp.load();
    out->kernel_launches = g_kernel_launches.load();
}

void speckv_ext_reset_stats(void) {
    g_n_comp = 0;
    g_n_decomp = 0;
    g_n_xlate = 0;
    g_b_comp = 0;
    g_b_decomp = 0;
    g_kernel_launches = 0;
}

}  // extern "C"

// c_api.cu -- the eight frozen entry points of libcxlspeckv.so (include/speckv.h).
//
// Behaviour follows host/src/speckv_c_api.cpp:13-121 call for call: process-global
// state behind one mutex, the same status codes for the same misuse, handles
// restarting at 1 after finalize -> init.  The reference's SpeckvDriver (an ioctl
// client of /dev/speckv0, host/src/speckv_driver.cpp) is replaced by a CUDA device.
#include <fcntl.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/speckv_ext.h"
#include "device_ctx.h"
#include "page_lookup.h"
#include "page_table.h"

namespace speckv {

struct PrefetchRecord {            // == SpeckvPrefetchReq + tokens, speckv_driver.hpp:17-24
    uint32_t req_id;
    uint16_t layer;
    uint32_t cur_pos, depth_k;
    std::vector<int32_t> tokens;
};

struct Runtime {
    PageTable table;
    int fd = -1;                   // the opened dev_path, when it is a filesystem path
    int cuda_device = -1;          // -1: no CUDA device (host bookkeeping only; setters fail like a dead ioctl)
    uint32_t prefetch_depth = 4;   // SpeculativePrefetcher default, speculative_prefetcher.h:36
    int scheme = SPECKV_COMP_INT8_DELTA_RLE;
    std::deque<PrefetchRecord> prefetch_log;   // bounded: 16 outstanding, speculative_prefetcher.cpp:168-171
    uint64_t prefetch_total = 0;
};

static std::mutex g_mutex;
static std::unique_ptr<Runtime> g_rt;

Runtime* runtime_locked() { return g_rt.get(); }
std::mutex& runtime_mutex() { return g_mutex; }

}  // namespace speckv

using namespace speckv;

extern "C" {

speckv_status_t speckv_init(const char* dev_path) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_rt) return SPECKV_ERR_GENERAL;            // already initialised, speckv_c_api.cpp:16-18
    if (!dev_path) return SPECKV_ERR_INVAL;
    auto rt = std::make_unique<Runtime>();
    const int ndev = device_count();
    if (std::strncmp(dev_path, "cuda", 4) == 0 && (dev_path[4] == '\0' || dev_path[4] == ':')) {
        int ord = dev_path[4] == ':' ? std::atoi(dev_path + 5) : 0;
        if (ord < 0 || ord >= ndev) return SPECKV_ERR_DRIVER;   // "driver not ok", speckv_c_api.cpp:22-24
        rt->cuda_device = ord;
    } else {
        // the reference opens the path O_RDWR and throws when that fails -> ERR_GENERAL
        // (speckv_driver.cpp:11-16, speckv_c_api.cpp:29-31)
        rt->fd = ::open(dev_path, O_RDWR);
        if (rt->fd < 0) return SPECKV_ERR_GENERAL;
        if (ndev > 0) {
            const char* env = std::getenv("SPECKV_CUDA_DEVICE");
            int ord = env ? std::atoi(env) : 0;
            if (ord < 0 || ord >= ndev) {
                ::close(rt->fd);
                return SPECKV_ERR_DRIVER;
            }
            rt->cuda_device = ord;
        }
    }
    if (rt->cuda_device >= 0 && cudaSetDevice(rt->cuda_device) != cudaSuccess) {
        cudaGetLastError();
        if (rt->fd >= 0) ::close(rt->fd);
        return SPECKV_ERR_DRIVER;
    }
    g_rt = std::move(rt);
    return SPECKV_OK;
}

void speckv_finalize(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt) return;
    if (g_rt->fd >= 0) ::close(g_rt->fd);
    if (g_rt->cuda_device >= 0) release_host_pipe();
    g_rt.reset();
}

speckv_status_t speckv_alloc(size_t bytes, const speckv_alloc_hint_t* hint, speckv_handle_t* out_handle) {
    (void)hint;                                      // ignored by the reference too, speckv_c_api.cpp:50
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt || !out_handle) return SPECKV_ERR_INVAL;
    *out_handle = g_rt->table.alloc(bytes);
    return SPECKV_OK;
}

speckv_status_t speckv_free(speckv_handle_t handle) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt) return SPECKV_ERR_INVAL;
    g_rt->table.free(handle);                        // unknown handle: OK, speckv_allocator.cpp:42
    return SPECKV_OK;
}

speckv_status_t speckv_access(speckv_handle_t handle, uint64_t offset_bytes, size_t length_bytes, void** out_gpu_ptr) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt || !out_gpu_ptr) return SPECKV_ERR_INVAL;
    bool fetched = false;
    const uint64_t addr = g_rt->table.access(handle, offset_bytes, length_bytes, &fetched);
    if (!addr) return SPECKV_ERR_GENERAL;            // *out_gpu_ptr untouched, speckv_c_api.cpp:76-79
    *out_gpu_ptr = reinterpret_cast<void*>(addr);
    return SPECKV_OK;
}

speckv_status_t speckv_prefetch(uint32_t req_id, uint16_t layer, uint32_t cur_pos, uint32_t depth_k,
                                const int32_t* recent_tokens, uint32_t history_len) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt || !recent_tokens || history_len == 0) return SPECKV_ERR_INVAL;
    PrefetchRecord r;
    r.req_id = req_id;
    r.layer = layer;
    r.cur_pos = cur_pos;
    r.depth_k = depth_k;
    r.tokens.assign(recent_tokens, recent_tokens + history_len);
    g_rt->prefetch_log.push_back(std::move(r));
    if (g_rt->prefetch_log.size() > 16) g_rt->prefetch_log.pop_front();
    ++g_rt->prefetch_total;
    return SPECKV_OK;                                // the reference discards the driver's status, speckv_allocator.cpp:89
}

speckv_status_t speckv_set_prefetch_depth(uint32_t depth_k) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt) return SPECKV_ERR_INVAL;
    if (g_rt->cuda_device < 0) return SPECKV_ERR_DRIVER;   // no device behind the handle: like a failed ioctl
    g_rt->prefetch_depth = depth_k;
    return SPECKV_OK;
}

speckv_status_t speckv_set_compression_scheme(speckv_comp_scheme_t scheme) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt) return SPECKV_ERR_INVAL;
    if (g_rt->cuda_device < 0) return SPECKV_ERR_DRIVER;
    if ((int)scheme < 0 || (int)scheme > 2) return SPECKV_ERR_DRIVER;  // the device has no such mode
    g_rt->scheme = (int)scheme;
    return SPECKV_OK;
}

speckv_status_t speckv_ext_page_table_export(speckv_handle_t handle, speckv_page_t* d_pages, size_t capacity,
                                             size_t* out_count, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_rt || !out_count) return SPECKV_ERR_INVAL;
    if (g_rt->cuda_device < 0) return SPECKV_ERR_DRIVER;
    KvAllocation* a = g_rt->table.find(handle);
    if (!a) return SPECKV_ERR_GENERAL;
    *out_count = a->pages.size();
    const size_t n = a->pages.size() < capacity ? a->pages.size() : capacity;
    if (n == 0) return SPECKV_OK;
    if (!d_pages) return SPECKV_ERR_INVAL;
    static_assert(sizeof(KvPage) == sizeof(speckv_page_t) && sizeof(KvPage) == 24, "page record layout");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t e = cudaMemcpyAsync(d_pages, a->pages.data(), n * sizeof(KvPage), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the host vector may change after we return
    return status_of(e);
}

speckv_status_t speckv_ext_page_lookup(const speckv_page_t* d_pages, size_t num_pages, uint64_t va_base,
                                       const uint64_t* d_va, uint64_t* d_pa, uint32_t* d_flags, size_t n,
                                       void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (n == 0) return SPECKV_OK;
    if (!d_va || !d_pa || (!d_pages && num_pages)) return SPECKV_ERR_INVAL;
    return status_of(launch_page_lookup(reinterpret_cast<const KvPageDev*>(d_pages), num_pages, va_base, d_va, d_pa,
                                        d_flags, n, current_sm_count(), static_cast<cudaStream_t>(cuda_stream)));
}

}  // extern "C"

// request_route.cu -- routing of prefetch requests to the rank that owns the predicted block.
//
// The reference's prefetcher enqueues PrefetchRequest records and "issues a DMA" per record
// (src/prefetcher/speculative_prefetcher.cpp:57-66,162-172).  With the stored blocks sharded over
// the GPUs of a box, every rank all-gathers the fixed-size request tables
// and keeps the requests for blocks it owns.  This kernel does that selection on the device:
// owner filter, stable compaction into a block-index list, request count in device memory -- so
// the decode that follows (speckv_ext_decompress_routed) needs no host round trip.
#include "../../include/speckv_ext.h"
#include "device_ctx.h"

namespace speckv {
namespace {

constexpr int kRouteThreads = 1024;

__global__ void __launch_bounds__(kRouteThreads)
route_requests_kernel(const speckv_prefetch_request_t* __restrict__ tables, uint32_t n_tables, uint32_t cap,
                      uint32_t n_blocks, uint32_t per_rank, uint32_t rank, uint32_t* __restrict__ block_index,
                      uint32_t* __restrict__ request_index, uint32_t* __restrict__ count) {
    __shared__ uint32_t warp_sum[kRouteThreads / 32];
    __shared__ uint32_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t t = 0; t < n_tables; ++t) {
        const speckv_prefetch_request_t* tab = tables + (size_t)t * (cap + 1);
        const uint32_t valid = (uint32_t)min((unsigned long long)tab[0].virtual_addr, (unsigned long long)cap);
        for (uint32_t i0 = 0; i0 < valid; i0 += kRouteThreads) {   // uniform bounds: every thread takes part in the barriers
            const uint32_t i = i0 + tid;
            bool mine = false;
            uint32_t local = 0;
            if (i < valid) {
                const uint32_t b = tab[1 + i].predicted_token_id % n_blocks;
                const uint32_t owner = b / per_rank;
                mine = owner == rank;
                local = b - owner * per_rank;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (lane == 0) warp_sum[wid] = (uint32_t)__popc(bal);
            __syncthreads();
            uint32_t before = 0, total = 0;
            for (int w = 0; w < kRouteThreads / 32; ++w) {
                const uint32_t v = warp_sum[w];
                if (w < (int)wid) before += v;
                total += v;
            }
            const uint32_t base = s_base;
            if (mine) {
                const uint32_t o = base + before + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                block_index[o] = local;
                if (request_index) request_index[o] = t * cap + i;
            }
            __syncthreads();
            if (tid == 0) s_base = base + total;
            __syncthreads();
        }
    }
    if (tid == 0) *count = s_base;
}

}  // namespace
}  // namespace speckv

using namespace speckv;

extern "C" speckv_status_t speckv_ext_route_requests(const speckv_prefetch_request_t* d_tables, uint32_t n_tables,
                                                     uint32_t table_capacity, uint32_t n_blocks_total, uint32_t world,
                                                     uint32_t rank, uint32_t* d_block_index, uint32_t* d_request_index,
                                                     uint32_t* d_count, void* cuda_stream) {
    static_assert(sizeof(speckv_prefetch_request_t) == 32, "PrefetchRequest layout");
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!d_count || world == 0 || rank >= world || n_blocks_total == 0) return SPECKV_ERR_INVAL;
    if ((n_tables && table_capacity) && (!d_tables || !d_block_index)) return SPECKV_ERR_INVAL;
    const uint32_t per_rank = (n_blocks_total + world - 1) / world;
    route_requests_kernel<<<1, kRouteThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_tables, n_tables, table_capacity, n_blocks_total, per_rank, rank, d_block_index, d_request_index, d_count);
    count_launch();
    return status_of(cudaGetLastError());
}

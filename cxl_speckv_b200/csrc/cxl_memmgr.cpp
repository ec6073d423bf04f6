// cxl_memmgr.cpp -- the address map of the reference's CXLMemoryManager
// (src/cxl_memory/cxl_memory_manager.cpp:8-128) behind speckv_ext_memmgr_*.
//
// The reference keeps an unordered_map of heap-allocated MemoryPage objects under two mutexes and answers
// one address per call.  Its virtual addresses are bump-allocated from 0x1_0000_0000 in whole pages and never
// reused, so the map is dense in (va - base) / page: here it is a flat host array of 24-byte page records
// (speckv_page_t, the layout speckv_ext_page_lookup reads), appended by allocate() and copied to the device on
// export; translate_virtual_to_physical / is_in_cache for a batch of addresses are then ONE page_lookup launch
// (pa = entry.phys + offset, flags bit0 = L1, bit1 = L2; unknown address -> 0 like :116).
// Kept quirks: the L1 fallback test counts ALLOCATIONS, not pages (:37-39 with :295-316, the tier lists hold
// base addresses only); deallocate(va) drops the page entry at va alone (:81-104); nothing is ever reused.
// The residency side of the class (access tracking, promote / demote, LRU) is speckv_ext_policy_* (tier_policy.cu);
// set_tier() carries its decisions into this table.
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/speckv_ext.h"
#include "device_ctx.h"

namespace speckv {
namespace {
constexpr uint64_t kVaBase = 0x100000000ULL;                                                   // :18
constexpr uint64_t kPaBase[3] = {0x8000000000ULL, 0x10000000000ULL, 0x20000000000ULL};       // :19-21
constexpr uint32_t kNoTier = 255;
}  // namespace

struct MemMgr {
    std::mutex mu;
    uint64_t cap[3] = {0, 0, 0};
    uint64_t next_va = kVaBase, next_pa[3] = {kPaBase[0], kPaBase[1], kPaBase[2]}, n_alloc[3] = {0, 0, 0};
    std::vector<speckv_page_t> pages;   // virt_page_id == 0: no entry
    std::vector<uint8_t> tier, base;
};
}  // namespace speckv

struct speckv_memmgr {
    speckv::MemMgr m;
};

using namespace speckv;

static uint32_t flags_of(uint32_t tier) { return tier == 0 ? 1u : tier == 1 ? 2u : 0u; }

extern "C" {

speckv_status_t speckv_ext_memmgr_create(uint64_t l1_bytes, uint64_t l2_bytes, uint64_t l3_bytes, speckv_memmgr_t** out) {
    if (!out) return SPECKV_ERR_INVAL;
    speckv_memmgr* h = new (std::nothrow) speckv_memmgr();
    if (!h) return SPECKV_ERR_NOMEM;
    h->m.cap[0] = l1_bytes;
    h->m.cap[1] = l2_bytes;
    h->m.cap[2] = l3_bytes;
    *out = h;
    return SPECKV_OK;
}

void speckv_ext_memmgr_destroy(speckv_memmgr_t* h) { delete h; }

speckv_status_t speckv_ext_memmgr_allocate(speckv_memmgr_t* h, uint64_t size_bytes, uint32_t layer_id, int preferred_tier,
                                           uint64_t* out_va, int* out_tier) {
    (void)layer_id;
    if (!h || !out_va || preferred_tier < 0 || preferred_tier > 2) return SPECKV_ERR_INVAL;
    MemMgr& m = h->m;
    std::lock_guard<std::mutex> lk(m.mu);
    const uint64_t npages = (size_bytes + SPECKV_PAGE_BYTES - 1) / SPECKV_PAGE_BYTES, bytes = npages * SPECKV_PAGE_BYTES;
    int tier = preferred_tier;
    if (tier == 0 && m.n_alloc[0] * SPECKV_PAGE_BYTES + bytes > m.cap[0]) tier = 2;     // :37-39
    const uint64_t va = m.next_va, pa = m.next_pa[tier];
    const uint64_t first = (va - kVaBase) / SPECKV_PAGE_BYTES;
    try {
        m.pages.resize(first + npages);
        m.tier.resize(first + npages, (uint8_t)kNoTier);
        m.base.resize(first + npages, 0);
    } catch (...) {
        return SPECKV_ERR_NOMEM;
    }
    m.next_pa[tier] += bytes;
    m.n_alloc[tier] += 1;
    for (uint64_t i = 0; i < npages; ++i) {                                             // :62-75
        speckv_page_t& e = m.pages[first + i];
        e.virt_page_id = va + i * SPECKV_PAGE_BYTES;
        e.phys_page_id = pa + i * SPECKV_PAGE_BYTES;
        e.page_size = SPECKV_PAGE_BYTES;
        e.flags = flags_of((uint32_t)tier);
        m.tier[first + i] = (uint8_t)tier;
        m.base[first + i] = i == 0;
    }
    m.next_va += bytes;                                                                 // :77
    *out_va = va;
    if (out_tier) *out_tier = tier;
    return SPECKV_OK;
}

speckv_status_t speckv_ext_memmgr_deallocate(speckv_memmgr_t* h, uint64_t va) {
    if (!h) return SPECKV_ERR_INVAL;
    MemMgr& m = h->m;
    std::lock_guard<std::mutex> lk(m.mu);
    if (va < kVaBase || (va - kVaBase) % SPECKV_PAGE_BYTES) return SPECKV_OK;           // not a key of the page table: no-op, :84-85
    const uint64_t i = (va - kVaBase) / SPECKV_PAGE_BYTES;
    if (i >= m.pages.size() || m.pages[i].virt_page_id == 0) return SPECKV_OK;
    if (m.base[i] && m.n_alloc[m.tier[i]]) m.n_alloc[m.tier[i]] -= 1;
    std::memset(&m.pages[i], 0, sizeof(speckv_page_t));
    m.tier[i] = (uint8_t)kNoTier;
    m.base[i] = 0;
    return SPECKV_OK;
}

speckv_status_t speckv_ext_memmgr_set_tier(speckv_memmgr_t* h, uint64_t va, size_t n_pages, int tier) {
    if (!h || tier < 0 || tier > 2) return SPECKV_ERR_INVAL;
    MemMgr& m = h->m;
    std::lock_guard<std::mutex> lk(m.mu);
    if (va < kVaBase) return SPECKV_ERR_GENERAL;
    const uint64_t first = (va - kVaBase) / SPECKV_PAGE_BYTES;
    if (first + n_pages > m.pages.size()) return SPECKV_ERR_GENERAL;
    for (uint64_t i = first; i < first + n_pages; ++i) {
        if (m.pages[i].virt_page_id == 0) continue;
        m.tier[i] = (uint8_t)tier;
        m.pages[i].flags = flags_of((uint32_t)tier);
    }
    return SPECKV_OK;
}

speckv_status_t speckv_ext_memmgr_translate_host(speckv_memmgr_t* h, uint64_t va, uint64_t* out_pa, int* out_tier) {
    if (!h || !out_pa) return SPECKV_ERR_INVAL;
    MemMgr& m = h->m;
    std::lock_guard<std::mutex> lk(m.mu);
    *out_pa = 0;                                                                        // :116
    if (out_tier) *out_tier = (int)kNoTier;
    const uint64_t page_addr = va & ~(uint64_t)(SPECKV_PAGE_BYTES - 1);
    if (page_addr < kVaBase) return SPECKV_OK;
    const uint64_t i = (page_addr - kVaBase) / SPECKV_PAGE_BYTES;
    if (i >= m.pages.size() || m.pages[i].virt_page_id == 0) return SPECKV_OK;
    *out_pa = m.pages[i].phys_page_id + (va - page_addr);
    if (out_tier) *out_tier = m.tier[i];
    return SPECKV_OK;
}

speckv_status_t speckv_ext_memmgr_export(speckv_memmgr_t* h, speckv_page_t* d_pages, size_t capacity, size_t* out_count,
                                         uint64_t* out_va_base, void* cuda_stream) {
    if (!h || !out_count) return SPECKV_ERR_INVAL;
    MemMgr& m = h->m;
    std::lock_guard<std::mutex> lk(m.mu);
    *out_count = m.pages.size();
    if (out_va_base) *out_va_base = kVaBase;
    const size_t n = m.pages.size() < capacity ? m.pages.size() : capacity;
    if (n == 0 || !d_pages) return SPECKV_OK;
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t e = cudaMemcpyAsync(d_pages, m.pages.data(), n * sizeof(speckv_page_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the host table may change as soon as the lock is released
    return e == cudaSuccess ? SPECKV_OK : SPECKV_ERR_DRIVER;
}

}  // extern "C"

// ref_shim.cpp -- extern "C" entry points into the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; the reference translation units
// it is linked with are compiled where they lie under $REF (/root/reference)
// by oracle/Makefile into oracle/_ref/ (git-ignored, never copied into the repo).
// It exposes the reference's C++ software model (FPGACacheEngine,
// AddressTranslationUnit, LSTMPredictor, SpeculativePrefetcher,
// CXLMemoryManager) on raw buffers so the C restatement (speckv_oracle.c) and
// the CUDA path can be checked against it, and so bench.py can time the
// reference CPU path (`cpu_baseline.kind == "reference"`).
//
// `private` is re-defined to `public` for the reference headers only so that
// stage-level functions (quantize_to_int8 etc.) and the LSTM weights can be
// reached; no reference code is altered.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define private public
#include "fpga_engine/cache_engine.h"
#include "utils/address_translation.h"
#include "prefetcher/lstm_predictor.h"
#include "prefetcher/speculative_prefetcher.h"
#include "cxl_memory/cxl_memory_manager.h"
#undef private

using cxlspeckv::FPGACacheEngine;

extern "C" {

// ---- FPGACacheEngine (src/fpga_engine/cache_engine.cpp) --------------------
void* ref_engine_new() { return new FPGACacheEngine(); }
void ref_engine_free(void* e) { delete static_cast<FPGACacheEngine*>(e); }

// compress(): cache_engine.cpp:40-82.  Returns payload bytes; writes at most cap.
size_t ref_compress(void* e, const float* x, size_t n, float* scale, uint8_t* out, size_t cap,
                    size_t* original_size) {
    std::vector<float> v(x, x + n);
    auto c = static_cast<FPGACacheEngine*>(e)->compress(v, /*num_tokens=*/0, /*hidden_dim=*/0, 0);
    *scale = c.scale_factor;
    if (original_size) *original_size = c.original_size;
    size_t w = c.rle_data.size() < cap ? c.rle_data.size() : cap;
    if (w) std::memcpy(out, c.rle_data.data(), w);
    return c.compressed_size;
}

// decompress(): cache_engine.cpp:84-116.  Returns element count; writes at most cap.
size_t ref_decompress(void* e, float scale, const uint8_t* rle, size_t bytes, float* out, size_t cap) {
    FPGACacheEngine::CompressedData c;
    c.scale_factor = scale;
    c.rle_data.assign(reinterpret_cast<const int8_t*>(rle), reinterpret_cast<const int8_t*>(rle) + bytes);
    c.original_size = 0;
    c.compressed_size = bytes;
    auto v = static_cast<FPGACacheEngine*>(e)->decompress(c, 0, 0);
    size_t w = v.size() < cap ? v.size() : cap;
    if (w) std::memcpy(out, v.data(), w * sizeof(float));
    return v.size();
}

// stage-level access (private members, reached through the #define above)
float ref_scale(void* e, const float* x, size_t n) {
    std::vector<float> v(x, x + n);
    return static_cast<FPGACacheEngine*>(e)->compute_scale_factor(v);
}
void ref_quantize(void* e, const float* x, size_t n, float scale, int8_t* q) {
    std::vector<float> v(x, x + n);
    auto r = static_cast<FPGACacheEngine*>(e)->quantize_to_int8(v, scale);
    if (n) std::memcpy(q, r.data(), n);
}
void ref_delta_encode(void* e, const int8_t* q, size_t n, int8_t* d) {
    std::vector<int8_t> v(q, q + n);
    auto r = static_cast<FPGACacheEngine*>(e)->delta_encode(v);
    if (n) std::memcpy(d, r.data(), n);
}
uint64_t ref_engine_translate(void* e, uint64_t va) {
    return static_cast<FPGACacheEngine*>(e)->translate_address(va);
}
void ref_engine_stats(void* e, size_t* n_comp, size_t* n_decomp, double* avg_ratio, double* gbps) {
    auto s = static_cast<FPGACacheEngine*>(e)->get_statistics();
    *n_comp = s.total_compressions;
    *n_decomp = s.total_decompressions;
    *avg_ratio = s.avg_compression_ratio;
    *gbps = s.throughput_gbps;
}
double ref_engine_layer_ratio(void* e, uint32_t layer) {
    return static_cast<FPGACacheEngine*>(e)->get_compression_ratio(layer);
}

// ---- fp16/bf16 batch forms used as the CPU baseline ------------------------
static inline float widen(uint16_t b, int dtype) {
    if (dtype == 1) { uint32_t u = uint32_t(b) << 16; float f; std::memcpy(&f, &u, 4); return f; }
    _Float16 h; std::memcpy(&h, &b, 2); return float(h);
}
static inline uint16_t narrow(float v, int dtype) {
    if (dtype == 1) {
        uint32_t u; std::memcpy(&u, &v, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x0040u);
        u += 0x7fffu + ((u >> 16) & 1u);
        return uint16_t(u >> 16);
    }
    _Float16 h = _Float16(v); uint16_t o; std::memcpy(&o, &h, 2); return o;
}

// Round trip (compress then decompress) of n_groups independent groups through
// the reference engine, one engine instance per thread.  Groups are split
// contiguously over `threads` std::threads.  dtype: 0 fp16, 1 bf16.
int ref_roundtrip_batch(const uint16_t* in, int dtype, size_t group_elems, size_t n_groups,
                        uint8_t* payload, size_t slot_bytes, float* scales, uint32_t* comp_bytes,
                        uint16_t* out, int threads, int do_compress, int do_decompress) {
    if (threads < 1) threads = 1;
    auto work = [&](size_t g0, size_t g1) {
        FPGACacheEngine eng;
        std::vector<float> v(group_elems);
        for (size_t g = g0; g < g1; ++g) {
            if (do_compress) {
                const uint16_t* src = in + g * group_elems;
                for (size_t i = 0; i < group_elems; ++i) v[i] = widen(src[i], dtype);
                auto c = eng.compress(v, 0, 0, 0);
                scales[g] = c.scale_factor;
                comp_bytes[g] = uint32_t(c.compressed_size);
                size_t w = c.rle_data.size() < slot_bytes ? c.rle_data.size() : slot_bytes;
                if (w) std::memcpy(payload + g * slot_bytes, c.rle_data.data(), w);
            }
            if (do_decompress) {
                FPGACacheEngine::CompressedData c;
                c.scale_factor = scales[g];
                const int8_t* p = reinterpret_cast<const int8_t*>(payload + g * slot_bytes);
                c.rle_data.assign(p, p + comp_bytes[g]);
                c.original_size = 0;
                c.compressed_size = comp_bytes[g];
                auto y = eng.decompress(c, 0, 0);
                uint16_t* dst = out + g * group_elems;
                size_t m = y.size() < group_elems ? y.size() : group_elems;
                for (size_t i = 0; i < m; ++i) dst[i] = narrow(y[i], dtype);
            }
        }
    };
    if (threads == 1) { work(0, n_groups); return 0; }
    std::vector<std::thread> th;
    size_t per = (n_groups + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t g0 = std::min(n_groups, size_t(t) * per), g1 = std::min(n_groups, size_t(t + 1) * per);
        if (g0 < g1) th.emplace_back(work, g0, g1);
    }
    for (auto& t : th) t.join();
    return 0;
}

// ---- AddressTranslationUnit (src/utils/address_translation.cpp) ------------
void* ref_atu_new(size_t tlb_size) { return new cxlspeckv::AddressTranslationUnit(tlb_size); }
void ref_atu_free(void* a) { delete static_cast<cxlspeckv::AddressTranslationUnit*>(a); }
uint64_t ref_atu_translate(void* a, uint64_t va) {
    return static_cast<cxlspeckv::AddressTranslationUnit*>(a)->translate(va);
}
void ref_atu_invalidate(void* a, uint64_t va) { static_cast<cxlspeckv::AddressTranslationUnit*>(a)->invalidate(va); }
void ref_atu_invalidate_all(void* a) { static_cast<cxlspeckv::AddressTranslationUnit*>(a)->invalidate_all(); }
void ref_atu_stats(void* a, size_t* hits, size_t* misses) {
    auto s = static_cast<cxlspeckv::AddressTranslationUnit*>(a)->get_statistics();
    *hits = s.hits; *misses = s.misses;
}

// ---- LSTMPredictor (src/prefetcher/lstm_predictor.cpp) ----------------------
// The constructor draws weights from the process-global rand(); srand(seed)
// first so the caller can reproduce them (glibc default state == srand(1)).
void* ref_lstm_new(unsigned seed, size_t vocab, size_t emb, size_t hidden, size_t layers, size_t hist) {
    srand(seed);
    return new cxlspeckv::LSTMPredictor(vocab, emb, hidden, layers, hist);
}
void ref_lstm_free(void* p) { delete static_cast<cxlspeckv::LSTMPredictor*>(p); }
void ref_lstm_weights(void* p, float* emb, float* wout) {
    auto* l = static_cast<cxlspeckv::LSTMPredictor*>(p);
    std::memcpy(emb, l->embedding_weights_.data(), l->embedding_weights_.size() * sizeof(float));
    std::memcpy(wout, l->output_weights_.data(), l->output_weights_.size() * sizeof(float));
}
size_t ref_lstm_model_size(void* p) { return static_cast<cxlspeckv::LSTMPredictor*>(p)->get_model_size(); }
size_t ref_lstm_predict(void* p, const uint32_t* hist, size_t n_hist, size_t k, uint32_t* ids, float* conf) {
    std::vector<uint32_t> h(hist, hist + n_hist);
    auto r = static_cast<cxlspeckv::LSTMPredictor*>(p)->predict_top_k(h, k);
    for (size_t i = 0; i < r.size(); ++i) { ids[i] = r[i].first; conf[i] = r[i].second; }
    return r.size();
}

// ---- SpeculativePrefetcher (src/prefetcher/speculative_prefetcher.cpp) ------
// prefetch(): :25-82.  Returns the number of requests; fields are written
// column-wise.  The prefetcher owns a memory manager with nothing allocated,
// so no predicted block is "already cached" (is_in_cache false, :51-54).
struct RefPrefetcher {
    cxlspeckv::CXLMemoryManager mm;
    cxlspeckv::SpeculativePrefetcher pf;
    RefPrefetcher(size_t depth, size_t hist) : mm(), pf(&mm, depth, hist) {}
};
void* ref_prefetcher_new(unsigned seed, size_t depth, size_t hist) {
    srand(seed);
    return new RefPrefetcher(depth, hist);
}
void ref_prefetcher_free(void* p) { delete static_cast<RefPrefetcher*>(p); }
void ref_prefetcher_weights(void* p, float* emb, float* wout) {
    ref_lstm_weights(static_cast<RefPrefetcher*>(p)->pf.predictor_.get(), emb, wout);
}
size_t ref_prefetcher_prefetch(void* p, const uint32_t* hist, size_t n_hist, uint32_t layer, size_t depth,
                               uint64_t* va, uint32_t* layer_out, uint32_t* tok, float* conf) {
    std::vector<uint32_t> h(hist, hist + n_hist);
    auto r = static_cast<RefPrefetcher*>(p)->pf.prefetch(h, layer, depth);
    for (size_t i = 0; i < r.size(); ++i) {
        va[i] = r[i].virtual_addr; layer_out[i] = r[i].layer_id;
        tok[i] = r[i].predicted_token_id; conf[i] = r[i].confidence;
    }
    return r.size();
}
// adaptive depth: update_prediction_accuracy :99-120, get_adaptive_depth :122-124, set_prefetch_depth :144-147
void ref_prefetcher_feedback(void* p, int was_correct) {
    static_cast<RefPrefetcher*>(p)->pf.update_prediction_accuracy(0, was_correct != 0);
}
size_t ref_prefetcher_adaptive_depth(void* p) { return static_cast<RefPrefetcher*>(p)->pf.get_adaptive_depth(); }
void ref_prefetcher_set_depth(void* p, size_t d) { static_cast<RefPrefetcher*>(p)->pf.set_prefetch_depth(d); }
uint64_t ref_kv_address(void* p, uint32_t req, uint32_t layer, uint32_t pos) {
    return static_cast<RefPrefetcher*>(p)->pf.compute_kv_address(req, layer, pos);
}

// the prefetcher's own memory manager, populated: allocate :28-82 (residency filter of prefetch(), :51-54)
uint64_t ref_prefetcher_mm_allocate(void* p, size_t bytes, uint32_t layer, int tier) {
    return static_cast<RefPrefetcher*>(p)->mm.allocate(bytes, layer, static_cast<cxlspeckv::MemoryTier>(tier));
}
// handle_misprediction :84-97, get_statistics :126-135, reset_statistics :137-142, is_already_prefetched :174-185
void ref_prefetcher_mispredict(void* p, uint32_t actual, const uint32_t* predicted, size_t n) {
    static_cast<RefPrefetcher*>(p)->pf.handle_misprediction(actual, std::vector<uint32_t>(predicted, predicted + n));
}
void ref_prefetcher_stats(void* p, uint64_t* counters3, double* rates3, int reset) {
    auto& pf = static_cast<RefPrefetcher*>(p)->pf;
    const auto s = pf.get_statistics();
    counters3[0] = s.total_prefetches; counters3[1] = s.successful_prefetches; counters3[2] = s.mispredictions;
    rates3[0] = s.hit_rate; rates3[1] = s.precision; rates3[2] = s.avg_prediction_latency_us;
    if (reset) pf.reset_statistics();
}
int ref_prefetcher_is_outstanding(void* p, uint64_t va) {
    return static_cast<RefPrefetcher*>(p)->pf.is_already_prefetched(va) ? 1 : 0;
}
size_t ref_prefetcher_queue(void* p, uint64_t* va, uint32_t* tok, size_t cap) {   // outstanding_prefetches_, oldest first
    auto q = static_cast<RefPrefetcher*>(p)->pf.outstanding_prefetches_;
    size_t n = 0;
    while (!q.empty() && n < cap) { va[n] = q.front().virtual_addr; tok[n] = q.front().predicted_token_id; ++n; q.pop(); }
    return n;
}

// ---- CXLMemoryManager page table (src/cxl_memory/cxl_memory_manager.cpp) ----
void* ref_mm_new() { return new cxlspeckv::CXLMemoryManager(); }
void ref_mm_free(void* m) { delete static_cast<cxlspeckv::CXLMemoryManager*>(m); }
uint64_t ref_mm_allocate(void* m, size_t bytes, uint32_t layer, int tier) {
    return static_cast<cxlspeckv::CXLMemoryManager*>(m)->allocate(bytes, layer, static_cast<cxlspeckv::MemoryTier>(tier));
}
uint64_t ref_mm_translate(void* m, uint64_t va) {
    return static_cast<cxlspeckv::CXLMemoryManager*>(m)->translate_virtual_to_physical(va);
}
int ref_mm_is_in_cache(void* m, uint64_t va, int tier) {
    return static_cast<cxlspeckv::CXLMemoryManager*>(m)->is_in_cache(va, static_cast<cxlspeckv::MemoryTier>(tier)) ? 1 : 0;
}
// residency policy entry points (:130-258).  promote_to_l1 is only safe to call while L1 has room:
// its eviction path re-locks page_table_mutex_ (evict_l1_lru -> demote_to_l3) and never returns.
void ref_mm_touch(void* m, uint64_t va) { static_cast<cxlspeckv::CXLMemoryManager*>(m)->update_access_tracking(va); }
int ref_mm_is_hot(void* m, uint64_t va) { return static_cast<cxlspeckv::CXLMemoryManager*>(m)->is_hot_page(va) ? 1 : 0; }
int ref_mm_promote(void* m, uint64_t va) { return static_cast<cxlspeckv::CXLMemoryManager*>(m)->promote_to_l1(va) ? 1 : 0; }
int ref_mm_demote(void* m, uint64_t va) { return static_cast<cxlspeckv::CXLMemoryManager*>(m)->demote_to_l3(va) ? 1 : 0; }
void ref_mm_release(void* m, uint64_t va) { static_cast<cxlspeckv::CXLMemoryManager*>(m)->deallocate(va); }
int ref_mm_tier(void* m, uint64_t va) {
    auto* mm = static_cast<cxlspeckv::CXLMemoryManager*>(m);
    for (int t = 0; t < 3; ++t)
        if (mm->is_in_cache(va, static_cast<cxlspeckv::MemoryTier>(t))) return t;
    return 255;
}
size_t ref_mm_lru(void* m, uint64_t* out, size_t cap) {   // l1_lru_list_, least recently used first
    auto* mm = static_cast<cxlspeckv::CXLMemoryManager*>(m);
    size_t k = 0;
    for (uint64_t va : mm->l1_lru_list_) {
        if (k < cap) out[k] = va;
        ++k;
    }
    return k;
}
// {l1_hits, l1_misses, l2_hits, l2_misses, l3_accesses, migrations_l1_to_l3, migrations_l3_to_l1}, rates
void ref_mm_stats(void* m, uint64_t* counters7, double* rates2) {
    auto st = static_cast<cxlspeckv::CXLMemoryManager*>(m)->get_statistics();
    counters7[0] = st.l1_hits; counters7[1] = st.l1_misses; counters7[2] = st.l2_hits; counters7[3] = st.l2_misses;
    counters7[4] = st.l3_accesses; counters7[5] = st.migrations_l1_to_l3; counters7[6] = st.migrations_l3_to_l1;
    rates2[0] = st.l1_hit_rate; rates2[1] = st.l2_hit_rate;
}

}  // extern "C"

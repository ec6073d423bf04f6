// prefetch_score.cu -- batched speculative-prefetch scoring.
//
// Replaces LSTMPredictor::predict_top_k (src/prefetcher/lstm_predictor.cpp:40-94) and the
// request emission of SpeculativePrefetcher::prefetch (speculative_prefetcher.cpp:25-82) for
// a whole batch of sequences per launch (the reference scores one sequence per call on the CPU:
// 17.8 ms each).
//
// What the reference's "LSTM" actually computes (and what is reproduced, step for step):
//   per token t of the 16-token window, per layer (the SAME embedded input and state both times,
//   weights ignored, lstm_predictor.cpp:116-147):
//       g = sum_{j<64} emb[tok][j] * 0.1f        sequential fp32 (mul, then add)      :138-140
//       c = 0.5f*c + 0.5f*tanh(g);  h = 0.5f*tanh(c)                                  :143-145
//   so all 128 hidden units carry the same value h.
//   logit_v = sum_{j<128} h * W[v][j]            sequential fp32 (mul, then add)      :166-173
//   p = softmax(logits) (max-subtracted), top-k by p                                  :176-188,:72-92
//   request i: va = (req << 32) | (layer << 16) | (i + 1)         speculative_prefetcher.cpp:153-160
// Kernels:
//   hidden_kernel   one warp per sequence: gather the window's embedding rows, then 16 x 2 cell updates on scalars
//   logits_kernel   CTA = 128 vocabulary rows staged in shared memory (W read once per batch
//                   tile: 16.4 MB total, L2 resident), each thread keeps the sequential
//                   j-order of the reference for BB sequences at a time
//   topk_kernel     CTA per sequence: one pass for the max and per-thread best-k lists, k block arg-max rounds over
//                   the list heads, one pass for the sum of exp
// tanh and the k reported exp values are evaluated in fp64 and rounded once to fp32 (glibc's float versions are
// within 1 ulp of that), the 32000 terms of the softmax sum with expf; parity is therefore "same top-k ids, confidences within 1e-7", not bitwise
// (SURVEY.md section 8a A11).
#include <cfloat>
#include <mutex>

#include "../../include/speckv_ext.h"
#include "device_ctx.h"

namespace speckv {

namespace {

struct Predictor {
    float* d_emb = nullptr;
    float* d_wout = nullptr;
    float* d_hidden = nullptr;   // [max_batch]
    float* d_logits = nullptr;   // [max_batch][vocab]
    uint32_t vocab = 0, emb_dim = 0, hidden = 0, layers = 0, hist_len = 0;
    size_t max_batch = 0;
    int device = -1;
};
std::mutex g_pred_mu;
Predictor g_pred;

__global__ void hidden_kernel(const uint32_t* __restrict__ tokens, uint32_t batch, uint32_t hist_len,
                              const float* __restrict__ emb, uint32_t vocab, uint32_t emb_dim, uint32_t hidden,
                              uint32_t layers, float* __restrict__ h_out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    float h = 0.0f, c = 0.0f;
    const uint32_t nj = emb_dim < hidden ? emb_dim : hidden;   // `j < input_dim && j < hidden_dim`, :138
    for (uint32_t t = 0; t < hist_len; ++t) {
        const uint32_t tok = tokens[(size_t)b * hist_len + t];
        float g = 0.0f;
        if (tok < vocab) {   // embed_token: out-of-vocabulary ids embed to zeros, :149-160
            const float* e = emb + (size_t)tok * emb_dim;
            for (uint32_t j = 0; j < nj; ++j) g = __fadd_rn(g, __fmul_rn(__ldg(e + j), 0.1f));
        }
        const float tg = (float)tanh((double)g);
        for (uint32_t l = 0; l < layers; ++l) {
            c = __fadd_rn(__fmul_rn(0.5f, c), __fmul_rn(0.5f, tg));
            h = __fmul_rn(0.5f, (float)tanh((double)c));
        }
    }
    h_out[b] = h;
}

// Same recurrence, one WARP per sequence: the lanes first gather the window's embedding rows into shared
// memory (hist_len x nj independent loads in flight instead of one dependent load per addition), then lane 0
// runs the reference's sequential sums over them.  84 us -> a few us for 256 sequences.
constexpr int kHiddenWarps = 4;
__global__ void __launch_bounds__(kHiddenWarps * 32)
hidden_warp_kernel(const uint32_t* __restrict__ tokens, uint32_t batch, uint32_t hist_len,
                   const float* __restrict__ emb, uint32_t vocab, uint32_t emb_dim, uint32_t hidden, uint32_t layers,
                   float* __restrict__ h_out) {
    extern __shared__ float se[];   // [kHiddenWarps][hist_len][nj]
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * kHiddenWarps + warp;
    if (b >= batch) return;
    const uint32_t nj = emb_dim < hidden ? emb_dim : hidden;
    float* my = se + (size_t)warp * hist_len * nj;
    for (uint32_t idx = lane; idx < hist_len * nj; idx += 32) {
        const uint32_t t = idx / nj, j = idx - t * nj;
        const uint32_t tok = tokens[(size_t)b * hist_len + t];
        my[idx] = tok < vocab ? __ldg(emb + (size_t)tok * emb_dim + j) : 0.0f;   // out-of-vocabulary ids embed to zeros
    }
    __syncwarp();
    if (lane != 0) return;
    float h = 0.0f, c = 0.0f;
    for (uint32_t t = 0; t < hist_len; ++t) {
        float g = 0.0f;
        for (uint32_t j = 0; j < nj; ++j) g = __fadd_rn(g, __fmul_rn(my[t * nj + j], 0.1f));
        const float tg = (float)tanh((double)g);
        for (uint32_t l = 0; l < layers; ++l) {
            c = __fadd_rn(__fmul_rn(0.5f, c), __fmul_rn(0.5f, tg));
            h = __fmul_rn(0.5f, (float)tanh((double)c));
        }
    }
    h_out[b] = h;
}

constexpr int kRows = 128;   // vocabulary rows per CTA
constexpr int kBB = 8;       // sequences per inner pass

constexpr int kLogitSplit = 2;   // threads per vocabulary row: each takes kBB of the kLogitSplit * kBB sequences of a pass

__global__ void __launch_bounds__(kRows * kLogitSplit)
logits_kernel(const float* __restrict__ wout, uint32_t vocab, uint32_t hidden, const float* __restrict__ h,
              uint32_t batch, float* __restrict__ logits) {
    extern __shared__ float wt[];   // [kRows][hidden + 1]
    const uint32_t v0 = blockIdx.x * kRows;
    const uint32_t ld = hidden + 1;
    const uint32_t rows = min((uint32_t)kRows, vocab - v0);
    for (uint32_t i = threadIdx.x; i < rows * hidden; i += kRows * kLogitSplit) {
        const uint32_t r = i / hidden, j = i - r * hidden;
        wt[r * ld + j] = __ldg(wout + (size_t)(v0 + r) * hidden + j);
    }
    __syncthreads();
    const uint32_t r = threadIdx.x % kRows, part = threadIdx.x / kRows;
    if (r >= rows) return;
    const float* w = wt + r * ld;
    for (uint32_t b0 = (blockIdx.y * kLogitSplit + part) * kBB; b0 < batch; b0 += gridDim.y * kLogitSplit * kBB) {
        float hb[kBB], acc[kBB];
#pragma unroll
        for (int i = 0; i < kBB; ++i) {
            hb[i] = b0 + i < batch ? h[b0 + i] : 0.0f;
            acc[i] = 0.0f;
        }
        for (uint32_t j = 0; j < hidden; ++j) {
            const float wj = w[j];
#pragma unroll
            for (int i = 0; i < kBB; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(hb[i], wj));   // logits[i] += h[j]*w, :170
        }
#pragma unroll
        for (int i = 0; i < kBB; ++i)
            if (b0 + i < batch) logits[(size_t)(b0 + i) * vocab + v0 + r] = acc[i];
    }
}

constexpr int kTopThreads = 256;
constexpr int kMaxK = 16;

// a > b in the order of the reference's sort by confidence (softmax is monotone in the logit); ties -> lower id
__device__ __forceinline__ bool better(float av, uint32_t ai, float bv, uint32_t bi) {
    return av > bv || (av == bv && ai < bi);
}

// CTA per sequence, two passes over its logits instead of k + 2:
//   pass 1  every thread keeps the best K of its strided share in registers (sorted insertion) and the max;
//           K rounds of block arg-max over the heads of the per-thread lists then give the global top-k
//           (the winner pops its head), identical to k rounds of arg-max over all logits;
//   pass 2  sum of exp(logit - max) (expf per term, accumulated in fp64, rounded once).
// The k confidences are exp(best - max) / sum with exp evaluated in fp64 and rounded once (glibc's expf is within
// 1 ulp of that).
template <int K>
__global__ void __launch_bounds__(kTopThreads)
topk_kernel(const float* __restrict__ logits, uint32_t vocab, uint32_t k, uint32_t req_id, uint32_t layer_id,
            uint32_t* __restrict__ ids, float* __restrict__ conf, uint64_t* __restrict__ va) {
    __shared__ float sf[kTopThreads / 32];
    __shared__ double sd[kTopThreads / 32];
    __shared__ float s_best[kTopThreads / 32];
    __shared__ uint32_t s_idx[kTopThreads / 32];
    __shared__ float s_bcast;
    __shared__ double s_sum;
    __shared__ float s_win_v[kMaxK];
    __shared__ uint32_t s_win_i[kMaxK];
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* lg = logits + (size_t)b * vocab;
    // ---- pass 1: max logit (max_element, :176) and the thread's best K ----
    float bv[K];
    uint32_t bi[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        bv[i] = -FLT_MAX;
        bi[i] = 0xffffffffu;
    }
    float m = -FLT_MAX;
    for (uint32_t v = tid; v < vocab; v += kTopThreads) {
        const float x = lg[v];
        m = fmaxf(m, x);
        if (better(x, v, bv[K - 1], bi[K - 1])) {
            bv[K - 1] = x;
            bi[K - 1] = v;
#pragma unroll
            for (int i = K - 1; i > 0; --i) {
                if (better(bv[i], bi[i], bv[i - 1], bi[i - 1])) {
                    const float tv = bv[i]; bv[i] = bv[i - 1]; bv[i - 1] = tv;
                    const uint32_t ti = bi[i]; bi[i] = bi[i - 1]; bi[i - 1] = ti;
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sf[wid] = m;
    __syncthreads();
    if (tid == 0) {
        float mm = sf[0];
        for (int w = 1; w < kTopThreads / 32; ++w) mm = fmaxf(mm, sf[w]);
        s_bcast = mm;
    }
    __syncthreads();
    const float mx = s_bcast;
    // ---- global top-k: k rounds of block arg-max over the list heads ----
    for (uint32_t r = 0; r < k; ++r) {
        float best = bv[0];
        uint32_t idx = bi[0];
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (better(ob, oi, best, idx)) {
                best = ob;
                idx = oi;
            }
        }
        if (lane == 0) {
            s_best[wid] = best;
            s_idx[wid] = idx;
        }
        __syncthreads();
        if (tid == 0) {
            float bb = s_best[0];
            uint32_t ii = s_idx[0];
            for (int w = 1; w < kTopThreads / 32; ++w)
                if (better(s_best[w], s_idx[w], bb, ii)) {
                    bb = s_best[w];
                    ii = s_idx[w];
                }
            s_win_v[r] = bb;
            s_win_i[r] = ii;
        }
        __syncthreads();
        if (bi[0] == s_win_i[r] && bi[0] != 0xffffffffu) {   // the winner pops its head
#pragma unroll
            for (int i = 0; i + 1 < K; ++i) {
                bv[i] = bv[i + 1];
                bi[i] = bi[i + 1];
            }
            bv[K - 1] = -FLT_MAX;
            bi[K - 1] = 0xffffffffu;
        }
    }
    // ---- pass 2: sum of exp (:178-181); accumulated in fp64, rounded once ----
    double sum = 0.0;
    for (uint32_t v = tid; v < vocab; v += kTopThreads) sum += (double)expf(__fsub_rn(lg[v], mx));
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) sd[wid] = sum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kTopThreads / 32; ++w) t += sd[w];
        s_sum = t;
    }
    __syncthreads();
    const float denom = (float)s_sum;
    if (tid < k) {
        const float e = (float)exp((double)__fsub_rn(s_win_v[tid], mx));
        ids[(size_t)b * k + tid] = s_win_i[tid];
        conf[(size_t)b * k + tid] = __fdiv_rn(e, denom);                                  // logits[i] /= sum_exp, :183-185
        va[(size_t)b * k + tid] = ((uint64_t)req_id << 32) | ((uint64_t)layer_id << 16) | (uint64_t)(tid + 1);
    }
}

void free_predictor() {
    if (g_pred.d_emb) cudaFree(g_pred.d_emb);
    if (g_pred.d_wout) cudaFree(g_pred.d_wout);
    if (g_pred.d_hidden) cudaFree(g_pred.d_hidden);
    if (g_pred.d_logits) cudaFree(g_pred.d_logits);
    g_pred = Predictor();
}

}  // namespace
}  // namespace speckv

using namespace speckv;

extern "C" {

speckv_status_t speckv_ext_predictor_load(const float* h_embedding, const float* h_output, uint32_t vocab,
                                          uint32_t emb_dim, uint32_t hidden, uint32_t layers, uint32_t history_len) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    if (!h_embedding || !h_output || !vocab || !emb_dim || !hidden || !history_len) return SPECKV_ERR_INVAL;
    if ((size_t)(hidden + 1) * kRows * sizeof(float) > 200 * 1024) return SPECKV_ERR_INVAL;
    std::lock_guard<std::mutex> lk(g_pred_mu);
    free_predictor();
    cudaError_t e = cudaGetDevice(&g_pred.device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g_pred.d_emb, (size_t)vocab * emb_dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&g_pred.d_wout, (size_t)vocab * hidden * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(g_pred.d_emb, h_embedding, (size_t)vocab * emb_dim * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g_pred.d_wout, h_output, (size_t)vocab * hidden * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        free_predictor();
        return status_of(e);
    }
    g_pred.vocab = vocab;
    g_pred.emb_dim = emb_dim;
    g_pred.hidden = hidden;
    g_pred.layers = layers;
    g_pred.hist_len = history_len;
    return SPECKV_OK;
}

void speckv_ext_predictor_unload(void) {
    std::lock_guard<std::mutex> lk(g_pred_mu);
    if (device_count() > 0) free_predictor();
}

speckv_status_t speckv_ext_prefetch_score(const uint32_t* d_tokens, uint32_t batch, uint32_t k, uint32_t req_id,
                                          uint32_t layer_id, uint32_t* d_ids, float* d_conf, uint64_t* d_va,
                                          void* cuda_stream) {
    if (device_count() <= 0) return SPECKV_ERR_DRIVER;
    std::lock_guard<std::mutex> lk(g_pred_mu);
    if (!g_pred.d_emb) return SPECKV_ERR_INVAL;   // no predictor loaded
    if (batch == 0) return SPECKV_OK;
    if (!d_tokens || !d_ids || !d_conf || !d_va || k == 0 || k > (uint32_t)kMaxK || k > g_pred.vocab) return SPECKV_ERR_INVAL;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t e = cudaSuccess;
    if (batch > g_pred.max_batch) {
        if (g_pred.d_hidden) cudaFree(g_pred.d_hidden);
        if (g_pred.d_logits) cudaFree(g_pred.d_logits);
        g_pred.d_hidden = g_pred.d_logits = nullptr;
        g_pred.max_batch = 0;
        e = cudaMalloc((void**)&g_pred.d_hidden, (size_t)batch * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&g_pred.d_logits, (size_t)batch * g_pred.vocab * sizeof(float));
        if (e != cudaSuccess) return status_of(e);
        g_pred.max_batch = batch;
    }
    const uint32_t nj = g_pred.emb_dim < g_pred.hidden ? g_pred.emb_dim : g_pred.hidden;
    const size_t hsmem = (size_t)kHiddenWarps * g_pred.hist_len * nj * sizeof(float);
    if (hsmem <= 48 * 1024)
        hidden_warp_kernel<<<(batch + kHiddenWarps - 1) / kHiddenWarps, kHiddenWarps * 32, hsmem, st>>>(
            d_tokens, batch, g_pred.hist_len, g_pred.d_emb, g_pred.vocab, g_pred.emb_dim, g_pred.hidden, g_pred.layers,
            g_pred.d_hidden);
    else   // very long windows / wide embeddings: the one-thread-per-sequence form needs no staging
        hidden_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_tokens, batch, g_pred.hist_len, g_pred.d_emb, g_pred.vocab,
                                                           g_pred.emb_dim, g_pred.hidden, g_pred.layers, g_pred.d_hidden);
    const size_t smem = (size_t)(g_pred.hidden + 1) * kRows * sizeof(float);
    e = cudaFuncSetAttribute(logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_of(e);
    const unsigned vtiles = (g_pred.vocab + kRows - 1) / kRows;
    unsigned ysplit = (batch + kBB * kLogitSplit - 1) / (kBB * kLogitSplit);
    if (ysplit > 4) ysplit = 4;   // a few batch slices per vocabulary tile keep all SMs busy
    logits_kernel<<<dim3(vtiles, ysplit), kRows * kLogitSplit, smem, st>>>(g_pred.d_wout, g_pred.vocab, g_pred.hidden, g_pred.d_hidden,
                                                             batch, g_pred.d_logits);
    if (k <= 4)
        topk_kernel<4><<<batch, kTopThreads, 0, st>>>(g_pred.d_logits, g_pred.vocab, k, req_id, layer_id, d_ids, d_conf, d_va);
    else if (k <= 8)
        topk_kernel<8><<<batch, kTopThreads, 0, st>>>(g_pred.d_logits, g_pred.vocab, k, req_id, layer_id, d_ids, d_conf, d_va);
    else
        topk_kernel<kMaxK><<<batch, kTopThreads, 0, st>>>(g_pred.d_logits, g_pred.vocab, k, req_id, layer_id, d_ids, d_conf, d_va);
    count_launch(3);
    return status_of(cudaGetLastError());
}

}  // extern "C"

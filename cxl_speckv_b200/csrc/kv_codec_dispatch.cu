// kv_codec_dispatch.cu -- picks the kernel family for a codec call.
//
// fp16/bf16 groups of R * 2048 elements (R a power of two up to 128) go to the tuned
// TMA/cluster kernels; every group those flag (runs of 8+ equal deltas, non-finite or
// tiny scales) and every other geometry goes to the generic kernel.  Both launches are
// ordered on the caller's stream; the flag array is a per-stream buffer the library keeps across calls.
#include <cstdlib>

#include "device_ctx.h"
#include "kv_codec.h"

namespace speckv {

static bool force_generic() {
    static const bool v = [] {
        const char* e = std::getenv("SPECKV_FORCE_GENERIC");
        return e && e[0] == '1';
    }();
    return v;
}

template <typename Fast, typename Generic>
static cudaError_t two_pass(const CodecArgs& a, cudaStream_t st, bool decompress, Fast fast, Generic generic) {
    const int R = force_generic() ? 0 : fast_regions(a, decompress);
    if (R == 0) return generic(a, st, nullptr);
    // the per-group flag words live in a buffer that belongs to the stream and stays allocated: no allocation on
    // the call path, and the two launches can be captured into a CUDA graph
    uint32_t* flags = nullptr;
    cudaError_t e = scratch_persistent(reinterpret_cast<void**>(&flags), (size_t)a.n_groups * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    e = fast(R, a, flags, st);
    if (e == cudaSuccess) e = generic(a, st, flags);
    return e;
}

cudaError_t launch_compress(const CodecArgs& a, cudaStream_t st) {
    return two_pass(a, st, false, launch_compress_fast,
                    [](const CodecArgs& x, cudaStream_t s, const uint32_t* f) { return launch_compress_generic(x, s, f); });
}

cudaError_t launch_decompress(const CodecArgs& a, cudaStream_t st) {
    return two_pass(a, st, true, launch_decompress_fast,
                    [](const CodecArgs& x, cudaStream_t s, const uint32_t* f) { return launch_decompress_generic(x, s, f); });
}

}  // namespace speckv

// kv_codec_dispatch.cu -- picks the kernel family for a codec call.
#include "kv_codec.h"

namespace speckv {

cudaError_t launch_compress(const CodecArgs& a, cudaStream_t st) { return launch_compress_generic(a, st); }

cudaError_t launch_decompress(const CodecArgs& a, cudaStream_t st) { return launch_decompress_generic(a, st); }

}  // namespace speckv

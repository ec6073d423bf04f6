// kv_codec.h -- internal launch interface between the C ABI (ext_api.cu) and the codec kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace speckv {

struct CodecArgs {
    const void* in = nullptr;        // compress: elements in; decompress: unused
    void* out = nullptr;             // decompress: elements out
    void* payload = nullptr;         // slots (written by compress, read by decompress)
    float* scales = nullptr;
    uint32_t* comp_bytes = nullptr;
    uint32_t* out_elems = nullptr;   // decompress, optional
    size_t slot_bytes = 0;
    uint32_t group_elems = 0;
    uint32_t n_groups = 0;
    int dtype = 0;                   // DT_F16 / DT_BF16 / DT_F32
    int scheme = 2;                  // speckv_comp_scheme_t
    int sm_count = 148;
};

// any geometry, any alignment (kv_codec_generic.cu)
cudaError_t launch_compress_generic(const CodecArgs& a, cudaStream_t st);
cudaError_t launch_decompress_generic(const CodecArgs& a, cudaStream_t st);

// dispatch: tuned kernels for the common geometries, generic otherwise (kv_codec_dispatch.cu)
cudaError_t launch_compress(const CodecArgs& a, cudaStream_t st);
cudaError_t launch_decompress(const CodecArgs& a, cudaStream_t st);

}  // namespace speckv

// codec_math.cuh -- bit-exact arithmetic of the KV block codec on sm_100a.
//
// Mirrors, operation for operation, what the reference computes on x86-64
// (src/fpga_engine/cache_engine.cpp):
//   scale  s = max|x| / 127.0f, or 1.0f if the max is 0             (:172-184)
//   code   q = int8( (int32) roundf( (x / s) * 127.0f ) )           (:186-196)
//            = low byte of the truncating float->int32 conversion; NaN and
//              out-of-range give 0x80000000 -> 0 (cvttss2si), so codes WRAP
//   value  y = ((float) q / 127.0f) * s                             (:275-284)
// Every float op is an explicit round-to-nearest intrinsic so nvcc can neither
// contract nor reassociate it; denormals are kept (no -ftz, no fast-math).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace speckv {

enum : int { DT_F16 = 0, DT_BF16 = 1, DT_F32 = 2 };

// ---- widening / narrowing at the fp16 / bf16 boundary ------------------------
template <typename T> __device__ __forceinline__ float widen(T v);
template <> __device__ __forceinline__ float widen<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float widen<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float widen<float>(float v) { return v; }

template <typename T> __device__ __forceinline__ T narrow(float v);
template <> __device__ __forceinline__ __half narrow<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 narrow<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float narrow<float>(float v) { return v; }

// ---- scale --------------------------------------------------------------------
// `abs_val > max_val` (cache_engine.cpp:176-179) never admits a NaN; fmaxf
// returns the non-NaN operand, so with m starting at 0 it is the same function.
__device__ __forceinline__ float absmax_step(float m, float x) { return fmaxf(m, fabsf(x)); }
__device__ __forceinline__ float scale_from_max(float m) { return (m > 0.0f) ? __fdiv_rn(m, 127.0f) : 1.0f; }
// The clamped schemes (ids 3 / 4: this library's extension, not reference behaviour) store s = max|x|: then
// (x / s) * 127 spans [-127, 127], the clamp of cache_engine.cpp:192 never acts and the wrapping cast is the
// identity on every code -- the quantiser below is the same code path with a different scale.
__device__ __forceinline__ float scale_for(float m, bool max_scale) {
    return max_scale ? ((m > 0.0f) ? m : 1.0f) : scale_from_max(m);
}

// ---- rounding + the wrapping cast ------------------------------------------------
// std::round (half away from zero) then static_cast<int8_t>:
//   roundf(t) == sign(t) * floor(|t| + 0.5); the add is done round-toward-zero
//   so that 0.49999997f + 0.5f does not round up to 1.0f.
// kRangeCheck is needed only when |t| may reach 2^31 or be +-inf (x86 yields
// the "integer indefinite" 0x80000000 there, whose low byte is 0; cvt.rzi
// would saturate instead).  NaN converts to 0 on both machines.
template <bool kRangeCheck>
__device__ __forceinline__ uint32_t round_wrap_u8(float t) {
    float a = __fadd_rz(fabsf(t), 0.5f);
    int k = __float2int_rz(a);
    if (kRangeCheck) {
        if (!(a < 2147483648.0f)) k = 0;
    }
    k = (t < 0.0f) ? -k : k;
    return (uint32_t)k & 0xffu;
}

// ---- quantiser --------------------------------------------------------------------
// Exact form: two IEEE operations as written in the reference.
__device__ __forceinline__ uint32_t quantize_exact(float x, float s) {
    float y = __fdiv_rn(x, s);
    return round_wrap_u8<true>(__fmul_rn(y, 127.0f));
}

// Fast form: the reciprocal of the scale is computed once per group as an unevaluated sum
//     rh = RN(1/s),  rl = RN(RN(1 - s*rh) * rh)          (recip_hi_lo: one Newton residual)
// and each element takes two operations:
//     y = RN(x*rh + RN(x*rl))
// which carries ~47 bits of x/s into the single final rounding.  oracle/verify_fastdiv.c proves by
// exhaustion that the resulting CODE equals the reference's for every (x, max) pair of fp16 values,
// and for every pair of bf16 values with 2^-60 <= max < 2^120; groups outside that domain (and fp32
// input) take quantize_exact.  |t| <= 16130 here, so no range check is needed.
// (The previous form, y0 = RN(x*rh); y = RN(y0 + RN(x - y0*s)*rh), is also exact but costs three.)
__device__ __forceinline__ void recip_hi_lo(float s, float& rh, float& rl) {
    rh = __frcp_rn(s);
    rl = __fmul_rn(__fmaf_rn(-s, rh, 1.0f), rh);
}
__device__ __forceinline__ uint32_t quantize_fast(float x, float rh, float rl) {
    const float y = __fmaf_rn(x, rh, __fmul_rn(x, rl));
    return round_wrap_u8<false>(__fmul_rn(y, 127.0f));
}

// Same value as an int (only its low byte is meaningful): t + copysign(0.5, t) added
// round-toward-zero, then truncated.  Used where the byte mask is applied later (packing).
__device__ __forceinline__ uint32_t quantize_fast_i(float x, float rh, float rl) {
    const float y = __fmaf_rn(x, rh, __fmul_rn(x, rl));
    const float t = __fmul_rn(y, 127.0f);
    const float half = __uint_as_float((__float_as_uint(t) & 0x80000000u) | 0x3f000000u);
    return (uint32_t)__float2int_rz(__fadd_rz(t, half));
}

// Two packed elements at once (the tuned kernels): the +-0.5 of the rounding step takes its sign from
// the INPUT (t has the sign of x whenever it is not zero, and +-0.5 both truncate to 0 when it is), so
// one LOP3 builds both halves as a packed fp16/bf16 pair and the mixed-precision add of sm_100
// (add.rz.f32.f16 / .bf16, SASS FHADD.RZ) consumes a half of it directly: 6.5 operations per element.
template <typename T> struct HalfConst;
template <> struct HalfConst<__half> { static constexpr uint32_t k = 0x38003800u; };          // (0.5h, 0.5h)
template <> struct HalfConst<__nv_bfloat16> { static constexpr uint32_t k = 0x3f003f00u; };   // (0.5bf, 0.5bf)
template <typename T> __device__ __forceinline__ float add_rz_mixed(uint16_t h, float t);
template <> __device__ __forceinline__ float add_rz_mixed<__half>(uint16_t h, float t) {
    float a;
    asm("add.rz.f32.f16 %0, %1, %2;" : "=f"(a) : "h"(h), "f"(t));
    return a;
}
template <> __device__ __forceinline__ float add_rz_mixed<__nv_bfloat16>(uint16_t h, float t) {
    float a;
    asm("add.rz.f32.bf16 %0, %1, %2;" : "=f"(a) : "h"(h), "f"(t));
    return a;
}
// Packed fp32 pairs (sm_100: mul / fma .f32x2, SASS FMUL2 / FFMA2): two IEEE operations per instruction, each half
// rounded exactly like the scalar form -- the same arithmetic in half the issue slots (measured: FMUL2 sustains the
// fp32 lane rate of FMUL).  ptxas treats mul.rn.f32x2 followed by add.rn.f32x2 as contractable even under
// --fmad=false (profiles/micro/f32x2_probe.cu), so the packed forms are used only where no such pair is adjacent:
// here the chain is mul -> fma -> mul -> (mixed-precision add, round toward zero).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t mul_f32x2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma_f32x2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// w = two packed elements, x0 / x1 = the same elements widened; half_k = HalfConst<T>::k held in a register
template <typename T>
__device__ __forceinline__ void quantize_fast_pair_i(uint32_t w, float x0, float x1, float rh, float rl, uint32_t half_k,
                                                     uint32_t& q0, uint32_t& q1) {
    uint32_t hs;
    asm("lop3.b32 %0, %1, 0x80008000, %2, 0xEA;" : "=r"(hs) : "r"(w), "r"(half_k));   // (w & signs) | halves
#ifdef SPECKV_SCALAR_QUANT
    const float t0 = __fmul_rn(__fmaf_rn(x0, rh, __fmul_rn(x0, rl)), 127.0f);
    const float t1 = __fmul_rn(__fmaf_rn(x1, rh, __fmul_rn(x1, rl)), 127.0f);
#else
    const f32x2_t x = pack_f32x2(x0, x1);
    const f32x2_t y = fma_f32x2(x, pack_f32x2(rh, rh), mul_f32x2(x, pack_f32x2(rl, rl)));
    float t0, t1;
    unpack_f32x2(mul_f32x2(y, pack_f32x2(127.0f, 127.0f)), t0, t1);
#endif
    q0 = (uint32_t)__float2int_rz(add_rz_mixed<T>((uint16_t)(hs & 0xffffu), t0));
    q1 = (uint32_t)__float2int_rz(add_rz_mixed<T>((uint16_t)(hs >> 16), t1));
}

// Is the fast form valid for a group whose max-abs is m (element type T)?
template <typename T> __device__ __forceinline__ bool fast_quant_ok(float m);
template <> __device__ __forceinline__ bool fast_quant_ok<__half>(float m) { return m > 0.0f && m < __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ bool fast_quant_ok<__nv_bfloat16>(float m) {
    return m >= 8.673617379884035e-19f /* 2^-60 */ && m < 1.329227995784916e36f /* 2^120: above, rl loses bits */;
}
template <> __device__ __forceinline__ bool fast_quant_ok<float>(float) { return false; }

// ---- dequantiser -------------------------------------------------------------------
// (float) q / 127.0f * s with q the signed code (the LOW BYTE of code; upper bits are ignored).
// q/127 is an IEEE division by a constant: with r127 = RN(1/127) and d127 = RN(1/127 - r127),
//     RN(q/127) == RN(q*r127 + RN(q*d127))        (one FMUL + one FMA)
// holds for all 256 codes (oracle/verify_fastdiv.c checks them: q/127 has a 7-bit periodic
// expansion, so it is never within the tiny error of a rounding boundary).
__device__ __forceinline__ float dequantize(uint32_t code, float s) {
    const float qf = (float)(int)(int8_t)code;
    const float r127 = __uint_as_float(0x3c010204u);   // RN(1/127)
    const float d127 = __uint_as_float(0x2e010204u);   // RN(1/127 - r127)
    const float y = __fmaf_rn(qf, r127, __fmul_rn(qf, d127));
    return __fmul_rn(y, s);
}

// Non-finite scale (a group that contained +-inf, or a caller-supplied NaN): x86
// produces the "real indefinite" NaN 0xFFC00000 for 0*inf and propagates a NaN
// operand quieted, whereas the GPU would return its canonical 0x7FFFFFFF.  Keep
// the fp32 output bit-identical to the reference in those groups too.
// Narrowing of those results at the fp16 / bf16 boundary: a NaN keeps its sign and the top bits of its payload,
// quieted (what x86's vcvtps2ph and the oracle's bf16 truncation produce: 0xFFC00000 -> 0xFE00 / 0xFFC0), where
// cvt.rn.f16.f32 / cvt.rn.bf16.f32 would return the canonical 0x7FFF.  Everything else rounds as narrow<T> does.
template <typename T> __device__ __forceinline__ T narrow_special(float v) { return narrow<T>(v); }
template <> __device__ __forceinline__ __half narrow_special<__half>(float v) {
    if (v != v) {
        const uint32_t b = __float_as_uint(v);
        return __ushort_as_half((unsigned short)(((b >> 16) & 0x8000u) | 0x7c00u | 0x0200u | ((b >> 13) & 0x03ffu)));
    }
    return __float2half_rn(v);
}
template <> __device__ __forceinline__ __nv_bfloat16 narrow_special<__nv_bfloat16>(float v) {
    if (v != v) return __ushort_as_bfloat16((unsigned short)((__float_as_uint(v) >> 16) | 0x0040u));
    return __float2bfloat16_rn(v);
}

__device__ __forceinline__ bool scale_is_special(float s) { return !(fabsf(s) < __int_as_float(0x7f800000)); }
__device__ __forceinline__ float dequantize_special(uint32_t code_u8, float s) {
    const int q = (int)(int8_t)code_u8;
    if (s != s) return __int_as_float(__float_as_int(s) | 0x00400000);
    if (q == 0) return __int_as_float(0xffc00000);
    return ((q < 0) != (s < 0.0f)) ? __int_as_float(0xff800000) : __int_as_float(0x7f800000);
}

}  // namespace speckv

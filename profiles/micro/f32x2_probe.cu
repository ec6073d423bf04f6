// f32x2_probe.cu -- does ptxas keep mul.rn.f32x2 + add.rn.f32x2 unfused (it emits FFMA2 even under --fmad=false),
// and what do FMUL2 / FADD2 / FFMA2 cost against the scalar forms?   nvcc -arch=sm_100a --fmad=false
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// mode 0: scalar mul_rn + add_rn (reference).  1: mul2 + add2 as written.  2: fma2(h,w,-0) then fma2(p,1,acc).
// 3: true fused fma2(h, w, acc) (what a contraction would compute)
template <int MODE>
__global__ void chain(const float* __restrict__ w, const float* __restrict__ h, float* out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float h0 = h[2 * t], h1 = h[2 * t + 1];
    float a0 = 0.f, a1 = 0.f;
    u64 acc = pk(0.f, 0.f), h2 = pk(h0, h1);
    const u64 one = pk(1.f, 1.f), nz = pk(-0.f, -0.f);
    for (int j = 0; j < n; ++j) {
        const float wj = w[j];
        if (MODE == 0) { a0 = __fadd_rn(a0, __fmul_rn(h0, wj)); a1 = __fadd_rn(a1, __fmul_rn(h1, wj)); }
        else if (MODE == 1) acc = add2(acc, mul2(h2, pk(wj, wj)));
        else if (MODE == 2) acc = fma2(fma2(h2, pk(wj, wj), nz), one, acc);
        else acc = fma2(h2, pk(wj, wj), acc);
    }
    if (MODE != 0) upk(acc, a0, a1);
    out[2 * t] = a0; out[2 * t + 1] = a1;
}

// throughput: 8 independent chains per thread, ITER steps
template <int MODE>
__global__ void tput(float* out, int iters, float seed) {
    float a[16]; u64 p[8];
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const float c = 1.0000001f; const u64 c2 = pk(c, c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[2 * i] = __fmul_rn(a[2 * i], c); a[2 * i + 1] = __fmul_rn(a[2 * i + 1], c); }      // 16 FMUL
            else if (MODE == 1) p[i] = mul2(p[i], c2);                                                              // 8 FMUL2
            else if (MODE == 2) { a[2 * i] = __fmaf_rn(a[2 * i], c, c); a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], c, c); }
            else if (MODE == 3) p[i] = fma2(p[i], c2, c2);
            else if (MODE == 4) { a[2 * i] = __fadd_rn(a[2 * i], c); a[2 * i + 1] = __fadd_rn(a[2 * i + 1], c); }
            else p[i] = add2(p[i], c2);
        }
    }
    float s = 0.f;
    if (MODE & 1) for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); s += x + y; }
    else for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    const int T = 4096, N = 128;
    float *w, *h, *o[4];
    cudaMallocManaged(&w, N * 4); cudaMallocManaged(&h, 2 * T * 4);
    for (int i = 0; i < 4; ++i) cudaMallocManaged(&o[i], 2 * T * 4);
    srand(1);
    for (int i = 0; i < N; ++i) w[i] = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    for (int i = 0; i < 2 * T; ++i) h[i] = (rand() / (float)RAND_MAX - 0.5f);
    chain<0><<<T / 128, 128>>>(w, h, o[0], N); chain<1><<<T / 128, 128>>>(w, h, o[1], N);
    chain<2><<<T / 128, 128>>>(w, h, o[2], N); chain<3><<<T / 128, 128>>>(w, h, o[3], N);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    int d1 = 0, d2 = 0, d3 = 0;
    for (int i = 0; i < 2 * T; ++i) { d1 += o[1][i] != o[0][i]; d2 += o[2][i] != o[0][i]; d3 += o[3][i] != o[0][i]; }
    printf("vs scalar mul+add: mul2+add2 differs in %d / %d, fma2(-0)+fma2(1) differs in %d, fused fma2 differs in %d\n", d1, 2 * T, d2, d3);
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    const char* names[6] = {"FMUL", "FMUL2", "FFMA", "FFMA2", "FADD", "FADD2"};
    for (int m = 0; m < 6; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            switch (m) {
                case 0: tput<0><<<148 * 8, 256>>>(out, iters, 1.f); break;
                case 1: tput<1><<<148 * 8, 256>>>(out, iters, 1.f); break;
                case 2: tput<2><<<148 * 8, 256>>>(out, iters, 1.f); break;
                case 3: tput<3><<<148 * 8, 256>>>(out, iters, 1.f); break;
                case 4: tput<4><<<148 * 8, 256>>>(out, iters, 1.f); break;
                default: tput<5><<<148 * 8, 256>>>(out, iters, 1.f); break;
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("%-6s %.3f ms  -> %.2f T fp32-lane-ops/s\n", names[m], ms, 148.0 * 8 * 256 * 16 * iters / ms / 1e9);
        }
    }
    return 0;
}

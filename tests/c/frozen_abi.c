/*
 * frozen_abi.c -- plain C host code against libcxlspeckv.so (test infrastructure).
 *
 * Proves that include/speckv.h and include/speckv_ext.h are valid C, that the library links from C
 * without any C++ / Python / torch in between, and replays the call scenarios of the reference's own
 * tests/test_c_api.c (init/finalize, alloc/free, access, prefetch, params) together with the error
 * conventions measured on the reference.  With a device argument of the form
 * "cuda:N" it also runs a known-answer codec round trip on the GPU through the C ABI only: the KAT-1
 * vector recorded from the reference's FPGACacheEngine (tests/golden/golden.json).
 *
 *   gcc -Iinclude tests/c/frozen_abi.c -Lcxl_speckv_b200 -lcxlspeckv [-DWITH_CUDA -lcudart] -o frozen_abi
 *   ./frozen_abi /dev/null        (CPU box: setters answer SPECKV_ERR_DRIVER, like the reference on a fake device)
 *   ./frozen_abi cuda:0           (GPU box)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "speckv.h"
#include "speckv_ext.h"

#ifdef WITH_CUDA
#include <cuda_runtime_api.h>
#endif

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static void frozen_calls(const char* dev, int have_gpu) {
    speckv_handle_t h = 0, h2 = 0, h3 = 0;
    speckv_alloc_hint_t hint = {0, 0};
    void* p = (void*)0x1;
    int32_t tokens[16];
    int i;
    for (i = 0; i < 16; ++i) tokens[i] = i + 1;

    /* before init: everything answers INVAL */
    CHECK(speckv_alloc(4096, &hint, &h) == SPECKV_ERR_INVAL);
    CHECK(speckv_free(1) == SPECKV_ERR_INVAL);
    CHECK(speckv_access(1, 0, 1, &p) == SPECKV_ERR_INVAL);
    CHECK(speckv_prefetch(1, 0, 100, 4, tokens, 16) == SPECKV_ERR_INVAL);
    CHECK(speckv_set_prefetch_depth(4) == SPECKV_ERR_INVAL);
    CHECK(speckv_set_compression_scheme(SPECKV_COMP_INT8) == SPECKV_ERR_INVAL);

    CHECK(speckv_init("/nonexistent/speckv0") == SPECKV_ERR_GENERAL);   /* unopenable device */
    CHECK(speckv_init(dev) == SPECKV_OK);
    CHECK(speckv_init(dev) == SPECKV_ERR_GENERAL);                      /* double init */

    /* alloc / access: handles start at 1; address = 0x4000000000 + (h << 20) + (page << 12) + (offset & 0xFFF) */
    CHECK(speckv_alloc(1024 * 1024, &hint, &h) == SPECKV_OK && h == 1);
    CHECK(speckv_alloc(8192, NULL, &h2) == SPECKV_OK && h2 == 2);       /* the hint is ignored */
    CHECK(speckv_alloc(5000, &hint, &h3) == SPECKV_OK && h3 == 3);
    CHECK(speckv_alloc(4096, &hint, NULL) == SPECKV_ERR_INVAL);
    CHECK(speckv_access(h, 100, 4, &p) == SPECKV_OK && (uint64_t)(uintptr_t)p == 0x4000100064ull);
    CHECK(speckv_access(h2, 8191, 1, &p) == SPECKV_OK && (uint64_t)(uintptr_t)p == 0x4000201fffull);
    CHECK(speckv_access(h3, 4096, 1, &p) == SPECKV_OK && (uint64_t)(uintptr_t)p == 0x4000301000ull);
    p = (void*)0x5;
    CHECK(speckv_access(h3, 8192, 1, &p) == SPECKV_ERR_GENERAL && p == (void*)0x5);   /* past the end: out untouched */
    CHECK(speckv_access(77, 0, 1, &p) == SPECKV_ERR_GENERAL && p == (void*)0x5);      /* unknown handle */
    CHECK(speckv_access(h, 0, 1, NULL) == SPECKV_ERR_INVAL);

    CHECK(speckv_prefetch(1, 0, 100, 4, tokens, 16) == SPECKV_OK);
    CHECK(speckv_prefetch(1, 0, 100, 4, NULL, 16) == SPECKV_ERR_INVAL);
    CHECK(speckv_prefetch(1, 0, 100, 4, tokens, 0) == SPECKV_ERR_INVAL);

    /* setters: OK with a GPU behind the handle, DRIVER on a fake device (the reference's ioctl fails there) */
    CHECK(speckv_set_prefetch_depth(8) == (have_gpu ? SPECKV_OK : SPECKV_ERR_DRIVER));
    CHECK(speckv_set_compression_scheme(SPECKV_COMP_INT8_DELTA_RLE) == (have_gpu ? SPECKV_OK : SPECKV_ERR_DRIVER));

    CHECK(speckv_free(h) == SPECKV_OK);
    CHECK(speckv_free(h) == SPECKV_OK);        /* unknown (already freed) handle: still OK */
    CHECK(speckv_free(12345) == SPECKV_OK);
    speckv_finalize();
    CHECK(speckv_free(h2) == SPECKV_ERR_INVAL);
    /* handles restart at 1 after finalize -> init */
    CHECK(speckv_init(dev) == SPECKV_OK);
    CHECK(speckv_alloc(4096, &hint, &h) == SPECKV_OK && h == 1);
    speckv_finalize();
}

static void ext_without_compute(int have_gpu) {
    CHECK(speckv_ext_slot_bytes(131072, SPECKV_COMP_INT8_DELTA_RLE) == 262144);
    CHECK(speckv_ext_slot_bytes(10, SPECKV_COMP_INT8_DELTA_RLE) == 32);
    CHECK(speckv_ext_slot_bytes(10, SPECKV_COMP_INT8) == 16);
    CHECK(speckv_ext_version() != NULL && strlen(speckv_ext_version()) > 0);
    CHECK((speckv_ext_device_count() > 0) == (have_gpu != 0));
    if (!have_gpu) {   /* no CPU fallback: compute entry points fail loudly */
        CHECK(speckv_ext_compress(NULL, SPECKV_DTYPE_F16, 2048, 1, NULL, 4096, NULL, NULL, SPECKV_COMP_INT8_DELTA_RLE, NULL) ==
              SPECKV_ERR_DRIVER);
        CHECK(speckv_ext_translate(NULL, NULL, 4, NULL) == SPECKV_ERR_DRIVER);
    }
}

#ifdef WITH_CUDA
/* KAT-1 (recorded from the reference): scale 0x1.020408p-6, 9 pairs = 18 bytes */
static void device_known_answer(void) {
    const float in[10] = {0.0f, 1.0f, -1.0f, 0.5f, 0.25f, 0.25f, 0.25f, 2.0f, -2.0f, 1e-3f};
    const signed char want[18] = {0, 1, -127, 1, -2, 1, 65, 1, 32, 1, 0, 2, 33, 1, -2, 1, 9, 1};
    const float want_out[10] = {0.0f,          -0.0157480314f, 0.0157480314f,   -0.00793601573f, -0.00396800786f,
                                -0.00396800786f, -0.00396800786f, 0.000124000246f, -0.000124000246f, 0.000992001966f};
    const size_t slot = speckv_ext_slot_bytes(10, SPECKV_COMP_INT8_DELTA_RLE);
    float *d_in = NULL, *d_scale = NULL, *d_out = NULL, scale = 0.0f, out[10];
    unsigned char* d_pay = NULL;
    unsigned char pay[32];
    uint32_t *d_comp = NULL, *d_n = NULL, comp = 0, n_out = 0;
    uint64_t va[3] = {0x100000123ull, 0x100000FFFull, 0xFFFF000000000ABCull}, pa[3] = {0, 0, 0}, *d_va = NULL, *d_pa = NULL;
    int i;
    CHECK(cudaMalloc((void**)&d_in, sizeof in) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_out, sizeof out) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_pay, slot) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_scale, 4) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_comp, 4) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_n, 4) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_va, sizeof va) == cudaSuccess);
    CHECK(cudaMalloc((void**)&d_pa, sizeof pa) == cudaSuccess);
    CHECK(cudaMemcpy(d_in, in, sizeof in, cudaMemcpyHostToDevice) == cudaSuccess);
    CHECK(speckv_ext_compress(d_in, SPECKV_DTYPE_F32, 10, 1, d_pay, slot, d_scale, d_comp, SPECKV_COMP_INT8_DELTA_RLE, NULL) ==
          SPECKV_OK);
    CHECK(speckv_ext_decompress(d_pay, slot, d_scale, d_comp, 10, 1, SPECKV_DTYPE_F32, d_out, d_n,
                                SPECKV_COMP_INT8_DELTA_RLE, NULL) == SPECKV_OK);
    CHECK(cudaMemcpy(&scale, d_scale, 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(cudaMemcpy(&comp, d_comp, 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(cudaMemcpy(&n_out, d_n, 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(cudaMemcpy(pay, d_pay, slot, cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(cudaMemcpy(out, d_out, sizeof out, cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(scale == 0x1.020408p-6f);
    CHECK(comp == 18 && n_out == 10);
    CHECK(memcmp(pay, want, 18) == 0);
    for (i = 0; i < 10; ++i) CHECK(out[i] == want_out[i]);
    /* translate_address: the three addresses measured on the reference */
    CHECK(cudaMemcpy(d_va, va, sizeof va, cudaMemcpyHostToDevice) == cudaSuccess);
    CHECK(speckv_ext_translate(d_va, d_pa, 3, NULL) == SPECKV_OK);
    CHECK(cudaMemcpy(pa, d_pa, sizeof pa, cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(pa[0] == 0x4100000123ull && pa[1] == 0x4100000FFFull && pa[2] == 0x4000000ABCull);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_pay); cudaFree(d_scale); cudaFree(d_comp); cudaFree(d_n);
    cudaFree(d_va); cudaFree(d_pa);
}
#endif

int main(int argc, char** argv) {
    const char* dev = argc > 1 ? argv[1] : "/dev/null";
    const int have_gpu = strncmp(dev, "cuda", 4) == 0;
    frozen_calls(dev, have_gpu);
    ext_without_compute(have_gpu);
#ifdef WITH_CUDA
    if (have_gpu) device_known_answer();
#endif
    if (failures) {
        printf("%d check(s) failed\n", failures);
        return 1;
    }
    printf("frozen ABI ok (%s)\n", dev);
    return 0;
}

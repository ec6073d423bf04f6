/*
 * speckv.h -- the frozen C ABI of libcxlspeckv.so (drop-in boundary).
 *
 * These eight entry points, their types and their error behaviour are exactly
 * what the reference exports from host/include/speckv.h:12-66 and implements in
 * host/src/speckv_c_api.cpp:13-121; the reference's ctypes binding
 * (host/python/speckv_ctypes.py:15-57) and tests/test_c_api.c bind to them
 * unchanged.  Behind them the B200 build keeps the reference's host page table
 * (host/src/speckv_allocator.cpp) and replaces the ioctl/FPGA data path
 * (host/src/speckv_driver.cpp) with sm_100a CUDA kernels, see speckv_ext.h.
 *
 * Error conventions (measured on the reference):
 *   any call before speckv_init             -> SPECKV_ERR_INVAL
 *   second speckv_init                      -> SPECKV_ERR_GENERAL
 *   device path that cannot be opened       -> SPECKV_ERR_GENERAL
 *   NULL out pointer / tokens, history 0    -> SPECKV_ERR_INVAL
 *   access: unknown handle, offset past end -> SPECKV_ERR_GENERAL, *out untouched
 *   free of an unknown handle               -> SPECKV_OK
 *   setters when the device rejects them    -> SPECKV_ERR_DRIVER
 */
#ifndef SPECKV_H
#define SPECKV_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SPECKV_API __attribute__((visibility("default")))
#else
#define SPECKV_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {                 /* reference: speckv.h:12-18 */
    SPECKV_OK          = 0,
    SPECKV_ERR_GENERAL = -1,
    SPECKV_ERR_DRIVER  = -2,
    SPECKV_ERR_NOMEM   = -3,
    SPECKV_ERR_INVAL   = -4,
} speckv_status_t;

typedef struct {               /* reference: speckv.h:21-24 */
    uint32_t preferred_node;
    uint32_t reserved;
} speckv_alloc_hint_t;

typedef uint64_t speckv_handle_t;   /* reference: speckv.h:27 */

typedef enum {                 /* reference: speckv.h:59-63 */
    SPECKV_COMP_FP16           = 0,
    SPECKV_COMP_INT8           = 1,
    SPECKV_COMP_INT8_DELTA_RLE = 2,
} speckv_comp_scheme_t;

/* reference: speckv.h:30-31, speckv_c_api.cpp:13-39.
 * dev_path: "cuda" / "cuda:<ordinal>" selects a B200 directly; any other string
 * is treated like the reference treats it -- a path that must open O_RDWR
 * (e.g. "/dev/null" in the reference's own smoke setup) -- and the CUDA device
 * is then taken from $SPECKV_CUDA_DEVICE (default 0). */
SPECKV_API speckv_status_t speckv_init(const char* dev_path);
SPECKV_API void speckv_finalize(void);

/* reference: speckv.h:34-39, speckv_c_api.cpp:41-64 */
SPECKV_API speckv_status_t speckv_alloc(size_t bytes, const speckv_alloc_hint_t* hint, speckv_handle_t* out_handle);
SPECKV_API speckv_status_t speckv_free(speckv_handle_t handle);

/* reference: speckv.h:44-47, speckv_c_api.cpp:66-83, speckv_allocator.cpp:54-74 */
SPECKV_API speckv_status_t speckv_access(speckv_handle_t handle, uint64_t offset_bytes, size_t length_bytes,
                              void** out_gpu_ptr);

/* reference: speckv.h:51-56, speckv_c_api.cpp:85-99 */
SPECKV_API speckv_status_t speckv_prefetch(uint32_t req_id, uint16_t layer, uint32_t cur_pos, uint32_t depth_k,
                                const int32_t* recent_tokens, uint32_t history_len);

/* reference: speckv.h:65-66, speckv_c_api.cpp:101-121 */
SPECKV_API speckv_status_t speckv_set_prefetch_depth(uint32_t depth_k);
SPECKV_API speckv_status_t speckv_set_compression_scheme(speckv_comp_scheme_t scheme);

#ifdef __cplusplus
}
#endif
#endif /* SPECKV_H */

"""cxl_speckv_b200 -- B200-native KV block codec path of CXL-SpecKV.

The package is a thin host-side mirror of the reference's Python surface
(host/python/speckv_ctypes.py, host/python/vllm_speckv_backend.py) over
libcxlspeckv.so, whose hot path is hand-written sm_100a CUDA.  There is no CPU
fallback: loading fails loudly when the library has not been built, and every
compute call fails with SPECKV_ERR_DRIVER when no CUDA device is present.
"""
from ._lib import (SpeckvError, lib, lib_path, SPECKV_OK, SPECKV_ERR_GENERAL, SPECKV_ERR_DRIVER,
                   SPECKV_ERR_NOMEM, SPECKV_ERR_INVAL, COMP_FP16, COMP_INT8, COMP_INT8_DELTA_RLE, COMP_INT8_CLAMP_DELTA_RLE, COMP_INT8_CLAMP,
                   DTYPE_F16, DTYPE_BF16, DTYPE_F32)
from .speckv_ctypes import SpeckvLib
from .vllm_speckv_backend import CxlSpeckvKVAllocator

__all__ = ["SpeckvError", "lib", "lib_path", "SpeckvLib", "CxlSpeckvKVAllocator",
           "SPECKV_OK", "SPECKV_ERR_GENERAL", "SPECKV_ERR_DRIVER", "SPECKV_ERR_NOMEM", "SPECKV_ERR_INVAL",
           "COMP_FP16", "COMP_INT8", "COMP_INT8_DELTA_RLE", "COMP_INT8_CLAMP_DELTA_RLE", "COMP_INT8_CLAMP", "DTYPE_F16", "DTYPE_BF16", "DTYPE_F32"]

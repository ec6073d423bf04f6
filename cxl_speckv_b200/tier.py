"""Host tier: a pinned-DRAM pool standing in for the CXL memory pool (speckv_ext_tier_*).
Offload = compress on the GPU and move the payload to host memory; restore = the inverse."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .codec import _DTYPES, _stream


class HostTier:
    def __init__(self, pool_bytes: int):
        self._h = C.c_void_p()
        check(lib().speckv_ext_tier_create(pool_bytes, C.byref(self._h)), "speckv_ext_tier_create")

    def close(self):
        if self._h:
            lib().speckv_ext_tier_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def offload(self, x: torch.Tensor, group_elems: int, block_ids: np.ndarray) -> None:
        x = x.contiguous()
        ids = np.ascontiguousarray(block_ids, dtype=np.uint64)
        n = x.numel() // group_elems
        assert ids.size == n and x.numel() % group_elems == 0
        with torch.cuda.device(x.device):
            check(lib().speckv_ext_tier_offload(self._h, x.data_ptr(), _DTYPES[x.dtype], group_elems, n, ids.ctypes.data,
                                                _stream()), "speckv_ext_tier_offload")

    def restore(self, block_ids: np.ndarray, group_elems: int, dtype: torch.dtype, device="cuda:0",
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        ids = np.ascontiguousarray(block_ids, dtype=np.uint64)
        if out is None:
            out = torch.empty((ids.size, group_elems), dtype=dtype, device=device)
        with torch.cuda.device(out.device):
            check(lib().speckv_ext_tier_restore(self._h, ids.ctypes.data, ids.size, group_elems, _DTYPES[dtype],
                                                out.data_ptr(), _stream()), "speckv_ext_tier_restore")
        return out

    def drop(self, block_ids: np.ndarray) -> None:
        ids = np.ascontiguousarray(block_ids, dtype=np.uint64)
        check(lib().speckv_ext_tier_drop(self._h, ids.ctypes.data, ids.size), "speckv_ext_tier_drop")

    def stats(self) -> dict:
        s = _lib.TierStats()
        lib().speckv_ext_tier_get_stats(self._h, C.byref(s))
        return {k: getattr(s, k) for k, _ in s._fields_}

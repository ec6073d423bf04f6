"""Batched speculative-prefetch scoring (speckv_ext_predictor_load / speckv_ext_prefetch_score):
host-side mirror of SpeculativePrefetcher::prefetch (src/prefetcher/speculative_prefetcher.cpp:25-82)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np
import torch

from ._lib import check, lib

HISTORY_LEN = 16   # speculative_prefetcher.h:37


def window_history(histories: Sequence[Sequence[int]], history_len: int = HISTORY_LEN) -> np.ndarray:
    """Last history_len tokens, left-padded with token 0 (lstm_predictor.cpp:46-51)."""
    out = np.zeros((len(histories), history_len), dtype=np.uint32)
    for i, h in enumerate(histories):
        h = list(h)[-history_len:]
        if h:
            out[i, history_len - len(h):] = np.asarray(h, dtype=np.uint32)
    return out


def load_predictor(embedding: np.ndarray, output: np.ndarray, layers: int = 2, history_len: int = HISTORY_LEN) -> None:
    emb = np.ascontiguousarray(embedding, dtype=np.float32)
    out = np.ascontiguousarray(output, dtype=np.float32)
    vocab, emb_dim = emb.shape
    assert out.shape[0] == vocab
    check(lib().speckv_ext_predictor_load(emb.ctypes.data, out.ctypes.data, vocab, emb_dim, out.shape[1], layers,
                                          history_len), "speckv_ext_predictor_load")


def score(tokens: torch.Tensor, k: int = 4, layer_id: int = 0, req_id: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """tokens: int32 CUDA tensor [batch, history_len] (see window_history).
    -> (ids [batch,k] int32, confidence [batch,k] float32, request addresses [batch,k] int64)."""
    tokens = tokens.to(torch.int32).contiguous()
    b = tokens.shape[0]
    ids = torch.empty((b, k), dtype=torch.int32, device=tokens.device)
    conf = torch.empty((b, k), dtype=torch.float32, device=tokens.device)
    va = torch.empty((b, k), dtype=torch.int64, device=tokens.device)
    with torch.cuda.device(tokens.device):
        st = lib().speckv_ext_prefetch_score(tokens.data_ptr(), b, k, req_id, layer_id, ids.data_ptr(), conf.data_ptr(),
                                             va.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(st, "speckv_ext_prefetch_score")
    return ids, conf, va

"""Batched speculative-prefetch scoring (speckv_ext_predictor_load / speckv_ext_prefetch_score):
host-side mirror of SpeculativePrefetcher::prefetch (src/prefetcher/speculative_prefetcher.cpp:25-82)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

REQUEST_BYTES = 32   # sizeof(PrefetchRequest), speculative_prefetcher.h:23-29

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


def table_records(batch: int, k: int) -> int:
    """Records in a request table: one header + batch * k requests."""
    return 1 + batch * k


def emit(tokens: torch.Tensor, k: int = 4, layer_id: int = 0, req_id: int = 0, req_ids: torch.Tensor | None = None,
         page_table: torch.Tensor | None = None, va_base: int = 0, timestamp: int = 0,
         table: torch.Tensor | None = None, want_predictions: bool = False):
    """SpeculativePrefetcher::prefetch for a batch, on the device (speckv_ext_prefetch_emit): scoring, residency
    filter against `page_table` (uint8 view of speckv_page_t records, or None), PrefetchRequest emission.
    -> table: uint8 [1 + batch * k, 32]; record 0's first 8 bytes hold the number of requests that follow.
    With want_predictions also (ids, conf) of the unfiltered predictions."""
    tokens = tokens.to(torch.int32).contiguous()
    b = tokens.shape[0]
    if table is None:
        table = torch.zeros((table_records(b, k), REQUEST_BYTES), dtype=torch.uint8, device=tokens.device)
    ids = conf = None
    if want_predictions:
        ids = torch.empty((b, k), dtype=torch.int32, device=tokens.device)
        conf = torch.empty((b, k), dtype=torch.float32, device=tokens.device)
    n_pages = 0 if page_table is None else page_table.numel() // 24
    with torch.cuda.device(tokens.device):
        st = lib().speckv_ext_prefetch_emit(tokens.data_ptr(), b, k, req_ids.data_ptr() if req_ids is not None else None,
                                            req_id, layer_id, page_table.data_ptr() if page_table is not None else None,
                                            n_pages, va_base, timestamp, table.data_ptr(),
                                            ids.data_ptr() if ids is not None else None,
                                            conf.data_ptr() if conf is not None else None,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(st, "speckv_ext_prefetch_emit")
    return (table, ids, conf) if want_predictions else table


def unpack_table(table: torch.Tensor):
    """Host view of a request table: (n, va uint64[n], layer uint32[n], token uint32[n], conf float32[n], ts uint64[n])."""
    raw = table.detach().cpu().numpy().reshape(-1, REQUEST_BYTES)
    n = int(raw[0, :8].copy().view(np.uint64)[0])
    body = raw[1:1 + n]
    return (n, body[:, 0:8].copy().view(np.uint64).ravel(), body[:, 8:12].copy().view(np.uint32).ravel(),
            body[:, 12:16].copy().view(np.uint32).ravel(), body[:, 16:20].copy().view(np.float32).ravel(),
            body[:, 24:32].copy().view(np.uint64).ravel())


def route(tables: torch.Tensor, n_tables: int, table_capacity: int, n_blocks_total: int, world: int, rank: int,
          block_index: torch.Tensor | None = None, count: torch.Tensor | None = None,
          request_index: torch.Tensor | None = None):
    """speckv_ext_route_requests: the requests of `tables` ([n_tables, 1 + table_capacity, 32] uint8) whose block
    (token % n_blocks_total) this rank owns -> (block_index int32 [n_tables * table_capacity], count int32 [1]), on the device."""
    dev = tables.device
    if block_index is None:
        block_index = torch.empty(n_tables * table_capacity, dtype=torch.int32, device=dev)
    if count is None:
        count = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = lib().speckv_ext_route_requests(tables.data_ptr(), n_tables, table_capacity, n_blocks_total, world, rank,
                                             block_index.data_ptr(),
                                             request_index.data_ptr() if request_index is not None else None,
                                             count.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(st, "speckv_ext_route_requests")
    return block_index, count


def handle_misprediction(actual_token: int, predicted) -> bool:
    """SpeculativePrefetcher::handle_misprediction (speculative_prefetcher.cpp:84-97) -> was the prediction correct."""
    p = np.ascontiguousarray(predicted, dtype=np.uint32)
    ok = C.c_int(0)
    check(lib().speckv_ext_prefetch_handle_misprediction(int(actual_token), p.ctypes.data_as(C.POINTER(C.c_uint32)), p.size,
                                                         C.byref(ok)), "speckv_ext_prefetch_handle_misprediction")
    return bool(ok.value)


def statistics(reset: bool = False) -> dict:
    s = _lib.PrefetchStats()
    check(lib().speckv_ext_prefetch_stats(C.byref(s), int(reset)), "speckv_ext_prefetch_stats")
    return {k: getattr(s, k) for k, _ in s._fields_}


def outstanding(vas=()):
    """(found flags for `vas`, queue of the most recent requests as a list of dicts, oldest first)."""
    v = np.ascontiguousarray(vas, dtype=np.uint64)
    found = np.zeros(v.size, dtype=np.uint8)
    q = (_lib.PrefetchRequest * 16)()
    n = C.c_uint32(0)
    check(lib().speckv_ext_prefetch_outstanding(v.ctypes.data_as(C.POINTER(C.c_uint64)), v.size,
                                                found.ctypes.data_as(C.POINTER(C.c_uint8)), q, C.byref(n), None),
          "speckv_ext_prefetch_outstanding")
    return found.astype(bool), [{"virtual_addr": q[i].virtual_addr, "layer_id": q[i].layer_id,
                                 "predicted_token_id": q[i].predicted_token_id, "confidence": q[i].confidence,
                                 "timestamp": q[i].timestamp} for i in range(n.value)]

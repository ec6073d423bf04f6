"""Multi-GPU sharding of the KV block codec path (one process per GPU).

KV blocks are independent units (one scale, one delta/RLE stream per group, SURVEY.md
section 8e), so they shard by layer (and, inside a layer, by KV head) with no data-path
collective.  The only exchange is the per-step all-gather of the page-table metadata
every rank needs to locate its peers' compressed blocks: one KvPageHandle-like record
per group (reference: host/include/speckv_allocator.hpp:22-27) -- here the group's
compressed size, from which slot occupancy and host-tier offsets follow.
Backend-agnostic: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_layers(n_layers: int, world: int, rank: int) -> List[int]:
    """Contiguous layer ownership (80 / G contiguous layers per GPU): rank r holds layers
    [r * n // world, (r + 1) * n // world) -- the split pipeline-parallel serving already uses, so compressed
    blocks never cross NVLink."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return list(range(rank * n_layers // world, (rank + 1) * n_layers // world))


def shard_groups(n_layers: int, groups_per_layer: int, world: int, rank: int) -> Tuple[List[int], int]:
    """-> (owned layers, number of groups this rank holds)."""
    layers = shard_layers(n_layers, world, rank)
    return layers, len(layers) * groups_per_layer


def owner_of(layer: int, n_layers: int, world: int) -> int:
    """Rank that holds `layer` under the contiguous split."""
    if not 0 <= layer < n_layers:
        raise ValueError("layer out of range")
    return ((layer + 1) * world - 1) // n_layers


def gather_page_metadata(comp_bytes: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather of the per-group compressed sizes (int32, same length on every rank).
    Returns a [world, n_groups] tensor; row r is rank r's table."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    n = comp_bytes.numel()
    if out is None:
        out = torch.empty(world * n, dtype=comp_bytes.dtype, device=comp_bytes.device)
    if world == 1:
        out.copy_(comp_bytes.reshape(-1))
    else:
        dist.all_gather_into_tensor(out, comp_bytes.reshape(-1).contiguous())
    return out.view(world, n)


def global_block_location(layer: int, block_in_layer: int, groups_per_layer: int, n_layers: int, world: int) -> Tuple[int, int]:
    """(rank, local group index) of a block under the contiguous layer split."""
    rank = owner_of(layer, n_layers, world)
    local_layer = layer - rank * n_layers // world
    return rank, local_layer * groups_per_layer + block_in_layer


# ---- fixed-size metadata records ---------------------------------------
# KvPageHandle {u64 virt_page_id, u64 phys_page_id, u32 size_bytes, u32 flags} = 24 bytes
# (host/include/speckv_allocator.hpp:22-27; the layout speckv_ext_page_table_export writes), and
# PrefetchRequest {u64 virtual_addr, u32 layer_id, u32 token_id, f32 confidence, u64 timestamp} = 32 bytes
# with natural alignment (src/prefetcher/speculative_prefetcher.h:23-29).
PAGE_RECORD_BYTES = 24
PREFETCH_RECORD_BYTES = 32


def gather_records(records: torch.Tensor, record_bytes: int) -> torch.Tensor:
    """All-gather of a table of fixed-size records (uint8 [n, record_bytes], same n on every rank;
    pad with zero records).  Returns uint8 [world, n, record_bytes]; one collective per step."""
    if records.dtype != torch.uint8 or records.dim() != 2 or records.shape[1] != record_bytes:
        raise ValueError(f"records must be uint8 [n, {record_bytes}]")
    world = dist.get_world_size() if dist.is_initialized() else 1
    flat = records.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=torch.uint8, device=records.device)
    if world == 1:
        out.copy_(flat)
    else:
        dist.all_gather_into_tensor(out, flat)
    return out.view(world, records.shape[0], record_bytes)


def pack_prefetch_requests(va: torch.Tensor, layer: int, token_ids: torch.Tensor, conf: torch.Tensor,
                           timestamp: int = 0) -> torch.Tensor:
    """The scoring kernel's outputs (va int64 / ids int32 / conf float32, any shape) as PrefetchRequest
    records: uint8 [n, 32]."""
    n = va.numel()
    rec = torch.zeros((n, 4), dtype=torch.int64, device=va.device)
    rec[:, 0] = va.reshape(-1)
    lo = torch.full((n,), int(layer), dtype=torch.int64, device=va.device)
    hi = token_ids.reshape(-1).to(torch.int64) & 0xFFFFFFFF
    rec[:, 1] = lo | (hi << 32)                                             # layer_id (low word), token_id (high word)
    rec[:, 2] = conf.reshape(-1).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF   # confidence, 4 bytes of padding
    rec[:, 3] = int(timestamp)
    return rec.view(torch.uint8).view(n, PREFETCH_RECORD_BYTES)


def unpack_prefetch_requests(rec: torch.Tensor):
    """Inverse of pack_prefetch_requests on the last dimension: (va int64, layer int32, token int32, conf float32)."""
    w = rec.contiguous().view(torch.int64).reshape(*rec.shape[:-1], 4)
    va = w[..., 0]
    layer = (w[..., 1] & 0xFFFFFFFF).to(torch.int32)
    tok = ((w[..., 1] >> 32) & 0xFFFFFFFF).to(torch.int32)
    conf = w[..., 2].to(torch.int32).view(torch.float32)    # low word (the cast keeps the low 32 bits)
    return va, layer, tok, conf

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
    """Round-robin layer ownership: layer l lives on rank l % world (what TP/PP serving
    already does, so compressed blocks never cross NVLink)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return [l for l in range(n_layers) if l % world == rank]


def shard_groups(n_layers: int, groups_per_layer: int, world: int, rank: int) -> Tuple[List[int], int]:
    """-> (owned layers, number of groups this rank holds)."""
    layers = shard_layers(n_layers, world, rank)
    return layers, len(layers) * groups_per_layer


def owner_of(layer: int, world: int) -> int:
    return layer % world


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


def global_block_location(layer: int, block_in_layer: int, groups_per_layer: int, world: int) -> Tuple[int, int]:
    """(rank, local group index) of a block under round-robin layer sharding."""
    rank = owner_of(layer, world)
    local_layer = layer // world
    return rank, local_layer * groups_per_layer + block_in_layer

"""CPU, world_size 2 over gloo: the N > 1 host logic of the codec path -- layer sharding and
the per-step all-gather of page-table metadata (the only collective in the design)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cxl_speckv_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_layers, gpl, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layers, n_groups = sharding.shard_groups(n_layers, gpl, world, rank)
        # synthetic per-group compressed sizes that encode (rank, layer, block) so the gather can be checked
        comp = torch.empty(n_groups, dtype=torch.int32)
        for li, layer in enumerate(layers):
            for b in range(gpl):
                comp[li * gpl + b] = 1_000_000 * rank + 1000 * layer + b
        # equal table lengths are required by all_gather_into_tensor: pad to the max over ranks
        n_max = max(len(sharding.shard_layers(n_layers, world, r)) for r in range(world)) * gpl
        padded = torch.full((n_max,), -1, dtype=torch.int32)
        padded[:n_groups] = comp
        table = sharding.gather_page_metadata(padded)
        ok = table.shape == (world, n_max)
        for layer in range(n_layers):
            for b in (0, gpl - 1):
                r, idx = sharding.global_block_location(layer, b, gpl, n_layers, world)
                ok &= int(table[r, idx]) == 1_000_000 * r + 1000 * layer + b
        ok &= sorted(sum((sharding.shard_layers(n_layers, world, r) for r in range(world)), [])) == list(range(n_layers))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_layers,gpl", [(80, 128), (5, 3)])
def test_layer_sharding_and_metadata_allgather(n_layers, gpl):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_layers, gpl, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_single_process_paths():
    assert sharding.shard_layers(80, 8, 3) == list(range(30, 40))
    assert sharding.owner_of(17, 80, 4) == 0 and sharding.owner_of(20, 80, 4) == 1 and sharding.owner_of(79, 80, 8) == 7
    assert sharding.global_block_location(27, 5, 128, 80, 4) == (1, 7 * 128 + 5)
    for n, w in ((80, 8), (5, 2), (7, 3), (3, 4)):
        for l in range(n):
            assert l in sharding.shard_layers(n, w, sharding.owner_of(l, n, w))
    t = torch.arange(6, dtype=torch.int32)
    assert torch.equal(sharding.gather_page_metadata(t), t.view(1, 6))
    with pytest.raises(ValueError):
        sharding.shard_layers(4, 2, 2)


def _record_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # prefetch requests of this rank's sequences (PrefetchRequest records)
        n = 12
        va = (torch.arange(n, dtype=torch.int64) + 1) | (rank << 32) | (5 << 16)
        ids = torch.arange(n, dtype=torch.int32) * 7 + rank
        conf = torch.linspace(0.9, 0.1, n) / (rank + 1)
        rec = sharding.pack_prefetch_requests(va, 5, ids, conf, timestamp=1000 + rank)
        allrec = sharding.gather_records(rec, sharding.PREFETCH_RECORD_BYTES)
        ok = allrec.shape == (world, n, 32)
        for r in range(world):
            v, l, t, c = sharding.unpack_prefetch_requests(allrec[r])
            ok &= torch.equal(v, (torch.arange(n, dtype=torch.int64) + 1) | (r << 32) | (5 << 16))
            ok &= bool((l == 5).all()) and torch.equal(t, torch.arange(n, dtype=torch.int32) * 7 + r)
            ok &= torch.equal(c, torch.linspace(0.9, 0.1, n) / (r + 1))
            ok &= int(allrec[r].view(torch.int64).view(n, 4)[0, 3]) == 1000 + r
        # page-table deltas: 24-byte KvPageHandle records {virt, phys, size, flags}
        pages = torch.zeros((4, 3), dtype=torch.int64)
        pages[:, 0] = ((rank + 1) << 32) | (torch.arange(4) << 12)
        pages[:, 1] = 0x4000000000 + ((rank + 1) << 20) + (torch.arange(4) << 12)
        pages[:, 2] = 4096 | (0b101 << 32)
        allpages = sharding.gather_records(pages.view(torch.uint8).view(4, 24), sharding.PAGE_RECORD_BYTES)
        for r in range(world):
            p = allpages[r].contiguous().view(torch.int64).view(4, 3)
            ok &= int(p[2, 0]) == ((r + 1) << 32) | (2 << 12) and int(p[3, 1]) == 0x4000000000 + ((r + 1) << 20) + (3 << 12)
            ok &= int(p[0, 2]) & 0xFFFFFFFF == 4096 and int(p[0, 2]) >> 32 == 0b101
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_fixed_size_record_allgather():
    """The per-step metadata exchange of SURVEY.md section 8e: PrefetchRequest (32 B) and KvPageHandle (24 B)
    records, one all-gather each, world size 2 over gloo."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_record_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

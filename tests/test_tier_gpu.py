"""GPU: offload to / restore from the pinned host tier (BASELINE config 4 geometry: 4 KiB page
groups).  The restored pages must equal a direct compress -> decompress on the device bit for
bit, which in turn is bit-exact against the oracle (tests/test_codec_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle.oracle import Port

pytestmark = pytest.mark.gpu

from cxl_speckv_b200 import codec  # noqa: E402
from cxl_speckv_b200.tier import HostTier  # noqa: E402
from cxl_speckv_b200 import SpeckvError  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("G,n_groups", [(2048, 50000), (131072, 1200), (1000, 333)])
def test_offload_restore_roundtrip(G, n_groups):
    torch.manual_seed(G)
    x = torch.randn(n_groups * G, device=DEV).half()
    x[3 * G:5 * G] = 0                                    # highly compressible pages (tiny payload)
    x[7 * G:8 * G] = 1.5
    tier = HostTier(pool_bytes=int(n_groups * G * 2 * 1.6) + (1 << 20))   # chunks need contiguous extents
    try:
        ids = (np.arange(n_groups, dtype=np.uint64) << np.uint64(12)) | (np.uint64(1) << np.uint64(32))   # virt_page_id style
        tier.offload(x, G, ids)
        st = tier.stats()
        assert st["blocks"] == n_groups and st["bytes_offloaded_raw"] == n_groups * G * 2
        c = codec.compress(x, G)
        comp = c.comp_bytes.cpu().numpy().view(np.uint32).astype(np.uint64)
        assert st["used_bytes"] == int(((comp + 15) // 16 * 16).sum()) == st["bytes_offloaded_stored"]
        want = codec.decompress(c)
        y = tier.restore(ids, G, torch.float16)
        assert torch.equal(y.view(torch.int16), want.view(torch.int16))
        # a scattered subset, in another order (prefetch / page-fault pattern)
        rng = np.random.default_rng(0)
        sel = rng.permutation(n_groups)[:257]
        y2 = tier.restore(ids[sel], G, torch.float16)
        assert torch.equal(y2.view(torch.int16), want[torch.from_numpy(sel).to(DEV)].view(torch.int16))
        # oracle spot check on the restored bytes
        for g in (0, 3, 7, n_groups - 1):
            s, p = Port.compress(x[g * G:(g + 1) * G].cpu().numpy().astype(np.float32))
            ref = Port.decompress(s, p, G).astype(np.float16)
            assert np.array_equal(y[g].cpu().numpy().view(np.uint16), ref.view(np.uint16))
        # drop half, re-offload new content under the same ids: space is reused
        tier.drop(ids[: n_groups // 2])
        assert tier.stats()["blocks"] == n_groups - n_groups // 2
        with pytest.raises(SpeckvError):
            tier.restore(ids[:4], G, torch.float16)       # unknown after drop -> SPECKV_ERR_GENERAL
        x2 = torch.randn((n_groups // 2) * G, device=DEV).half()
        tier.offload(x2, G, ids[: n_groups // 2])
        y3 = tier.restore(ids[: n_groups // 2], G, torch.float16)
        assert torch.equal(y3.view(torch.int16), codec.decompress(codec.compress(x2, G)).view(torch.int16))
        assert tier.stats()["used_bytes"] <= tier.stats()["pool_bytes"]
    finally:
        tier.close()


def test_pool_exhaustion_is_reported():
    G, n_groups = 2048, 4096
    x = torch.randn(n_groups * G, device=DEV).half()
    tier = HostTier(pool_bytes=1 << 20)                   # far too small
    try:
        with pytest.raises(SpeckvError) as ei:
            tier.offload(x, G, np.arange(n_groups, dtype=np.uint64))
        assert ei.value.status == -3                       # SPECKV_ERR_NOMEM
    finally:
        tier.close()

"""Helper of tests/test_codec_gpu.py::test_tuned_and_generic_kernels_agree_at_scale (test infrastructure):
compresses / decompresses a seeded data set and prints digests that do not depend on the bytes a slot holds
past its payload.  Run in a fresh process so that SPECKV_FORCE_GENERIC (read once per process) can differ."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cxl_speckv_b200 import codec


def digests(G: int, n_groups: int, dtype: str):
    torch.manual_seed(4321)
    dt = torch.float16 if dtype == "f16" else torch.bfloat16
    x = torch.randn(n_groups, G, device="cuda:0")
    # a third of the groups carries structure the tuned kernels hand to the generic ones: zero tails
    # (partially filled blocks), constant stretches, one NaN, one all-zero group
    x[1::3, G // 3:] = 0.0
    x[2::9, 100:1000] = 0.75
    x[4, 17] = float("nan")
    x[6] = 0.0
    x = x.to(dt).reshape(-1)
    c = codec.compress(x, G)
    y = codec.decompress(c)
    torch.cuda.synchronize()
    comp = c.comp_bytes.to(torch.int64) & 0xFFFFFFFF
    sb = c.payload.shape[1]
    pay_sum = 0
    rows = max(1, (1 << 28) // sb)                      # mask in slabs of ~256 MiB
    col = torch.arange(sb, device="cuda:0").unsqueeze(0)
    weights = (torch.arange(sb, device="cuda:0", dtype=torch.int64) % 251 + 1).unsqueeze(0)
    for r0 in range(0, n_groups, rows):
        p = c.payload[r0:r0 + rows].to(torch.int64)
        m = col < comp[r0:r0 + rows].unsqueeze(1)
        pay_sum += int((p * weights * m).sum().item())
    yw = y.view(torch.int16).to(torch.int64).reshape(n_groups, G)
    ew = (torch.arange(G, device="cuda:0", dtype=torch.int64) % 127 + 1).unsqueeze(0)
    out_sum = 0
    for r0 in range(0, n_groups, max(1, (1 << 26) // G)):
        out_sum += int((yw[r0:r0 + max(1, (1 << 26) // G)] * ew).sum().item())
    return {"comp_sum": int(comp.sum().item()), "comp_xor": int(torch.bitwise_xor(comp[::2], comp[1::2]).sum().item()),
            "scale_bits_sum": int((c.scales.view(torch.int32).to(torch.int64) & 0xFFFFFFFF).sum().item()),
            "payload_weighted_sum": pay_sum, "output_weighted_sum": out_sum,
            "launches": codec.stats()["kernel_launches"]}


if __name__ == "__main__":
    print(json.dumps(digests(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])))

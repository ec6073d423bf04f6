#!/usr/bin/env python
"""bench.py -- KV block codec throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

One "step" = one pass of the hot path over one batch of synthetic KV: every group of
the workload is compressed (scale -> wrapped int8 -> delta -> byte-pair RLE) and then
decompressed again, layer by layer.  `value` = uncompressed fp16 KV bytes that made the
round trip per second, whole job (all ranks), inputs resident in HBM.  `e2e` = the same
metric through the host-buffer C ABI (speckv_ext_compress_host / _decompress_host) with
the H2D / D2H copies inside the timed region.

Workloads (BASELINE.json `configs`):
  cfg2  Llama-2-7B KV, batch 32 x 4K ctx per GPU: 32 L x 2 x 32 H x 32 x 4 blocks of
        1024 tok x 128 d = 262144 groups = 64 GiB fp16 per GPU (default; weak scaling)
  cfg3  Llama-2-70B GQA KV, 80 L x 8 KVH x 8K ctx: 10240 groups = 2.5 GiB, layers
        sharded 80/N across ranks (strong scaling)
  cfg4p Llama-3-8B 32K ctx as 4 KiB page groups (2048 elems): 1 Mi groups = 4 GiB
One process per GPU (torchrun), no data-path collective; at N > 1 the only exchange is
an NCCL all-gather of the per-group page-table metadata (comp_bytes) per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

G_BLOCK = 131072          # 1024 tokens x 128 dims: one (layer, head, K|V) block
METRIC = "kv_compress_decompress_roundtrip_throughput"
UNIT = "GB/s"             # 1e9 bytes of uncompressed fp16 KV per second


def workload_spec(name: str, world: int, rank: int, scale: float):
    """-> dict(label, group_elems, n_groups (this rank), chunk_groups, scaling)"""
    if name == "cfg2":
        layers, per_layer = 32, 2 * 32 * 32 * 4          # K|V x heads x batch x (4096/1024)
        layers = max(1, int(round(layers * scale)))
        return dict(label=f"cfg2: Llama-2-7B KV, batch 32 x 4K ctx per GPU ({layers} layers x {per_layer} blocks of 1024x128 fp16)",
                    group_elems=G_BLOCK, n_groups=layers * per_layer, chunk_groups=per_layer, scaling="weak",
                    global_batch=32 * world)
    if name == "cfg3":
        layers, per_layer = 80, 2 * 8 * 8                # K|V x KV heads x (8192/1024)
        mine = [l for l in range(layers) if l % world == rank] if world > 1 else list(range(layers))
        return dict(label=f"cfg3: Llama-2-70B GQA KV, 80 L x 8 KVH x 8K ctx, layers sharded {layers}/{world}",
                    group_elems=G_BLOCK, n_groups=len(mine) * per_layer, chunk_groups=per_layer * 20, scaling="strong",
                    global_batch=1)
    if name == "cfg4p":
        layers, per_layer = 32, 2 * 8 * 32768 * 128 // 2048
        layers = max(1, int(round(layers * scale)))
        return dict(label=f"cfg4p: Llama-3-8B 32K ctx paged KV, {layers} layers x {per_layer} page groups of 2048 fp16",
                    group_elems=2048, n_groups=layers * per_layer, chunk_groups=per_layer, scaling="weak",
                    global_batch=world)
    raise SystemExit(f"unknown workload {name}")


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str):
    """dram bytes per launch from the committed ncu --set full capture, if one matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------
# reference arm: the reference's own C++ engine (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------
def cpu_roundtrip_rate(group_elems: int, budget_s: float, threads: int, seed: int = 1234):
    """Times the reference CPU path (oracle/_ref when it was built here, else the C port) on a
    bounded sample of the workload: round trip of `n` groups of N(0,1) fp16.  -> dict"""
    import numpy as np
    from oracle.oracle import Port, Ref

    rng = np.random.default_rng(seed)
    kind = "reference" if Ref.available() else "port"

    def run(x16, ngroups):
        t0 = time.perf_counter()
        if kind == "reference":
            Ref.roundtrip_batch(x16, 0, group_elems, threads)
        else:
            payload, scales, comp = Port.compress_batch(x16, group_elems, threads=threads)
            Port.decompress_batch(payload, scales, comp, group_elems, 0, threads=threads)
        return time.perf_counter() - t0

    probe_groups = max(threads, 8) if group_elems >= 65536 else max(threads * 16, 256)
    xp = rng.standard_normal(probe_groups * group_elems).astype(np.float16)
    run(xp, probe_groups)                                    # warm-up (page faults, thread start)
    t_probe = run(xp, probe_groups)
    # one pass over <= 512 MiB of KV, repeated until the time budget is used
    cap = max(probe_groups, (512 << 20) // (group_elems * 2))
    n = int(max(probe_groups, min(probe_groups * budget_s / max(t_probe, 1e-6), cap)))
    n = (n // threads) * threads or threads
    x = rng.standard_normal(n * group_elems).astype(np.float16)
    dt, reps = 0.0, 0
    while dt < budget_s and reps < 1000:
        dt += run(x, n)
        reps += 1
    gbs = reps * n * group_elems * 2 / dt / 1e9
    return {"value": gbs, "unit": UNIT, "cores": threads, "kind": kind, "seconds": dt,
            "sample": f"{reps} x {n} groups of {group_elems} fp16 N(0,1) elements ({n * group_elems * 2 / 2**20:.0f} MiB per pass), "
                      f"compress+decompress round trip, {threads} threads, {dt:.1f} s"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    spec = workload_spec(args.workload, world, 0, args.scale)
    threads = os.cpu_count() or 1
    per_step_budget = max(2.0, min(10.0, 90.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_roundtrip_rate(spec["group_elems"], per_step_budget, threads, seed=1234 + i)
        if i >= args.warmup:
            vals.append(r)
    tot_bytes = sum(v["value"] * v["seconds"] for v in vals)
    tot_s = sum(v["seconds"] for v in vals)
    value = tot_bytes / tot_s
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(vals),
            "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None, "dtype": "f32+i8",
            "data": "synthetic N(0,1) fp16 KV (numpy default_rng)",
            "config": {"workload": spec["label"], "group_elems": spec["group_elems"],
                       "note": "reference CPU engine (FPGACacheEngine::compress/decompress) on host cores; "
                               "each step is a bounded sample of the workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": vals[-1]["kind"],
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------
def bind_to_gpu_cpus(local_rank):
    """Pin this rank to the host cores NVML lists as local to its GPU, so that the pinned host
    buffers of the e2e leg are allocated on the GPU's own NUMA node (matters at 8 ranks per box).
    Returns the previous affinity so that the CPU baseline can take all cores back."""
    before = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(before) + 64) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1} & before
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    return before


def run_cuda(args, rank, world, local_rank):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import cxl_speckv_b200 as pkg
    from cxl_speckv_b200 import codec

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    all_cpus = bind_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = pkg.lib()

    spec = workload_spec(args.workload, world, rank, args.scale)
    G, n_groups, cg = spec["group_elems"], spec["n_groups"], spec["chunk_groups"]
    sb = codec.slot_bytes(G)
    n_chunks = (n_groups + cg - 1) // cg
    # ---- resident synthetic KV (N(0,1), seed 1234 + rank), payload slots, one output chunk ------
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    x = torch.empty(n_groups * G, dtype=torch.float16, device=dev)
    for c0 in range(0, n_groups, cg):
        c1 = min(n_groups, c0 + cg)
        x[c0 * G:c1 * G].normal_(generator=gen)
    payload = torch.empty((n_groups, sb), dtype=torch.uint8, device=dev)
    scales = torch.empty(n_groups, dtype=torch.float32, device=dev)
    comp = torch.empty(n_groups, dtype=torch.int32, device=dev)
    out = torch.empty((min(cg, n_groups), G), dtype=torch.float16, device=dev)
    gathered = torch.empty(world * n_groups, dtype=torch.int32, device=dev) if world > 1 else None
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def compress_all():
        for c0 in range(0, n_groups, cg):
            n = min(cg, n_groups - c0)
            st = L.speckv_ext_compress(x.data_ptr() + c0 * G * 2, 0, G, n, payload.data_ptr() + c0 * sb, sb,
                                       scales.data_ptr() + c0 * 4, comp.data_ptr() + c0 * 4, 2, stream)
            assert st == 0, st

    def decompress_all():
        for c0 in range(0, n_groups, cg):
            n = min(cg, n_groups - c0)
            st = L.speckv_ext_decompress(payload.data_ptr() + c0 * sb, sb, scales.data_ptr() + c0 * 4,
                                         comp.data_ptr() + c0 * 4, G, n, 0, out.data_ptr(), None, 2, stream)
            assert st == 0, st

    def step(ev=None):
        if ev:
            ev[0].record()
        compress_all()
        if ev:
            ev[1].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, comp)     # page-table metadata exchange
        if ev:
            ev[2].record()
        decompress_all()
        if ev:
            ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    codec_stats0 = codec.stats()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        step(evs[i])
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = codec.stats()["kernel_launches"] - codec_stats0["kernel_launches"]
    elapsed_ms = t_start.elapsed_time(t_end)
    t_comp = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    t_dec = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    tt = torch.tensor([elapsed_ms, t_comp, t_dec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed_ms, t_comp_max, t_dec_max = tt.tolist()

    kv_bytes = n_groups * G * 2                                   # this rank, per step
    comp_total = int(comp.to(torch.int64).sum().item())           # payload bytes c, this rank
    tot = torch.tensor([kv_bytes, comp_total, n_groups], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    kv_bytes_all, comp_all, groups_all = tot.tolist()
    ms_per_step = elapsed_ms / args.steps
    value = kv_bytes_all / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (per launch = one chunk of cg groups) -------------------
    peak, peak_src = hbm_peak()
    alg_comp = 2 * G * n_groups + comp_total + 12 * n_groups       # 2n read + c write + header
    alg_dec = comp_total + 12 * n_groups + 2 * G * n_groups        # c read + header + 2n write
    comp_gbs = alg_comp / (t_comp * 1e-3) / 1e9
    dec_gbs = alg_dec / (t_dec * 1e-3) / 1e9
    dom = "compress" if t_comp >= t_dec else "decompress"
    dom_bytes_per_launch = (alg_comp if dom == "compress" else alg_dec) / n_chunks
    dom_ms_per_launch = (t_comp if dom == "compress" else t_dec) / n_chunks
    achieved = comp_gbs if dom == "compress" else dec_gbs
    roofline = {"bound": "hbm", "kernel": f"kv_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": ncu_traffic(f"{args.workload}:{dom}"),
                "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
                "compress": {"GB/s": comp_gbs, "frac": comp_gbs / peak, "ms_per_step": t_comp,
                             "kv_GB/s": kv_bytes / (t_comp * 1e-3) / 1e9},
                "decompress": {"GB/s": dec_gbs, "frac": dec_gbs / peak, "ms_per_step": t_dec,
                               "kv_GB/s": kv_bytes / (t_dec * 1e-3) / 1e9},
                "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----------------------
    e2e = None
    if not args.no_e2e:
        ng = min(n_groups, max(1, int(args.e2e_mib * 2**20) // (G * 2)))
        nb_in, nb_pay = ng * G * 2, ng * sb
        h_in = L.speckv_ext_host_alloc(nb_in)
        h_pay = L.speckv_ext_host_alloc(nb_pay)
        h_out = L.speckv_ext_host_alloc(nb_in)
        h_sc = L.speckv_ext_host_alloc(ng * 4)
        h_cb = L.speckv_ext_host_alloc(ng * 4)
        assert h_in and h_pay and h_out and h_sc and h_cb, "pinned host allocation failed"
        x_host = x[:ng * G].cpu().numpy()                      # keep alive across the memmove
        C.memmove(h_in, x_host.ctypes.data, nb_in)
        del x_host

        def e2e_step():
            st = L.speckv_ext_compress_host(h_in, 0, G, ng, h_pay, sb, h_sc, h_cb, 2)
            assert st == 0, st
            st = L.speckv_ext_decompress_host(h_pay, sb, h_sc, h_cb, G, ng, 0, h_out, None, 2)
            assert st == 0, st

        e2e_step()
        e2e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        tb = torch.tensor([float(nb_in)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        e2e = {"value": tb.item() * e2e_steps / te.item() / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": nb_in + nb_pay + 8 * ng, "d2h_bytes_per_step": nb_pay + 8 * ng + nb_in,
               "steps": e2e_steps, "api": "speckv_ext_compress_host + speckv_ext_decompress_host (pinned host buffers)",
               "sample": f"{ng} groups ({nb_in / 2**20:.0f} MiB fp16) per rank per step"}
        # the host path must reproduce the device path bit for bit
        y = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint16)), shape=(ng * G,))
        codec_out = codec.decompress(codec.compress(x[:min(ng, 64) * G], G))
        assert np.array_equal(y[:min(ng, 64) * G], codec_out.view(torch.int16).cpu().numpy().view(np.uint16).ravel())
        for p in (h_in, h_pay, h_out, h_sc, h_cb):
            L.speckv_ext_host_free(p)

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)
        cpu = cpu_roundtrip_rate(G, args.cpu_seconds, os.cpu_count() or 1)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": spec["scaling"], "vs_baseline": None, "dtype": "f32+i8", "data": "synthetic N(0,1) fp16 KV (torch seed 1234+rank)",
                "config": {"workload": spec["label"], "group_elems": G, "groups_per_rank": n_groups,
                           "groups_total": int(groups_all), "kv_bytes_per_step_total": int(kv_bytes_all),
                           "launch_chunk_groups": cg, "global_batch": spec["global_batch"],
                           "parallelism": f"independent shards x{world}, all-gather of comp_bytes only" if world > 1 else "single GPU",
                           "cache": "inputs larger than L2 (no flush needed)" if kv_bytes > (1 << 30) else "inputs + outputs exceed L2 per step"},
                "per_gpu_value": value / world,
                "compression_ratio": {"vs_fp16": kv_bytes_all / comp_all, "vs_fp32_reference_accounting": 2 * kv_bytes_all / comp_all},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# auxiliary reports (not the driver's bench line): BASELINE configs 4 and 5
# --------------------------------------------------------------------------------------
def run_cfg4(args, local_rank):
    """Llama-3-8B 32K-context paged KV: translate every page address, page-table lookup, then
    offload -> restore through the pinned host pool (PCIe GB/s of stored bytes), verify bytes."""
    import ctypes as C

    import numpy as np
    import torch

    import cxl_speckv_b200 as pkg
    from cxl_speckv_b200 import codec
    from cxl_speckv_b200.tier import HostTier

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = pkg.lib()
    layers = max(1, int(round(32 * args.scale)))
    G, pages = 2048, layers * 2 * 8 * 32768 * 128 // 2048
    x = torch.empty(pages * G, dtype=torch.float16, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    x.normal_(generator=gen)

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # address translation of every page + page-table lookup through the exported table
    assert L.speckv_init(f"cuda:{local_rank}".encode()) == 0
    h = C.c_uint64(); assert L.speckv_alloc(pages * 4096, None, C.byref(h)) == 0
    tbl = torch.empty(pages * 3, dtype=torch.int64, device=dev); cnt = C.c_size_t()
    assert L.speckv_ext_page_table_export(h.value, tbl.data_ptr(), pages, C.byref(cnt), None) == 0
    va = (torch.arange(pages, dtype=torch.int64, device=dev) << 12) + (h.value << 32) + 0x123
    pa = torch.empty_like(va); fl = torch.empty(pages, dtype=torch.int32, device=dev)
    t_tr = timed(lambda: codec.translate(va, out=pa))
    t_lk = timed(lambda: L.speckv_ext_page_lookup(tbl.data_ptr(), pages, h.value << 32, va.data_ptr(), pa.data_ptr(), fl.data_ptr(), pages, None))
    assert int(pa[5].item()) == 0x4000000000 + (h.value << 20) + (5 << 12) + 0x123
    L.speckv_finalize()

    tier = HostTier(int(pages * G * 2 * 1.02) + (64 << 20))
    ids = np.arange(pages, dtype=np.uint64) << np.uint64(12)
    tier.offload(x[:65536 * G], G, ids[:65536]); tier.drop(ids[:65536])      # warm-up: staging buffers, pinned mirrors
    t0 = time.perf_counter(); tier.offload(x, G, ids); torch.cuda.synchronize(); t_off = time.perf_counter() - t0
    out = torch.empty((pages, G), dtype=torch.float16, device=dev)
    st_off = tier.stats()
    tier.restore(ids[:65536], G, torch.float16, out=out[:65536])             # warm-up
    t0 = time.perf_counter(); tier.restore(ids, G, torch.float16, out=out); torch.cuda.synchronize(); t_res = time.perf_counter() - t0
    st = tier.stats()
    want = codec.decompress(codec.compress(x[:4096 * G], G))
    assert torch.equal(out[:4096].view(torch.int16), want.view(torch.int16))
    tier.close()
    del out

    # residency policy over the same pages: one decode step touches every page; then a prefetch-sized promotion
    from cxl_speckv_b200.tier import TierPolicy
    pol = TierPolicy(pages, l1_pages=pages // 16)
    allp = np.arange(pages, dtype=np.uint64)
    pol.place(allp[:pages // 16], 0); pol.place(allp[pages // 16:], 2)
    perm = torch.randperm(pages, device=dev, generator=gen)
    t_touch = timed(lambda: pol.touch(perm))
    batch = allp[pages // 16:][:4096]
    torch.cuda.synchronize(); t0 = time.perf_counter(); ok, ev = pol.promote(batch); t_prom = time.perf_counter() - t0
    assert ok.all() and ev.size == 4096
    pol.close()
    print(json.dumps({"report": "cfg4", "workload": f"Llama-3-8B 32K ctx paged KV, {layers} layers, {pages} pages of 4 KiB",
                      "policy": {"touch_Gpages_per_s": pages / t_touch / 1e6, "touch_ms_all_pages": t_touch,
                                 "promote_4096_with_eviction_ms": t_prom * 1e3},
                      "translate_Gaddr_per_s": pages / t_tr / 1e6, "translate_GB/s": pages * 16 / t_tr / 1e6,
                      "page_lookup_Gaddr_per_s": pages / t_lk / 1e6,
                      "offload": {"s": t_off, "stored_bytes": st_off["last_offload_stored_bytes"], "pcie_GB/s": st_off["last_offload_stored_bytes"] / t_off / 1e9,
                                  "kv_GB/s": pages * G * 2 / t_off / 1e9},
                      "restore": {"s": t_res, "pcie_GB/s": st["last_restore_stored_bytes"] / t_res / 1e9, "kv_GB/s": pages * G * 2 / t_res / 1e9},
                      "verified": "restored pages == device compress->decompress, bit for bit"}), flush=True)


def run_cfg5(args, rank, world, local_rank):
    """Batch-256 decode step: LSTM prefetch scoring (k = 4) + decompress of the predicted blocks.
    N GPUs: the 256 sequences split over the ranks; every rank scores its share, the PrefetchRequest records
    (32 bytes each) are all-gathered over NCCL -- the only collective -- and every rank
    decodes the predicted blocks it owns (stored blocks shard round-robin, like layers)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from cxl_speckv_b200 import codec, prefetch, sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(1)
    emb = ((rng.random((32000, 64), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
    wout = ((rng.random((32000, 128), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
    prefetch.load_predictor(emb, wout)
    batch, k = 256, 4
    all_toks = np.random.default_rng(7).integers(0, 32000, (batch, 16)).astype(np.int32)
    per = (batch + world - 1) // world
    toks = torch.from_numpy(all_toks[rank * per:(rank + 1) * per]).to(dev)
    G, n_blocks = 131072, 4096                                    # 1 GiB of stored blocks to pick from, in all
    local_blocks = n_blocks // world                              # block b lives on rank b % world as local block b // world
    x = torch.empty(local_blocks * G, device=dev, dtype=torch.float16).normal_()
    c = codec.compress(x, G)
    out = torch.empty((batch * k, G), dtype=torch.float16, device=dev)   # worst case: every prediction lands here

    def step():
        ids, conf, va = prefetch.score(toks, k=k, layer_id=0)
        if world == 1:
            blocks = (ids.view(-1) % n_blocks).to(torch.int32)
            codec.decompress_indexed(c, blocks, out=out[:blocks.numel()])
            return blocks.numel()
        rec = sharding.pack_prefetch_requests(va, 0, ids, conf)
        allrec = sharding.gather_records(rec, sharding.PREFETCH_RECORD_BYTES)          # [world, per * k, 32]
        _, _, tok, _ = sharding.unpack_prefetch_requests(allrec.view(-1, sharding.PREFETCH_RECORD_BYTES))
        blocks = tok.to(torch.int64) % n_blocks
        mine = blocks[blocks % world == rank] // world
        n = int(mine.numel())                                      # host sync: the size of this rank's decode batch
        if n:
            codec.decompress_indexed(c, mine.to(torch.int32), out=out[:n])
        return n

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b, m = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 10
    a.record()
    for _ in range(reps):
        prefetch.score(toks, k=k, layer_id=0)
    m.record()
    decoded = 0
    for _ in range(reps):
        decoded += step()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(m) / reps, m.elapsed_time(b) / reps, float(decoded) / reps], device=dev, dtype=torch.float64)
    tot = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                    # times: max over ranks
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)                  # decoded blocks: sum over ranks
    t_score, t_step, n_dec = float(t[0]), float(t[1]), float(tot[2])
    if rank == 0:
        print(json.dumps({"report": "cfg5", "n_gpus": world,
                          "workload": "256 sequences x 16-token history, vocab 32000, k=4 -> 1024 predicted blocks of 1024x128 fp16",
                          "score_ms": t_score, "score_seq_per_s": batch / t_score * 1e3, "reference_cpu_ms_per_sequence": 17.8,
                          "step_ms(score+exchange+decompress)": t_step, "blocks_decoded_per_step": n_dec,
                          "step_seq_per_s": batch / t_step * 1e3,
                          "decompress_kv_GB/s": n_dec * G * 2 / (max(t_step - t_score, 1e-9) * 1e-3) / 1e9,
                          "exchange": "none (1 GPU)" if world == 1 else f"all-gather of {batch * k} PrefetchRequest records (32 B) over NCCL"}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ratios(args, local_rank):
    """Compression ratio and codec throughput per value distribution:
    N(0,1) (the headline: incompressible for this format), all-zero, runs of 300, bf16 N(0,1)."""
    import torch

    from cxl_speckv_b200 import codec

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    G, n_groups = G_BLOCK, 2048                                   # 512 MiB of fp16 per distribution
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    base = torch.empty(n_groups * G, dtype=torch.float16, device=dev).normal_(generator=gen)
    dists = {
        "normal_fp16": base,
        "normal_bf16": base.to(torch.bfloat16),
        "zeros_fp16": torch.zeros_like(base),
        "runs300_fp16": base[: (n_groups * G + 299) // 300].repeat_interleave(300)[: n_groups * G].contiguous(),
    }
    out = {}
    for name, x in dists.items():
        c = codec.compress(x, G)
        y = codec.decompress(c)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(5):
            codec.compress(x, G, out=c)
        ev[1].record()
        for _ in range(5):
            codec.decompress(c, out=y)
        ev[2].record(); torch.cuda.synchronize()
        tc, td = ev[0].elapsed_time(ev[1]) / 5, ev[1].elapsed_time(ev[2]) / 5
        cb = float(c.comp_bytes.to(torch.int64).sum().item())
        raw = n_groups * G * 2
        mse = ((y.float() - x.view(n_groups, G).float()) ** 2).mean().item()
        out[name] = {"ratio_vs_fp16": raw / cb, "ratio_fp32_reference_accounting": 2 * raw / cb,
                     "compress_kv_GB/s": raw / tc / 1e6, "decompress_kv_GB/s": raw / td / 1e6,
                     "compress_algorithmic_GB/s": (raw + cb) / tc / 1e6, "decompress_algorithmic_GB/s": (raw + cb) / td / 1e6,
                     "roundtrip_mse": mse}
    print(json.dumps({"report": "ratios", "workload": f"{n_groups} groups of 1024x128 per distribution", "distributions": out}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4p", "cfg4", "cfg5", "ratios"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the workload's layers (debug only; 1.0 = the named config)")
    ap.add_argument("--e2e-mib", type=float, default=2048.0, help="host-buffer sample per e2e step (MiB of fp16 KV)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "cfg4":
        run_cfg4(args, local_rank)
    elif args.workload == "cfg5":
        run_cfg5(args, rank, world, local_rank)
    elif args.workload == "ratios":
        run_ratios(args, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""
bench.py -- headline benchmark of the compress/decompress hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--layers L]

Workload (configs[1] of BASELINE.json): W4A16 group_size=128 quantize + pack_to_int32 over every
Linear weight of a Llama-3-8B-shaped model (32 layers x {q,k,v,o,gate,up,down} = 224 bf16 tensors,
6.98 G elements, 13.96 GB), synthetic N(0, 0.02^2) weights, scales from the min/max observer rule.
One "step" = one pass of the hot path over all 224 tensors (a single multi-tensor launch).
`value` = weight bytes processed per second with tensors resident in HBM; `e2e` = the same pass
through the compressor plugin API on HOST (pinned) state dicts, H2D and D2H inside the timed region.
The inputs (14 GB) are far larger than the 126 MB L2, so every step streams from HBM.

Under torchrun (N > 1) every rank owns its own full-size tensor set (weak scaling, no data-path
collective); timing = max over ranks.

--impl reference times the CPU restatement of the reference path (oracle/, plain C + OpenMP on all
host threads; the reference itself is Python and cannot travel to the GPU box) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "weight_GBps_w4a16_g128_quantize_pack_llama3_8b"
UNIT = "GB/s"
LAYER_SHAPES = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
GROUP = 128
BITS = 4
ALG_BYTES_PER_ELEM = 2 + BITS / 8 + 2 / GROUP          # bf16 in + packed out + bf16 scale = 2.515625
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def profile_traffic(kernel: str):
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


def traffic_per_launch(kernel: str, n_elems: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture
    (profiles/ncu_ops_r1.md, taken on 4 layers), scaled by element count to this launch's size"""
    t = profile_traffic(kernel)
    if not t:
        return None
    return int(t["dram_bytes_per_launch"] / t["elements_per_launch"] * n_elems)


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every few ms during the timed region
    (nvidia-smi itself takes ~100 ms per query, too coarse for a sub-second region)"""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self._h is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_weights(device, layers: int, seed0: int):
    """bf16 weights + observer scales, generated on the device"""
    ws, scs = [], []
    for li in range(layers):
        for ti, (r, c) in enumerate(LAYER_SHAPES):
            g = torch.Generator(device=device).manual_seed(seed0 + li * len(LAYER_SHAPES) + ti)
            w = torch.empty(r, c, dtype=torch.bfloat16, device=device)
            step = 2048
            for r0 in range(0, r, step):  # bounded fp32 temporaries
                w[r0:r0 + step] = (torch.randn(min(step, r - r0), c, device=device, generator=g) * 0.02).bfloat16()
            # calculate_qparams, symmetric int4: scale = max|w| / 7.5 in the weight dtype (utils/helpers.py:79-87)
            sc = (w.unflatten(-1, (-1, GROUP)).abs().amax(-1).float() / 7.5).bfloat16()
            ws.append(w)
            scs.append(sc)
    return ws, scs


def args_w4():
    from types import SimpleNamespace
    return SimpleNamespace(strategy="group", group_size=GROUP, block_structure=None, num_bits=BITS, type="int", symmetric=True)


def time_steps(fn, steps: int, warmup: int, dist_on: bool, sampler=None):
    """W untimed steps, then exactly K steps bracketed by barrier + synchronize; CUDA events on the launching stream"""
    import torch.distributed as dist

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.__enter__()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.__exit__()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_b200(a):
    import torch.distributed as dist

    from compressed_tensors_b200 import _native as N
    from compressed_tensors_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    if a.gpus != world:
        print(f"[bench] note: --gpus {a.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    layers = a.layers
    ws, scs = make_weights(dev, layers, 1000 + rank * 100000)
    n_elems = sum(w.numel() for w in ws)
    weight_bytes = n_elems * 2
    alg_bytes = n_elems * ALG_BYTES_PER_ELEM
    qargs = args_w4()

    # device-resident problem table for the multi-tensor launch
    outs = [torch.empty(w.shape[0], w.shape[1] * BITS // 32, dtype=torch.int32, device=dev) for w in ws]
    probs = []
    for w, sc, o in zip(ws, scs, outs):
        p = ops._resolve(w, sc, None, qargs, None)
        d = ops._desc(p, w.dtype, sc.dtype, None, torch.bfloat16, torch.int8, None, N.Q_INT, BITS)
        probs.append((d, w, sc, None, o))

    def step():
        ops.batched(N.OP_QUANTIZE_PACK, probs, local)

    l0 = N.launch_count()
    cs = ClockSampler(local)
    ms = time_steps(step, a.steps, a.warmup, dist_on, sampler=cs)
    launches = (N.launch_count() - l0) - a.warmup  # one launch per step
    clocks = cs.summary()
    ms_per_step = ms / a.steps
    value = world * weight_bytes / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = peaks()
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9

    # secondary ops of the metric (same tensors), device-resident, reported beside the headline
    extra = {}
    if not a.no_extra:
        def rate(fn, bytes_alg, nelem_bytes):
            # secondary ops: one CUDA-event pair per launch, median over the launches (a host hiccup between two
            # sub-millisecond launches would otherwise dominate a short timed region); the headline keeps the contract's
            # single bracket around exactly K steps
            k = max(5, a.steps)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
            for e0, e1 in ev:
                e0.record()
                fn()
                e1.record()
            torch.cuda.synchronize()
            ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
            t = ms[len(ms) // 2]
            return {"weight_GBps": round(nelem_bytes / (t * 1e-3) / 1e9, 1), "hbm_GBps": round(bytes_alg / (t * 1e-3) / 1e9, 1),
                    "frac_of_peak": round(bytes_alg / (t * 1e-3) / 1e9 / peak, 3), "ms": round(t, 3), "ms_worst": round(ms[-1], 3), "launches": k}

        # decompress: unpack + dequantize
        dq = [torch.empty_like(w) for w in ws]
        dprobs = []
        for w, sc, o, dst in zip(ws, scs, outs, dq):
            p = ops._resolve(torch.empty(w.shape, dtype=torch.int8, device="meta"), sc, None, qargs, None)
            d = ops._desc(p, None, sc.dtype, None, None, torch.int8, torch.bfloat16, N.Q_INT, BITS)
            dprobs.append((d, o, sc, None, dst))
        extra["w4a16_unpack_dequantize"] = rate(lambda: ops.batched(N.OP_UNPACK_DEQUANTIZE, dprobs, local), alg_bytes, weight_bytes)
        del dq, dprobs
        # FP8 per-tensor quantize / dequantize (configs[2])
        from types import SimpleNamespace
        f8 = SimpleNamespace(strategy="tensor", group_size=None, block_structure=None, num_bits=8, type="float", symmetric=True)
        s8 = [(w.abs().max().float() / 448).bfloat16().reshape(1) for w in ws]
        q8 = [torch.empty(w.shape, dtype=torch.float8_e4m3fn, device=dev) for w in ws]
        qprobs, dqprobs = [], []
        back = [torch.empty_like(w) for w in ws]
        for w, sc, q, b in zip(ws, s8, q8, back):
            p = ops._resolve(w, sc, None, f8, None)
            qprobs.append((ops._desc(p, w.dtype, sc.dtype, None, torch.bfloat16, torch.float8_e4m3fn, None, N.Q_FLOAT, 8), w, sc, None, q))
            dqprobs.append((ops._desc(p, None, sc.dtype, None, None, torch.float8_e4m3fn, torch.bfloat16, N.Q_INT, 8), q, sc, None, b))
        extra["fp8_quantize"] = rate(lambda: ops.batched(N.OP_QUANTIZE, qprobs, local), n_elems * 3.0, weight_bytes)
        extra["fp8_dequantize"] = rate(lambda: ops.batched(N.OP_DEQUANTIZE, dqprobs, local), n_elems * 3.0, weight_bytes)
        del q8, back, qprobs, dqprobs
        # standalone int4 pack / unpack on int8 codes (one big tensor set: largest shape x 8)
        codes = [torch.randint(-8, 8, (14336, 4096), dtype=torch.int8, device=dev) for _ in range(32)]  # 1.88 G codes: 2.8 GB of traffic per launch
        nel = sum(c.numel() for c in codes)
        pk = [torch.empty(c.shape[0], c.shape[1] // 8, dtype=torch.int32, device=dev) for c in codes]
        pdesc = []
        for c in codes:
            d = N.QuantDesc()
            d.rows, d.cols, d.num_bits = c.shape[0], c.shape[1], 4
            pdesc.append(d)
        pack_probs = [(d, c, None, None, o) for d, c, o in zip(pdesc, codes, pk)]
        unpack_probs = [(d, o, None, None, c) for d, c, o in zip(pdesc, codes, pk)]
        extra["int4_pack"] = rate(lambda: ops.batched(N.OP_PACK_INT32, pack_probs, local), nel * 1.5, nel)
        extra["int4_unpack"] = rate(lambda: ops.batched(N.OP_UNPACK_INT32, unpack_probs, local), nel * 1.5, nel)
        del codes, pk
        # NVFP4: fp4 e2m1, groups of 16, bf16 group scales (fp8-representable, as after calibration) and a
        # float32 global scale per tensor; decompress reads the scales as stored (float8_e4m3fn)
        nv = SimpleNamespace(strategy="tensor_group", group_size=16, block_structure=None, num_bits=4, type="float", symmetric=True)
        gss = [(448.0 * 6.0 / w.abs().max().float()).reshape(1) for w in ws]
        s8s = [(w.unflatten(-1, (-1, 16)).abs().amax(-1).float() / 6.0 * g).clamp(2.0 ** -9, 448.0).to(torch.float8_e4m3fn) for w, g in zip(ws, gss)]
        sbs = [s.to(torch.bfloat16) for s in s8s]
        nib = [torch.empty(w.shape[0], w.shape[1] // 2, dtype=torch.uint8, device=dev) for w in ws]
        nback = [torch.empty_like(w) for w in ws]
        cprobs, uprobs = [], []
        for w, sb, s8_, g, o, b in zip(ws, sbs, s8s, gss, nib, nback):
            p = ops._resolve(w, sb, None, nv, None)
            d = ops._desc(p, w.dtype, sb.dtype, None, torch.float32, w.dtype, None, N.Q_FP4, 4, torch.float32)
            d.global_scale = g.data_ptr()
            cprobs.append((d, w, sb, None, o))
            d2 = ops._desc(p, None, torch.float32, None, None, None, torch.bfloat16, N.Q_FP4, 4, torch.float32)
            d2.scale_dtype = N.DT[torch.float8_e4m3fn]
            d2.global_scale = g.data_ptr()
            uprobs.append((d2, o, s8_, None, b))
        extra["nvfp4_quantize_pack"] = rate(lambda: ops.batched(N.OP_QUANTIZE_PACK_FP4, cprobs, local), n_elems * (2 + 2 / 16 + 0.5), weight_bytes)
        extra["nvfp4_unpack_dequantize"] = rate(lambda: ops.batched(N.OP_UNPACK_DEQUANTIZE_FP4, uprobs, local), n_elems * (0.5 + 1 / 16 + 2), weight_bytes)
        del nib, nback, cprobs, uprobs, sbs, s8s

    # end to end through the plugin API on host (pinned) state dicts
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(ws, scs, a, dist_on, world)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu = cpu_baseline(max_seconds=20.0)

    if dist_on:
        dist.barrier()
    if rank == 0:
        tn = N.lib()
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"W4A16 g128 symmetric quantize+pack_to_int32, Llama-3-8B-shaped Linear weights, {layers} layers x 7 = {len(ws)} bf16 tensors per GPU, {n_elems/1e9:.3f} G elements",
                       "l2": "inputs (%.1f GB per step) >> 126 MB L2, no flush needed" % (weight_bytes / 1e9),
                       "launch": "one multi-tensor persistent launch per step", "pipe": os.environ.get("CT_B200_PIPE", "tma")},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic_per_launch("quantize_pack", n_elems), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "frac_of_8TBps_nominal": round(achieved / 8000.0, 4)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra:
            line["ops"] = extra
        print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()


def _bind_to_gpu_numa(device_index: int):
    """Opt-in experiment (CT_BENCH_NUMA=1, default off, not part of any reported number yet): run this rank on the CPUs NVML names as
    local to its GPU, so that the pinned host buffers of the e2e leg are first-touched on that socket.  Returns the previous affinity
    (to restore) or None when anything is missing."""
    if os.environ.get("CT_BENCH_NUMA", "0") != "1" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        before = os.sched_getaffinity(0)
        cpus &= before
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return before
    except Exception:  # noqa: BLE001  (no NVML, no permission: keep the launcher's affinity)
        return None


def run_e2e(ws, scs, a, dist_on, world):
    """The same pass end to end through the public API: ModelCompressor.compress_model() on a HOST-resident
    model (pinned weights and scales), i.e. the call llm-compressor makes before save_pretrained.  Every step
    uploads all weights + scales (H2D), runs the kernels and brings the packed words back (D2H); the timed
    region is the compress_model call itself.  Modules are reset to their uncompressed state between steps
    (pointer swaps, untimed)."""
    from compressed_tensors_b200.compressors import ModelCompressor
    from compressed_tensors_b200.quantization import QuantizationConfig, QuantizationStatus, apply_quantization_config
    from compressed_tensors_b200.utils import replace_direct_state_dict

    layers = min(a.e2e_layers, len(ws) // len(LAYER_SHAPES))
    n = layers * len(LAYER_SHAPES)
    previous_affinity = _bind_to_gpu_numa(ws[0].device.index or 0)
    hw = [w.cpu().pin_memory() for w in ws[:n]]
    hs = [s.cpu().pin_memory() for s in scs[:n]]
    wbytes = sum(t.numel() * 2 for t in hw)
    h2d = wbytes + sum(t.numel() * 2 for t in hs)
    d2h = sum(t.numel() // 8 * 4 for t in hw)

    model = torch.nn.Module()
    mods = []
    for i, w in enumerate(hw):
        lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device="meta", dtype=torch.bfloat16)
        lin.weight = torch.nn.Parameter(w, requires_grad=False)
        model.add_module(f"linear_{i}", lin)
        mods.append(lin)
    apply_quantization_config(model, QuantizationConfig(config_groups={"W4A16": ["Linear"]}))
    mc = ModelCompressor.from_pretrained_model(model)

    def reset():
        for lin, w, sc in zip(mods, hw, hs):
            replace_direct_state_dict(lin, {"weight": w, "weight_scale": sc})
            lin.quantization_status = QuantizationStatus.FROZEN
        mc.remove_decompression_hook(model)

    steps = max(2, min(a.steps, 5))
    times = []
    for k in range(2 + steps):
        reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mc.compress_model(model, distributed=False)   # every rank holds its own shard of the job (weak scaling), no exchange
        torch.cuda.synchronize()
        if k >= 2:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    if dist_on:
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    packed = mods[0].weight_packed
    assert packed.dtype == torch.int32 and not packed.is_cuda and mods[-1].quantization_status == QuantizationStatus.COMPRESSED
    out = {"value": round(world * wbytes / dt / 1e9, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "api": "ModelCompressor.compress_model(model) on a host-resident (pinned) model", "tensors_per_step": n, "steps": steps,
           "ms_per_step": round(dt * 1e3, 2)}
    if previous_affinity is not None:
        out["cpu_affinity"] = "GPU-local CPUs (CT_BENCH_NUMA=1)"
        os.sched_setaffinity(0, previous_affinity)      # the CPU baseline leg counts its threads from the launcher's affinity
    return out


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port (plain C + OpenMP), bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def oracle_compress_layer(ws, scs, tmp):
    """quantize(int8) -> pack_to_int32 for each tensor of the sample, through the C oracle"""
    import ctypes

    import oracle

    L = oracle.lib()
    for w, sc, (q8, out) in zip(ws, scs, tmp):
        r, c = w.shape
        L.orc_quantize_pack(oracle._p(w), 2, oracle._p(sc), 2, ctypes.c_void_p(0), -1, ctypes.c_void_p(0), oracle._p(out), oracle._p(q8),
                            ctypes.c_int64(r), ctypes.c_int64(c), ctypes.c_int64(1), ctypes.c_int64(GROUP), ctypes.c_int64(c // GROUP), 2, BITS)


def cpu_sample():
    """one layer of the workload (7 tensors, 218 M elements, 436 MB of bf16)"""
    ws, scs, tmp = [], [], []
    for ti, (r, c) in enumerate(LAYER_SHAPES):
        g = torch.Generator().manual_seed(1000 + ti)
        w = (torch.randn(r, c, generator=g) * 0.02).bfloat16()
        sc = (w.unflatten(-1, (-1, GROUP)).abs().amax(-1).float() / 7.5).bfloat16()
        ws.append(w)
        scs.append(sc)
        tmp.append((torch.empty(r, c, dtype=torch.int8), torch.empty(r, c // 8, dtype=torch.int32)))
    return ws, scs, tmp


def cpu_baseline(max_seconds: float):
    import oracle

    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    ws, scs, tmp = cpu_sample()
    wbytes = sum(w.numel() * 2 for w in ws)
    oracle_compress_layer(ws, scs, tmp)  # warm-up (also builds the .so)
    t0 = time.perf_counter()
    reps = 0
    while True:
        oracle_compress_layer(ws, scs, tmp)
        reps += 1
        if time.perf_counter() - t0 > max_seconds / 2 or reps >= 20:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(wbytes / dt / 1e9, 3), "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"oracle/ct_oracle.c (C + OpenMP) on 1 of 32 layers (7 tensors, {wbytes/1e6:.0f} MB bf16), {reps} reps, {dt*1e3:.0f} ms each"}


def run_reference(a):
    """reference arm: the CPU restatement of the reference path on the host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time a 1-thread baseline
    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    ws, scs, tmp = cpu_sample()
    wbytes = sum(w.numel() * 2 for w in ws)
    for _ in range(max(1, min(a.warmup, 3))):
        oracle_compress_layer(ws, scs, tmp)
    steps = max(1, a.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        oracle_compress_layer(ws, scs, tmp)
        done += 1
        if time.perf_counter() - t0 > 150:
            break
    dt = (time.perf_counter() - t0) / done
    v = round(wbytes / dt / 1e9, 3)
    sample = f"1 of 32 Llama-3-8B layers per step (7 tensors, {wbytes/1e6:.0f} MB bf16), oracle/ct_oracle.c C+OpenMP port of the reference's torch-eager path"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": done, "warmup": a.warmup,
        "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "W4A16 g128 symmetric quantize+pack_to_int32, Llama-3-8B-shaped Linear weights (bounded sample: 1 layer per step)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=32, help="Llama-3-8B layers per GPU (32 = the full model)")
    ap.add_argument("--e2e-layers", type=int, default=32)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()

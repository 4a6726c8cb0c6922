#!/usr/bin/env python
"""
Tuning sweep on one B200: times the headline multi-tensor ops for a list of
(pipe, stages, ctas_per_sm) settings on `--layers` Llama-3-8B layers (>> L2) and prints one JSON
line per setting.  Run under gpurun; results go to gpurun_out/sweep.jsonl.
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from compressed_tensors_b200 import _native as N  # noqa: E402
from compressed_tensors_b200 import ops  # noqa: E402


def build_problems(layers: int):
    dev = torch.device("cuda", 0)
    ws, scs = bench.make_weights(dev, layers, 1000)
    n = sum(w.numel() for w in ws)
    qa = bench.args_w4()
    outs = [torch.empty(w.shape[0], w.shape[1] // 8, dtype=torch.int32, device=dev) for w in ws]
    back = [torch.empty_like(w) for w in ws]
    f8 = SimpleNamespace(strategy="tensor", group_size=None, block_structure=None, num_bits=8, type="float", symmetric=True)
    s8 = [(w.abs().max().float() / 448).bfloat16().reshape(1) for w in ws]
    q8 = [torch.empty(w.shape, dtype=torch.float8_e4m3fn, device=dev) for w in ws]
    P = {"quantpack": [], "unpackdeq": [], "fp8_q": [], "fp8_dq": [], "fake_w4": [], "observe_qp": [], "nvfp4_qp": [], "nvfp4_ud": []}
    nv = SimpleNamespace(strategy="tensor_group", group_size=16, block_structure=None, num_bits=4, type="float", symmetric=True)
    keep = []
    sc_out = [torch.empty_like(s) for s in scs]
    for w, sc, o, b, s, q, so in zip(ws, scs, outs, back, s8, q8, sc_out):
        p = ops._resolve(w, sc, None, qa, None)
        P["observe_qp"].append((ops._desc(p, w.dtype, sc.dtype, None, torch.bfloat16, torch.int8, None, N.Q_INT, 4), w, so, None, o))
        P["quantpack"].append((ops._desc(p, w.dtype, sc.dtype, None, torch.bfloat16, torch.int8, None, N.Q_INT, 4), w, sc, None, o))
        P["unpackdeq"].append((ops._desc(p, None, sc.dtype, None, None, torch.int8, torch.bfloat16, N.Q_INT, 4), o, sc, None, b))
        P["fake_w4"].append((ops._desc(p, w.dtype, sc.dtype, None, torch.bfloat16, None, torch.bfloat16, N.Q_INT, 4), w, sc, None, b))
        p8 = ops._resolve(w, s, None, f8, None)
        P["fp8_q"].append((ops._desc(p8, w.dtype, s.dtype, None, torch.bfloat16, torch.float8_e4m3fn, None, N.Q_FLOAT, 8), w, s, None, q))
        g = (448.0 * 6.0 / w.abs().max().float()).reshape(1)
        s8n = (w.unflatten(-1, (-1, 16)).abs().amax(-1).float() / 6.0 * g).clamp(2.0 ** -9, 448.0).to(torch.float8_e4m3fn)
        sbn = s8n.to(torch.bfloat16)
        nib = torch.empty(w.shape[0], w.shape[1] // 2, dtype=torch.uint8, device=dev)
        pn = ops._resolve(w, sbn, None, nv, None)
        dn = ops._desc(pn, w.dtype, sbn.dtype, None, torch.float32, w.dtype, None, N.Q_FP4, 4, torch.float32)
        dn.global_scale = g.data_ptr()
        du = ops._desc(pn, None, torch.float32, None, None, None, torch.bfloat16, N.Q_FP4, 4, torch.float32)
        du.scale_dtype = N.DT[torch.float8_e4m3fn]
        du.global_scale = g.data_ptr()
        keep.append(g)
        P["nvfp4_qp"].append((dn, w, sbn, None, nib))
        P.setdefault("nvfp4_observe_qp", []).append((dn, w, torch.empty_like(s8n), None, nib))
        P["nvfp4_ud"].append((du, nib, s8n, None, b))
        P["fp8_dq"].append((ops._desc(p8, None, s.dtype, None, None, torch.float8_e4m3fn, torch.bfloat16, N.Q_INT, 8), q, s, None, b))
    OPS = {"quantpack": (N.OP_QUANTIZE_PACK, 2.515625), "unpackdeq": (N.OP_UNPACK_DEQUANTIZE, 2.515625),
           "observe_qp": (N.OP_OBSERVE_QUANTIZE_PACK, 2.515625), "fp8_q": (N.OP_QUANTIZE, 3.0), "fp8_dq": (N.OP_DEQUANTIZE, 3.0), "fake_w4": (N.OP_FAKE_QUANTIZE, 4.0 + 2 / 128),
           "nvfp4_qp": (N.OP_QUANTIZE_PACK_FP4, 2 + 2 / 16 + 0.5), "nvfp4_observe_qp": (N.OP_OBSERVE_QUANTIZE_PACK_FP4, 2 + 1 / 16 + 0.5), "nvfp4_ud": (N.OP_UNPACK_DEQUANTIZE_FP4, 0.5 + 1 / 16 + 2)}
    P["_keep"] = keep   # global-scale tensors referenced by address from the descriptors
    return P, OPS, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--configs", default="tma:4:3,tma:3:4,tma:2:5,tma:2:6,tma:6:2,tma:8:2,direct:0:4,direct:0:6,direct:0:8")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    a = ap.parse_args()
    torch.cuda.set_device(0)
    P, OPS, n = build_problems(a.layers)
    peak, _ = bench.peaks()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as f:
        for cfg in a.configs.split(","):
            pipe, stages, ctas = cfg.split(":")
            N.set_tuning({"tma": 1, "tmas": 2}.get(pipe, 0), int(stages) or 4, int(ctas))   # tmas = TMA ring, static deal
            rec = {"pipe": pipe, "stages": int(stages), "ctas_per_sm": int(ctas)}
            for name, (op, bpe) in OPS.items():
                ts = sorted(bench.time_steps(lambda: ops.batched(op, P[name], 0), a.steps, 3, False) / a.steps for _ in range(5))
                rec[name] = round(n * bpe / (ts[0] * 1e-3) / 1e9, 1)          # best of 5
                rec[name + "_med"] = round(n * bpe / (ts[2] * 1e-3) / 1e9, 1)  # median of 5
                rec[name + "_frac"] = round(rec[name + "_med"] / peak, 3)
            line = json.dumps(rec)
            print(line, flush=True)
            f.write(line + "\n")


if __name__ == "__main__":
    main()

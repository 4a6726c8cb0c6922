#!/usr/bin/env python
"""
Golden vectors for the checkpoint converters, produced by importing the reference like make_golden.py:

  AutoAWQConverter.process            entrypoints/convert/converters/autoawq.py:109-129, 179-262
  FP8BlockDequantizer._create_dequantized_weight   converters/fp8block_dequantizer.py:111-158
  CompressedTensorsDequantizer.process (pack-quantized W4A16 asym group, int8 channel, NVFP4)   converters/ct_dequantizer.py:63-99

    python tests/golden/make_golden_convert.py   ->  tests/golden/convert.pt.gz
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference from a temp copy)

from compressed_tensors.entrypoints.convert import AutoAWQConverter, FP8BlockDequantizer  # noqa: E402


def main():
    g = torch.Generator().manual_seed(777)
    out = {"awq": [], "fp8block": []}
    for (k, n, gsz, zp) in ((64, 256, 32, True), (128, 64, 128, True), (200, 24, 8, False), (8, 8, 8, True), (136, 520, 8, True)):
        qweight = torch.randint(-2**31, 2**31 - 1, (k, n // 8), generator=g, dtype=torch.int64).to(torch.int32)
        qzeros = torch.randint(-2**31, 2**31 - 1, (k // gsz, n // 8), generator=g, dtype=torch.int64).to(torch.int32)
        scales = (torch.rand(k // gsz, n, generator=g) * 0.02).to(torch.float16)
        conv = AutoAWQConverter(group_size=gsz, zero_point=zp)
        t = {"m.q_proj.qweight": qweight.clone(), "m.q_proj.scales": scales.clone()}
        if zp:
            t["m.q_proj.qzeros"] = qzeros.clone()
        res = conv.process(t)
        out["awq"].append(dict(qweight=qweight, qzeros=qzeros if zp else None, scales=scales, group_size=gsz, zero_point=zp,
                               result={k_: v for k_, v in res.items()}))
    for (r, c, bs, dt) in ((256, 256, (128, 128), torch.bfloat16), (200, 300, (128, 128), torch.bfloat16), (96, 136, (32, 64), torch.float16),
                           (130, 8, (128, 128), torch.bfloat16)):
        w = (torch.randn(r, c, generator=g) * 3).to(torch.float8_e4m3fn)
        s = torch.randn(-(-r // bs[0]), -(-c // bs[1]), generator=g).abs() * 0.01 + 1e-4
        conv = FP8BlockDequantizer(weight_block_size=bs, dtype=dt)
        out["fp8block"].append(dict(weight=w, scale_inv=s, block=list(bs), out=conv._create_dequantized_weight(w, s)))
    out["awq"] = [{k: v for k, v in c.items() if v is not None} for c in out["awq"]]
    mg.save("convert.pt", out)


if __name__ == "__main__":
    main()

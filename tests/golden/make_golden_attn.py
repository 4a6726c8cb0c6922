#!/usr/bin/env python
"""
Golden vectors for the ATTN_HEAD strategy (one scale per attention head, shape [heads, 1, 1], broadcast against
[batch, heads, seq, head_dim]; reference quantization/lifecycle/forward.py:229-241 and initialize.py:242-250), produced by importing
the reference exactly like make_golden.py does.

    python tests/golden/make_golden_attn.py   ->  tests/golden/attn.pt.gz
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference from a temp copy)

from compressed_tensors.quantization import QuantizationArgs  # noqa: E402
from compressed_tensors.quantization.lifecycle.forward import dequantize, fake_quantize, quantize  # noqa: E402
from compressed_tensors.quantization.utils import calculate_qparams  # noqa: E402


def main():
    g = torch.Generator().manual_seed(77)
    cases = []
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        for kw in (dict(num_bits=8, type="int", symmetric=True), dict(num_bits=8, type="int", symmetric=False), dict(num_bits=4, type="int", symmetric=True),
                   dict(num_bits=8, type="float", symmetric=True)):
            b, h, s, d = 2, 6, 11, 40
            x = (torch.randn(b, h, s, d, generator=g) * torch.rand(1, h, 1, 1, generator=g) * 3).to(dt)
            args = QuantizationArgs(strategy="attn_head", **kw)
            lo, hi = x.amin((0, 2, 3)).reshape(h, 1, 1), x.amax((0, 2, 3)).reshape(h, 1, 1)      # the observer's reduction for [H, 1, 1] qparams
            scale, zp = calculate_qparams(lo, hi, args)
            qdt = torch.float8_e4m3fn if kw["type"] == "float" else torch.int8
            q = quantize(x, scale, zp, args, dtype=qdt)
            cases.append(dict(args=args.model_dump(mode="json"), x=x, scale=scale, zp=zp, q=q, dq=dequantize(q, scale, zp, args=args),
                              fq=fake_quantize(x, scale, zp, args)))
    mg.save("attn.pt", cases)


if __name__ == "__main__":
    main()

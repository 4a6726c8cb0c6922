#!/usr/bin/env python
"""
Golden vectors for the observer rule `calculate_qparams` (reference
quantization/utils/helpers.py:50-137), produced by importing the reference exactly like
make_golden.py does.  Min/max tensors are given in the weight dtype, as the memoryless min-max
observer and the reference's test fixtures (tests/conftest.py:21-102) produce them.

    python tests/golden/make_golden_qparams.py   ->  tests/golden/qparams.pt.gz
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference from a temp copy)

from compressed_tensors.quantization import QuantizationArgs  # noqa: E402
from compressed_tensors.quantization.utils import calculate_qparams  # noqa: E402


def main():
    g = torch.Generator().manual_seed(2024)
    n = 4096
    cases = []
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        mag = torch.rand(n, generator=g) * 0.2
        mn = (-torch.rand(n, generator=g) * mag).to(dt)
        mx = (torch.rand(n, generator=g) * mag).to(dt)
        # specials: all-zero group, one-sided groups, tiny and huge ranges, equal min/max
        mn[:8] = torch.tensor([0, 0, -1e-3, 0.5, -3.0, -1e-30, -6e4, -0.0]).to(dt)
        mx[:8] = torch.tensor([0, 2.0, 0, 0.75, -1.0, 1e-30, 6e4, 0.0]).to(dt)
        for kw in (dict(num_bits=4, symmetric=True), dict(num_bits=4, symmetric=False), dict(num_bits=8, symmetric=True),
                   dict(num_bits=8, symmetric=False), dict(num_bits=2, symmetric=False), dict(num_bits=8, type="float", symmetric=True)):
            args = QuantizationArgs(strategy="channel", **kw)
            s, z = calculate_qparams(mn.reshape(-1, 1), mx.reshape(-1, 1), args)
            cases.append(dict(args=args.model_dump(mode="json"), min=mn.reshape(-1, 1), max=mx.reshape(-1, 1), scale=s, zp=z))
    mg.save("qparams.pt", cases)


if __name__ == "__main__":
    main()

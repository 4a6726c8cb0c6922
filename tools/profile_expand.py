#!/usr/bin/env python
"""ncu target: the row-expansion kernel (ct_bitmask_decompress with row_offsets) on one [14336, 8192] bf16 tensor at 50 % and 10 % density:
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'expand_rows|move_vec16' -c 4 -o gpurun_out/r2_expand python tools/profile_expand.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compressed_tensors_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
sp = (torch.randn(14336, 8192, device=dev) * 0.02).to(torch.bfloat16)
for density in (0.5, 0.1):
    un = torch.where(torch.rand(sp.shape, device=dev) < density, sp, torch.zeros_like(sp))
    vals, mask, offs = ops.bitmask_compress(un)
    for _ in range(2):
        out = ops.bitmask_decompress(vals, mask, offs, un.shape)
    torch.cuda.synchronize()
    assert torch.equal(out, un)
print("ok")

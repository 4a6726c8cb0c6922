#!/usr/bin/env python
"""ncu target for the per-tensor observer PAIR (reduction -> quantize): does the quantize pass find the tensor in L2?
    ncu --cache-control none --clock-control none --replay-mode application --kernel-name-base demangled \
        -k regex:'minmax_qparams|QuantizeOp' --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
        --csv --log-file gpurun_out/observer_pair.csv python tools/profile_observer.py
--cache-control none: ncu must not flush the L2 between the two kernels; the script evicts it itself (a 512 MB write) before every pair.
A 33.5 MB tensor (fits the 126 MB L2, loaded with evict_last by the reduction) and a 235 MB one (does not fit)."""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compressed_tensors_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
f8 = SimpleNamespace(strategy="tensor", group_size=None, block_structure=None, num_bits=8, type="float", symmetric=True)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for shape in ((4096, 4096), (14336, 8192)):
    x = (torch.randn(shape, device=dev) * 0.02).to(torch.bfloat16)
    for _ in range(3):
        flush.fill_(1)                      # evict the L2 (not matched by the kernel filter)
        torch.cuda.synchronize()
        q, s, z = ops.observe_quantize(x, f8)
        torch.cuda.synchronize()
print("ok")

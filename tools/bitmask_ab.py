"""A/B of the unstructured-bitmask move kernel's segments-per-iteration (CT_B200_BITMASK_SEGMENTS = 1, 2, 4) on one [14336, 8192] bf16
tensor with 50 % zeros: CUDA-event medians of the public ops (count + scan + nnz read-back + move for compress; move for decompress)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressed_tensors_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
sp = (torch.randn(14336, 8192, device=dev) * 0.02).to(torch.bfloat16)
un = torch.where(torch.rand(sp.shape, device=dev) < 0.5, sp, torch.zeros_like(sp))
n = un.numel()
uv, um, uo = ops.bitmask_compress(un)


def med(fn, reps=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]


for seg in ("1", "2", "4"):
    os.environ["CT_B200_BITMASK_SEGMENTS"] = seg
    c = med(lambda: ops.bitmask_compress(un))
    d = med(lambda: ops.bitmask_decompress(uv, um, uo, un.shape))
    print(json.dumps({"segments": int(seg), "compress_us": round(c * 1e3, 1), "compress_GBps": round(n * 3.125 / c / 1e6, 1),
                      "decompress_us": round(d * 1e3, 1), "decompress_GBps": round(n * 3.125 / d / 1e6, 1)}), flush=True)

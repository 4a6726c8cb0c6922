#!/usr/bin/env python
"""Read-only / write-only / copy HBM rates of stock torch kernels on this GPU: the practical ceilings the
streaming kernels are compared with (MEASURED_PEAKS.json holds only the copy number)."""
import json

import torch


def rate(fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return round(nbytes / (best * 1e-3) / 1e9, 1)


def main():
    n = 2 << 30  # 2 Gi bf16 elements = 4 GiB per tensor
    x = torch.ones(n, dtype=torch.bfloat16, device="cuda")
    y = torch.empty_like(x)
    h = torch.empty(n // 2, dtype=torch.bfloat16, device="cuda")
    out = {
        "write_only_fill": rate(lambda: y.fill_(1.0), n * 2),
        "write_only_zero": rate(lambda: y.zero_(), n * 2),
        "read_only_amax": rate(lambda: x.amax(), n * 2),
        "copy": rate(lambda: y.copy_(x), n * 4),
        "read2_write1_add": rate(lambda: torch.add(x[: n // 2], x[n // 2:], out=h), n * 3),
        "read1_write2_cat": rate(lambda: torch.cat([h, h], out=y), n * 3),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()

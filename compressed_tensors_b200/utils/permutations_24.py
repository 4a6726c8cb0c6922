"""
Marlin-24 tile permutations (index arithmetic only; mirror of utils/permutations_24.py:20-53).

Marlin consumes weights in [16*2, 64] tiles laid out for the m16n8k16 tensor-core fragments:
`perm` reorders the 1024 values of a tile (then interleaves them for 4- or 8-bit packing),
`scale_perm` / `scale_perm_single` reorder group / channel scales to match.
"""
from __future__ import annotations

import torch

__all__ = ["get_permutations_24"]


def get_permutations_24(num_bits: int):
    if num_bits == 4:
        interleave = [0, 2, 4, 6, 1, 3, 5, 7]
    elif num_bits == 8:
        interleave = [0, 2, 1, 3]
    else:
        raise ValueError("num_bits must be 4 or 8, got {}".format(num_bits))
    order = []
    for lane in range(32):
        col = lane // 4
        rows = [2 * (lane % 4), 2 * (lane % 4) + 1, 2 * (lane % 4 + 4), 2 * (lane % 4 + 4) + 1]
        base = [16 * row + (col // 2) * 256 + 8 * (col % 2) + 4 * block for block in (0, 1) for row in rows]
        for j in range(4):
            order.extend(p + j for p in base)
    n = len(interleave)
    perm = [order[g * n + i] for g in range(len(order) // n) for i in interleave]
    scale_perm = [8 * i + j for i in range(8) for j in (0, 4, 1, 5, 2, 6, 3, 7)]
    scale_perm_single = [8 * i + j for i in range(8) for j in range(8)]
    return torch.tensor(perm, dtype=torch.int64), scale_perm, scale_perm_single

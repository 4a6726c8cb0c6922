"""
Bitmask sparse compressors named by the north star: `sparse-24-bitmask` and `sparse-bitmask`
(CompressionFormat values config/base.py:17-18 of the reference).

PARITY UNPINNED: the compressor classes themselves were removed from the reference snapshot
(only the config classes, `pack_bitmasks` / `unpack_bitmasks` and `tensor_follows_mask_structure`
remain).  The storage layout follows the format's public description as
restated in oracle/ct_oracle.c; the only normative piece, the mask bit order of pack_bitmasks
(utils/helpers.py:306-343), is pinned by golden vectors.

Local state-dict keys:  weight -> compressed, bitmask, shape (+ row_offsets for sparse-bitmask).
"""
from __future__ import annotations

import torch

from ... import ops
from ...config import CompressionFormat
from ..base import BaseCompressor

__all__ = ["Sparse24BitMaskCompressor", "BitmaskCompressor"]


@BaseCompressor.register(name=CompressionFormat.sparse_24_bitmask.value)
class Sparse24BitMaskCompressor(BaseCompressor):
    """2:4 structured: keep the 2 largest-magnitude values of every 4 (ties: lower column)"""

    @classmethod
    def compression_param_names(cls, scheme=None) -> tuple:
        return ("compressed", "bitmask", "shape")

    @classmethod
    def compress(cls, state_dict, scheme=None):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        if weight.ndim != 2:
            raise ValueError("sparse-24-bitmask compresses 2-D weights")
        if weight.device.type == "meta":
            state_dict["compressed"] = torch.empty((weight.shape[0], weight.shape[1] // 2), dtype=weight.dtype, device="meta")
            state_dict["bitmask"] = torch.empty((weight.shape[0], (weight.shape[1] + 7) // 8), dtype=torch.uint8, device="meta")
        else:
            state_dict["compressed"], state_dict["bitmask"] = ops.sparse24_compress(weight)
        state_dict["shape"] = torch.tensor(weight.shape)
        return state_dict

    @classmethod
    def decompress(cls, state_dict, scheme=None):
        state_dict = state_dict.copy()
        values, bitmask, shape = state_dict.pop("compressed"), state_dict.pop("bitmask"), state_dict.pop("shape")
        shape = tuple(int(v) for v in shape.tolist())
        if values.device.type == "meta":
            state_dict["weight"] = torch.empty(shape, dtype=values.dtype, device="meta")
        else:
            state_dict["weight"] = ops.sparse24_decompress(values, bitmask, shape)
        return state_dict

    @classmethod
    def can_compress(cls, module_type: type, scheme=None) -> bool:
        return False  # never inferred: must be requested explicitly, like the legacy sparse path


@BaseCompressor.register(name=CompressionFormat.sparse_bitmask.value)
class BitmaskCompressor(BaseCompressor):
    """unstructured: non-zero values in row-major order + bitmask + per-row value offsets"""

    @classmethod
    def compression_param_names(cls, scheme=None) -> tuple:
        return ("compressed", "bitmask", "shape", "row_offsets")

    @classmethod
    def compress(cls, state_dict, scheme=None):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        if weight.ndim != 2:
            raise ValueError("sparse-bitmask compresses 2-D weights")
        state_dict["compressed"], state_dict["bitmask"], state_dict["row_offsets"] = ops.bitmask_compress(weight)
        state_dict["shape"] = torch.tensor(weight.shape)
        return state_dict

    @classmethod
    def decompress(cls, state_dict, scheme=None):
        state_dict = state_dict.copy()
        values, bitmask = state_dict.pop("compressed"), state_dict.pop("bitmask")
        offsets, shape = state_dict.pop("row_offsets"), state_dict.pop("shape")
        state_dict["weight"] = ops.bitmask_decompress(values, bitmask, offsets, tuple(int(v) for v in shape.tolist()))
        return state_dict

    @classmethod
    def can_compress(cls, module_type: type, scheme=None) -> bool:
        return False

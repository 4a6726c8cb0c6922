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

__all__ = ["Sparse24BitMaskCompressor", "BitmaskCompressor", "Sparse24PackQuantizedCompressor", "SPARSE24_PACK_QUANTIZED"]


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


# ------------------------------------------------------------------------------------------------------------------------
# BASELINE config 4: "Sparse24BitMask + int4".  In the reference's history a 2:4 model with int4 weights stacked a sparsity
# compressor on a quantization compressor (or used the Marlin-24 kernel layout); neither survives in the snapshot.  This plugin
# is the stack  sparse-24-bitmask . pack-quantized  as ONE compressor with one fused kernel per direction.  The name is this
# engine's own (not a CompressionFormat of the reference); PARITY UNPINNED as a composite, the pieces are pinned
# (tests/test_gpu_sparse24q.py).
# Local keys: weight -> weight_packed (int32 [R, C * bits / 64], kept codes only), bitmask (uint8 [R, C/8]), weight_shape;
#             weight_scale kept; weight_zero_point packed along dim 0 like pack-quantized when asymmetric, dropped when symmetric
# ------------------------------------------------------------------------------------------------------------------------
SPARSE24_PACK_QUANTIZED = "sparse-24-pack-quantized"


@BaseCompressor.register(name=SPARSE24_PACK_QUANTIZED)
class Sparse24PackQuantizedCompressor(BaseCompressor):
    @classmethod
    def compression_param_names(cls, scheme) -> tuple:
        names = ("weight_packed", "bitmask", "weight_scale", "weight_shape")
        if scheme is not None and scheme.weights is not None and not scheme.weights.symmetric:
            names += ("weight_zero_point",)
        return names

    @classmethod
    def compress(cls, state_dict, scheme):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        args = scheme.weights
        zp = state_dict.get("weight_zero_point", None)
        if weight.ndim != 2 or weight.shape[1] % 4 != 0:
            raise ValueError("sparse-24-pack-quantized compresses 2-D weights whose column count is a multiple of 4")
        if weight.device.type == "meta":
            words = -(-(weight.shape[1] // 2) * args.num_bits // 32)
            state_dict["weight_packed"] = torch.empty((weight.shape[0], words), dtype=torch.int32, device="meta")
            state_dict["bitmask"] = torch.empty((weight.shape[0], (weight.shape[1] + 7) // 8), dtype=torch.uint8, device="meta")
        else:
            state_dict["weight_packed"], state_dict["bitmask"] = ops.sparse24_quantize_pack(
                weight, state_dict["weight_scale"], zp if not args.symmetric else None, args)
        state_dict["weight_shape"] = torch.tensor(weight.shape)
        if not args.symmetric and args.strategy in ("group", "channel") and zp is not None:
            state_dict["weight_zero_point"] = ops.pack_to_int32(zp, args.num_bits, packed_dim=0).contiguous()
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme):
        state_dict = state_dict.copy()
        packed, bitmask = state_dict.pop("weight_packed"), state_dict.pop("bitmask")
        scale, args = state_dict["weight_scale"], scheme.weights
        shape = tuple(int(v) for v in state_dict["weight_shape"].tolist())
        zp = state_dict.get("weight_zero_point", None)
        if zp is not None and not args.symmetric and args.strategy in ("group", "channel"):
            zp = ops.unpack_from_int32(zp, args.num_bits, (*shape[:-1], scale.shape[-1]), packed_dim=0)
            state_dict["weight_zero_point"] = zp
        if packed.device.type == "meta":
            state_dict["weight"] = torch.empty(shape, dtype=scale.dtype, device="meta")
        else:
            state_dict["weight"] = ops.sparse24_unpack_dequantize(packed, bitmask, scale, zp, args.num_bits, shape)
        return state_dict

    @classmethod
    def can_compress(cls, module_type: type, scheme=None) -> bool:
        return False   # never inferred and not a CompressionFormat member: fetched explicitly, BaseCompressor.get_value_from_registry(SPARSE24_PACK_QUANTIZED)

"""
pack-quantized: 1..8-bit integer weights stored as a dense little-endian bitstream in int32 words
(mirror of compressors/pack_quantized/base.py:35-177).

compress  : ONE fused kernel bf16/fp16/fp32 weight -> packed words (quantize + pack_to_int32, the
            int8 intermediate never exists in memory); asymmetric group/channel zero points are
            packed along dim 0.
decompress: ONE fused kernel packed words -> dequantized weight in the scale's dtype
            (unpack_from_int32 + dequantize with the strategy inferred from the scale shape).
"""
from __future__ import annotations

import math

import torch

from ... import ops
from ...config import CompressionFormat
from ...quantization import ActivationOrdering, QuantizationScheme, QuantizationStrategy, QuantizationType
from ...utils.helpers import getattr_chain
from ..base import COMPRESSIBLE_MODULE_TYPES, BaseCompressor

__all__ = ["PackedQuantizationCompressor", "PACK_ZP_STRATS"]

PACK_ZP_STRATS = [QuantizationStrategy.GROUP.value, QuantizationStrategy.CHANNEL.value]


@BaseCompressor.register(name=CompressionFormat.pack_quantized.value)
class PackedQuantizationCompressor(BaseCompressor):
    @classmethod
    def compression_param_names(cls, scheme: QuantizationScheme) -> tuple:
        names = ("weight_packed", "weight_scale", "weight_shape")
        if not getattr_chain(scheme, "weights.symmetric", True):
            names += ("weight_zero_point",)
        if getattr_chain(scheme, "weights.actorder", None) == ActivationOrdering.GROUP:
            names += ("weight_g_idx",)
        if getattr_chain(scheme, "input_activations.strategy", None) == QuantizationStrategy.TENSOR_GROUP:
            names += ("input_global_scale",)
        return names

    @classmethod
    def compress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        scale = state_dict.get("weight_scale")
        zero_point = state_dict.get("weight_zero_point", None)
        g_idx = state_dict.get("weight_g_idx", None)
        args = scheme.weights

        if weight.device.type == "meta":
            words = math.ceil(weight.shape[-1] * args.num_bits / 32)
            state_dict["weight_packed"] = torch.empty((*weight.shape[:-1], words), dtype=torch.int32, device="meta")
            state_dict["weight_shape"] = torch.tensor(weight.shape)
            # the reference's shape-only path leaves the zero point unpacked; here it takes the owner's (packed) shape
            # so that module-parallel ranks can receive the owner's tensors by plain broadcast
            if not args.symmetric and args.strategy in PACK_ZP_STRATS and zero_point is not None:
                state_dict["weight_zero_point"] = ops.pack_to_int32(zero_point, args.num_bits, packed_dim=0).contiguous()
            return cls._remove_symmetric_zp(state_dict, scheme)

        state_dict["weight_packed"] = ops.quantize_pack(weight, scale, zero_point, args, g_idx=g_idx)
        state_dict["weight_shape"] = torch.tensor(weight.shape)
        if not args.symmetric and args.strategy in PACK_ZP_STRATS:
            assert zero_point is not None, "Asymmetric quant requires zero-point values"
            state_dict["weight_zero_point"] = ops.pack_to_int32(zero_point, args.num_bits, packed_dim=0).contiguous()
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        packed = state_dict.pop("weight_packed")
        scale = state_dict.get("weight_scale")
        zero_point = state_dict.get("weight_zero_point", None)
        g_idx = state_dict.get("weight_g_idx", None)
        original_shape = state_dict.get("weight_shape")
        args = scheme.weights
        shape = tuple(int(v) for v in original_shape.tolist())

        if not args.symmetric and args.strategy in PACK_ZP_STRATS:
            assert zero_point is not None, "Asymmetric quant requires zero-point values"
            zp_shape = (*shape[:-1], scale.shape[-1])
            zero_point = ops.unpack_from_int32(zero_point, args.num_bits, zp_shape, packed_dim=0)   # shape-only on meta
            state_dict["weight_zero_point"] = zero_point

        if packed.device.type == "meta":
            state_dict["weight"] = torch.empty(shape, dtype=scale.dtype, device="meta")
            return state_dict

        state_dict["weight"] = ops.unpack_dequantize(packed, scale, zero_point, args.num_bits, shape, g_idx=g_idx)
        return state_dict

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        if scheme.input_activations is not None and scheme.input_activations.type == QuantizationType.FLOAT.value:
            return False
        return (module_type in COMPRESSIBLE_MODULE_TYPES and scheme.weights is not None
                and 1 <= scheme.weights.num_bits <= 8 and scheme.weights.type == QuantizationType.INT.value)

"""
nvfp4-pack-quantized (mirror of compressors/nvfp4/base.py:27-139): FP4 E2M1 weights in groups of 16 with float8_e4m3fn
group scales and one float32 global scale per tensor; two values per byte.

compress   = ONE kernel  weight (bf16/fp16/fp32) + scale + global scale -> nibbles   (the reference: quantize -> pack)
decompress = ONE kernel  nibbles + STORED fp8 scale + global scale -> bfloat16       (the reference: unpack -> scale.to ->
             dequantize); like the reference the result is bfloat16 whatever the original weight dtype was
             (unpack_fp4_from_uint8's default dtype, base.py:116-125).
"""
from __future__ import annotations

import torch

from ... import ops
from ...config import CompressionFormat
from ...quantization import QuantizationArgs, QuantizationScheme, QuantizationType
from ...utils.helpers import getattr_chain
from ..base import COMPRESSIBLE_MODULE_TYPES, BaseCompressor

__all__ = ["NVFP4PackedCompressor"]


@BaseCompressor.register(name=CompressionFormat.nvfp4_pack_quantized.value)
class NVFP4PackedCompressor(BaseCompressor):
    _stored_scale = "fp8"   # how weight_scale is held in the compressed state dict

    @classmethod
    def compression_param_names(cls, scheme: QuantizationScheme) -> tuple:
        names = ("weight_packed", "weight_scale", "weight_global_scale")
        if not getattr_chain(scheme, "weights.symmetric", True):
            names += ("weight_zero_point",)
        if not getattr_chain(scheme, "input_activations.dynamic", True):
            names += ("input_global_scale",)
        return names

    @classmethod
    def _compress_scale(cls, scale: torch.Tensor, weights: QuantizationArgs) -> torch.Tensor:
        return scale.to(weights.scale_dtype or torch.float8_e4m3fn)

    @classmethod
    def _decompress_scale(cls, scale: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        return scale.to(dtype)

    @classmethod
    def compress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        scale = state_dict.pop("weight_scale")
        args = scheme.weights
        state_dict["weight_packed"] = ops.quantize_pack_fp4(
            weight, scale, state_dict.get("weight_zero_point", None), args, global_scale=state_dict.get("weight_global_scale", None))
        state_dict["weight_scale"] = cls._compress_scale(scale, args)
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        packed = state_dict.pop("weight_packed")
        scale = state_dict.get("weight_scale")
        dense = torch.bfloat16
        stored = cls._stored_scale if scale.dtype in (torch.float8_e4m3fn, torch.uint8) else None
        state_dict["weight"] = ops.unpack_dequantize_fp4(packed, scale, state_dict.get("weight_global_scale", None), dtype=dense,
                                                         stored_scale=stored)
        state_dict["weight_scale"] = torch.nn.Parameter(cls._decompress_scale(scale, dense), requires_grad=False)
        return state_dict

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        w = scheme.weights
        return (module_type in COMPRESSIBLE_MODULE_TYPES and w is not None and w.num_bits == 4
                and w.type == QuantizationType.FLOAT.value and w.group_size == 16)

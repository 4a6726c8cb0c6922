"""
mxfp8-quantized (mirror of compressors/mxfp8/base.py:28-104): float8_e4m3fn weights in groups of 32 with power-of-two
scales stored as E8M0 exponents.  compress = the naive FP8 group quantize kernel + one E8M0 encode of the scales;
decompress = one E8M0 decode + the FP8 dequantize kernel (bfloat16 out, the decoded scale's dtype).
"""
from __future__ import annotations

import torch

from ...config import CompressionFormat
from ...quantization import QuantizationArgs, QuantizationScheme, QuantizationType
from ..base import COMPRESSIBLE_MODULE_TYPES, BaseCompressor
from ..mx_utils import compress_mx_scale, decompress_mx_scale
from ..naive_quantized.base import NaiveQuantizationCompressor

__all__ = ["MXFP8QuantizationCompressor"]


@BaseCompressor.register(name=CompressionFormat.mxfp8_quantized.value)
class MXFP8QuantizationCompressor(NaiveQuantizationCompressor):
    @classmethod
    def _compress_scale(cls, scale: torch.Tensor, weights: QuantizationArgs) -> torch.Tensor:
        return compress_mx_scale(scale, weights.scale_dtype or torch.uint8)

    @classmethod
    def _decompress_scale(cls, scale: torch.Tensor) -> torch.Tensor:
        return decompress_mx_scale(scale)

    @classmethod
    def compress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = NaiveQuantizationCompressor.compress(state_dict, scheme)
        state_dict["weight_scale"] = cls._compress_scale(state_dict["weight_scale"], scheme.weights)
        return state_dict

    @classmethod
    def decompress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        state_dict["weight_scale"] = cls._decompress_scale(state_dict["weight_scale"])
        return NaiveQuantizationCompressor.decompress(state_dict, scheme)

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        w = scheme.weights
        return (module_type in COMPRESSIBLE_MODULE_TYPES and w is not None and w.num_bits == 8
                and w.type == QuantizationType.FLOAT.value and w.group_size == 32 and w.scale_dtype == torch.uint8)

"""
mxfp4-pack-quantized (mirror of compressors/mxfp4/base.py:27-65): NVFP4's nibble layout with groups of 32, power-of-two
scales stored as E8M0 exponents (uint8) and no global scale.
"""
from __future__ import annotations

import torch

from ...config import CompressionFormat
from ...quantization import QuantizationArgs, QuantizationScheme, QuantizationType
from ...utils.helpers import getattr_chain
from ..base import COMPRESSIBLE_MODULE_TYPES, BaseCompressor
from ..mx_utils import compress_mx_scale, decompress_mx_scale
from ..nvfp4.base import NVFP4PackedCompressor

__all__ = ["MXFP4PackedCompressor"]


@BaseCompressor.register(name=CompressionFormat.mxfp4_pack_quantized.value)
class MXFP4PackedCompressor(NVFP4PackedCompressor):
    _stored_scale = "e8m0"

    @classmethod
    def compression_param_names(cls, scheme: QuantizationScheme) -> tuple:
        names = ("weight_packed", "weight_scale")   # GROUP strategy: no weight_global_scale
        if not getattr_chain(scheme, "weights.symmetric", True):
            names += ("weight_zero_point",)
        if not getattr_chain(scheme, "input_activations.dynamic", True):
            names += ("input_global_scale",)
        return names

    @classmethod
    def _compress_scale(cls, scale: torch.Tensor, weights: QuantizationArgs) -> torch.Tensor:
        return compress_mx_scale(scale, weights.scale_dtype or torch.uint8)

    @classmethod
    def _decompress_scale(cls, scale: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        return decompress_mx_scale(scale).to(dtype)

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        w = scheme.weights
        return (module_type in COMPRESSIBLE_MODULE_TYPES and w is not None and w.num_bits == 4
                and w.type == QuantizationType.FLOAT.value and w.group_size == 32)

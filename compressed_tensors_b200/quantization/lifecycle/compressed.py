"""
compress_quantized_weights: in-place "weight -> quantized dtype" for a frozen module (mirror of
quantization/lifecycle/compressed.py:24-60); the quantize call is the CUDA kernel.
"""
from __future__ import annotations

import torch
from torch.nn import Module

from ..quant_config import QuantizationStatus
from .forward import quantize

__all__ = ["compress_quantized_weights"]


def compress_quantized_weights(module: Module):
    scheme = getattr(module, "quantization_scheme", None)
    if not scheme or not scheme.weights:
        return
    status = getattr(module, "quantization_status", None)
    if status is QuantizationStatus.COMPRESSED:
        return
    weight = getattr(module, "weight", None)
    scale = getattr(module, "weight_scale", None)
    zero_point = getattr(module, "weight_zero_point", None)
    g_idx = getattr(module, "weight_g_idx", None)
    if weight is None or scale is None:
        return
    module.weight.requires_grad = False
    module.weight.data = quantize(x=weight.data, scale=scale, zero_point=zero_point, g_idx=g_idx, args=scheme.weights,
                                  dtype=scheme.weights.pytorch_dtype())
    module.quantization_status = QuantizationStatus.COMPRESSED

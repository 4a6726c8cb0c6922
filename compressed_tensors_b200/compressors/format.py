"""format inference (mirror of compressors/format.py:18-114): first compressor in priority order
whose can_compress() accepts the module type + scheme."""
from __future__ import annotations

from typing import List, Optional

import torch

from ..config import CompressionFormat
from ..quantization import QuantizationScheme
from ..quantization.utils.helpers import is_module_quantized

__all__ = ["infer_model_format", "infer_module_format", "COMPRESSION_FORMAT_PRIORITY"]

COMPRESSION_FORMAT_PRIORITY: List[CompressionFormat] = [
    CompressionFormat.mxfp4_pack_quantized,
    CompressionFormat.mxfp8_quantized,
    CompressionFormat.nvfp4_pack_quantized,
    CompressionFormat.int_quantized,
    CompressionFormat.pack_quantized,
    CompressionFormat.float_quantized,
    CompressionFormat.naive_quantized,
    CompressionFormat.dense,
]


def infer_module_format(module_type: type, scheme: QuantizationScheme) -> CompressionFormat:
    from .base import BaseCompressor

    for fmt in COMPRESSION_FORMAT_PRIORITY:
        if BaseCompressor.get_value_from_registry(fmt.value).can_compress(module_type, scheme):
            return fmt
    raise StopIteration


def infer_model_format(model: torch.nn.Module, force_compression_format: Optional[str] = None) -> CompressionFormat:
    formats = set()
    for _, module in model.named_modules(remove_duplicate=True):
        if not is_module_quantized(module):
            continue
        scheme: QuantizationScheme = module.quantization_scheme
        fmt = infer_module_format(type(module), scheme)
        if force_compression_format is not None:
            fmt = force_compression_format
        elif scheme.format is not None:
            fmt = scheme.format
        scheme.format = CompressionFormat(fmt)
        if scheme.format != CompressionFormat.dense:
            formats.add(scheme.format)
    if not formats:
        return CompressionFormat.dense
    return next(iter(formats)) if len(formats) == 1 else CompressionFormat.mixed_precision

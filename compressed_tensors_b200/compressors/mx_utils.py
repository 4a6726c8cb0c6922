"""E8M0 encode / decode of MX scales (mirror of compressors/mx_utils.py:18-44); each is one CUDA kernel."""
import torch

from .. import ops

__all__ = ["compress_mx_scale", "decompress_mx_scale"]


def compress_mx_scale(scale: torch.Tensor, scale_dtype: torch.dtype) -> torch.Tensor:
    """float scales -> biased power-of-two exponents 127 + floor(log2(scale)) in `scale_dtype` (typically uint8)"""
    return ops.compress_mx_scale(scale, scale_dtype)


def decompress_mx_scale(scale: torch.Tensor) -> torch.Tensor:
    """uint8 E8M0 exponents -> bfloat16 scales 2 ** (e - 127)"""
    return ops.decompress_mx_scale(scale)

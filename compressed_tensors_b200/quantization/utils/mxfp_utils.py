"""
MX (MXFP4 / MXFP8) scale generation on the observer side (mirror of quantization/utils/mxfp_utils.py:38-139).
These run on qparam-sized tensors (one value per 32 weights) with plain torch ops on the tensors' own device;
the weight-sized work of the MX formats (quantize, nibble packing, E8M0 encode of stored scales) is in the CUDA library.
"""
from __future__ import annotations

import math

import torch

from ..quant_args import FP4_E2M1_DATA, FP8_E4M3_DATA, QuantizationArgs, QuantizationType

__all__ = ["maybe_convert_from_mx_exp", "generate_mx_scales", "round_to_power_2", "should_generate_mx_scales"]

# floor(log2(max of the element format)): fp4 e2m1 -> 2, fp8 e4m3 -> 8  (mxfp_utils.py:32-35)
_ELEM_OFFSET = {4: int(math.floor(math.log2(FP4_E2M1_DATA.max))), 8: int(math.floor(math.log2(FP8_E4M3_DATA.max)))}
# (mantissa bits, exponent bits, integer view) per float dtype
_LAYOUT = {torch.bfloat16: (7, 8, torch.uint16), torch.float16: (10, 5, torch.uint16), torch.float32: (23, 8, torch.uint32),
           torch.float64: (52, 11, torch.uint64)}


def should_generate_mx_scales(args: QuantizationArgs) -> bool:
    return (args.num_bits in (4, 8) and args.type == QuantizationType.FLOAT.value and args.group_size == 32
            and args.scale_dtype == torch.uint8)


def maybe_convert_from_mx_exp(args: QuantizationArgs, scale: torch.Tensor) -> torch.Tensor:
    """E8M0 exponents -> float power-of-two scales for MX args, anything else unchanged (mxfp_utils.py:47-67)"""
    if not should_generate_mx_scales(args):
        return scale
    exp = scale.to(torch.int32) - 127
    return (2.00 ** exp.to(torch.float)).to(scale.dtype)


def round_to_power_2(x: torch.Tensor) -> torch.Tensor:
    """keep sign + exponent after adding half of the fp4 mantissa step: the power of two the MX spec rounds a group
    maximum to (mxfp_utils.py:70-121)"""
    if x.dtype not in _LAYOUT:
        raise TypeError(f"Unsupported dtype {x.dtype}")
    mant, expo, view = _LAYOUT[x.dtype]
    wide = torch.int64 if view is torch.uint64 else torch.int32
    bits = x.view(view).to(wide)
    bump = 1 << (mant - FP4_E2M1_DATA.mantissa - 1)
    keep = ((1 << (expo + 1)) - 1) << mant
    snapped = torch.bitwise_and(bits + bump, keep)
    if view is torch.uint16:
        return snapped.to(view).view(x.dtype)
    return snapped.view(x.dtype)


def generate_mx_scales(x: torch.Tensor, num_bits: int = 4) -> torch.Tensor:
    """per-group max |x| -> biased E8M0 exponent (still in x's float dtype; rounded to uint8 by the caller), :124-139"""
    return 127 + torch.floor(torch.log2(round_to_power_2(x))) - _ELEM_OFFSET[num_bits]

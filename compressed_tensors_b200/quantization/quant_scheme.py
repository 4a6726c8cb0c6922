"""
QuantizationScheme + the preset table (mirror of quantization/quant_scheme.py:26-96, :104-160,
:164-439 of the reference).  The preset names and their argument values are the public
vocabulary ("W4A16", "FP8", ...) used by llm-compressor recipes and config.json.
"""
from __future__ import annotations

import warnings
from copy import deepcopy

import torch
from pydantic import BaseModel, ConfigDict, model_validator

from ..config import CompressionFormat
from .quant_args import FP8_E4M3_DATA, DynamicType, QuantizationArgs, QuantizationStrategy, QuantizationType

__all__ = ["QuantizationScheme", "preset_name_to_scheme", "is_preset_scheme", "PRESET_SCHEMES"]


class QuantizationScheme(BaseModel, use_enum_values=True):
    """how the weights / inputs / outputs of the targeted modules are quantized, and stored (format)"""

    targets: list[str]
    weights: QuantizationArgs | None = None
    input_activations: QuantizationArgs | None = None
    output_activations: QuantizationArgs | None = None
    format: CompressionFormat | None = None

    model_config = ConfigDict(extra="forbid")

    @model_validator(mode="after")
    def _check(self):
        ins, outs, w = self.input_activations, self.output_activations, self.weights
        if ins is not None:
            ok = (QuantizationStrategy.TOKEN, QuantizationStrategy.TENSOR, QuantizationStrategy.GROUP,
                  QuantizationStrategy.TENSOR_GROUP, QuantizationStrategy.ATTN_HEAD)
            if ins.strategy not in ok:
                raise NotImplementedError(f"Using {ins.strategy} strategy is not supported for activation quantization")
            if ins.actorder is not None:
                raise ValueError("Cannot apply actorder to input activations")
        if outs is not None and outs.actorder is not None:
            raise ValueError("Cannot apply actorder to output activations")
        if self.format == CompressionFormat.mixed_precision:
            raise ValueError("mixed-precision cannot be set as a format for a QuantizationScheme")
        if (ins and w and w.strategy == QuantizationStrategy.GROUP and ins.strategy == QuantizationStrategy.GROUP
                and w.group_size != ins.group_size):
            warnings.warn(
                "Using GROUP strategy for both weights and input_activations with different group sizes "
                f"({w.group_size} vs {ins.group_size}) may complicate fused kernel implementations. "
                "Consider using TENSOR_GROUP strategy for both or matching group sizes.",
                UserWarning, stacklevel=2,
            )
        return self


def _A(**kw) -> QuantizationArgs:
    return QuantizationArgs(**kw)


def _int_weights(bits: int, act_bits: int = 16) -> dict:
    """integer W{bits}A{act_bits}: group-128 symmetric weights, dynamic per-token int activations below 16 bits"""
    if not 2 <= bits <= 8:
        raise ValueError(f"weight_bits must be 2-8, got {bits}")
    if act_bits not in (4, 8, 16):
        raise ValueError(f"act_bits must be 4, 8, or 16, got {act_bits}")
    if bits > act_bits:
        raise ValueError(f"weight_bits ({bits}) must be <= act_bits ({act_bits})")
    out = dict(weights=_A(num_bits=bits, type="int", strategy="group", group_size=128, symmetric=True, dynamic=False))
    if act_bits < 16:
        out["input_activations"] = _A(num_bits=act_bits, type="int", strategy="token", symmetric=True, dynamic=True)
    return out


def _fp4(group, scale_dtype, strategy, acts=None) -> dict:
    w = _A(num_bits=4, type="float", strategy=strategy, symmetric=True, dynamic=False, group_size=group,
           scale_dtype=scale_dtype, zp_dtype=scale_dtype)
    return dict(weights=w) if acts is None else dict(weights=w, input_activations=acts)


_F8 = FP8_E4M3_DATA.dtype
_U8 = torch.uint8

PRESET_SCHEMES: dict[str, dict] = {
    "UNQUANTIZED": dict(),
    "W8A8": dict(
        weights=_A(num_bits=8, type="int", strategy="channel", symmetric=True, dynamic=False),
        input_activations=_A(num_bits=8, type="int", strategy="token", symmetric=True, dynamic=True),
    ),
    "W4A16_ASYM": dict(weights=_A(num_bits=4, type="int", strategy="group", group_size=128, symmetric=False, dynamic=False)),
    "W4AFP8": dict(
        weights=_A(num_bits=4, type="int", strategy="group", group_size=128, symmetric=True, dynamic=False),
        input_activations=_A(num_bits=8, type="float", strategy="token", symmetric=True, dynamic=True, observer=None),
    ),
    "FP8": dict(
        weights=_A(num_bits=8, type="float", strategy="tensor", symmetric=True, dynamic=False),
        input_activations=_A(num_bits=8, type="float", strategy="tensor", symmetric=True, dynamic=False, observer="static_minmax"),
    ),
    "FP8_DYNAMIC": dict(
        weights=_A(num_bits=8, type="float", strategy="channel", symmetric=True, dynamic=False),
        input_activations=_A(num_bits=8, type="float", strategy="token", symmetric=True, dynamic=True),
    ),
    "FP8_BLOCK": dict(
        weights=_A(num_bits=8, type="float", strategy="block", symmetric=True, dynamic=False, block_structure=[128, 128]),
        input_activations=_A(num_bits=8, type="float", strategy="group", symmetric=True, dynamic=True, group_size=128),
    ),
    "NVFP4A16": _fp4(16, _F8, "tensor_group"),
    "NVFP4": _fp4(16, _F8, "tensor_group", acts=_A(num_bits=4, type="float", strategy="tensor_group", symmetric=True,
                                                    dynamic=DynamicType.LOCAL, group_size=16, observer="static_minmax",
                                                    scale_dtype=_F8, zp_dtype=_F8)),
    "MXFP4A16": _fp4(32, _U8, "group"),
    "MXFP4": _fp4(32, _U8, "group", acts=_A(num_bits=4, type="float", strategy="group", dynamic=True, symmetric=True,
                                             group_size=32, scale_dtype=_U8, zp_dtype=_U8)),
    "MXFP8A16": dict(weights=_A(num_bits=8, type="float", strategy="group", symmetric=True, dynamic=False, group_size=32,
                                scale_dtype=_U8, zp_dtype=_U8)),
    "MXFP8": dict(
        weights=_A(num_bits=8, type="float", strategy="group", symmetric=True, dynamic=False, group_size=32, scale_dtype=_U8, zp_dtype=_U8),
        input_activations=_A(num_bits=8, type="float", strategy="group", dynamic=True, symmetric=True, group_size=32,
                             scale_dtype=_U8, zp_dtype=_U8),
    ),
}
PRESET_SCHEMES["INT8"] = PRESET_SCHEMES["W8A8"]
for _w, _a in ((2, 4), (2, 8), (2, 16), (3, 4), (3, 8), (3, 16), (4, 4), (4, 8), (4, 16), (5, 8), (5, 16),
               (6, 8), (6, 16), (7, 8), (7, 16), (8, 16)):
    PRESET_SCHEMES[f"W{_w}A{_a}"] = _int_weights(_w, _a)


def preset_name_to_scheme(name: str, targets: list[str]) -> QuantizationScheme:
    key = name.upper()
    if key not in PRESET_SCHEMES:
        raise KeyError(f"Unknown preset scheme name {key}, available names: {list(PRESET_SCHEMES.keys())}")
    return QuantizationScheme(targets=targets, **deepcopy(PRESET_SCHEMES[key]))


def is_preset_scheme(name: str) -> bool:
    return name.upper() in PRESET_SCHEMES

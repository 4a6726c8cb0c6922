"""
QuantizationArgs and friends -- the typed description of HOW one tensor is quantized.

Mirror of the reference's schema (quantization/quant_args.py:49-429) for the fields and
validation the compress/decompress path depends on: field names, defaults, the strategy
inference from group_size, the zero-point dtype default, `pytorch_dtype()`, and the JSON
shape produced by model_dump (what lands in config.json).
"""
from __future__ import annotations

import warnings
from enum import Enum
from typing import Any

import torch
from pydantic import BaseModel, ConfigDict, Field, field_serializer, field_validator, model_validator
from pydantic_core import core_schema

__all__ = [
    "FloatArgs",
    "BFLOAT16_DATA",
    "FLOAT16_DATA",
    "FLOAT32_DATA",
    "FLOAT64_DATA",
    "FP8_DTYPE",
    "FP8_E4M3_DATA",
    "FP4_E2M1_DATA",
    "QuantizationType",
    "QuantizationStrategy",
    "QuantizationArgs",
    "ActivationOrdering",
    "DynamicType",
    "TorchDtype",
    "round_to_quantized_type_args",
    "round_to_quantized_type_dtype",
]

FP8_DTYPE = torch.float8_e4m3fn


class FloatArgs:
    """format constants of a floating-point type (quant_args.py:40-46): the E8M0 / MX helpers read exponent and mantissa widths from these"""
    exponent: int
    mantissa: int
    bits = None
    max = None
    min = None
    dtype = None


class BFLOAT16_DATA(FloatArgs):
    exponent, mantissa = 8, 7


class FLOAT16_DATA(FloatArgs):
    exponent, mantissa = 5, 10


class FLOAT32_DATA(FloatArgs):
    exponent, mantissa = 8, 23


class FLOAT64_DATA(FloatArgs):
    exponent, mantissa = 11, 52


class FP8_E4M3_DATA(FloatArgs):
    exponent, mantissa, bits = 4, 3, 8
    max = torch.finfo(torch.float8_e4m3fn).max   # 448
    min = torch.finfo(torch.float8_e4m3fn).min
    dtype = torch.float8_e4m3fn


class FP4_E2M1_DATA(FloatArgs):
    exponent, mantissa, bits = 2, 1, 4
    max, min = 6.0, -6.0
    dtype = None

    @staticmethod
    def cast_to_fp4(x: torch.Tensor) -> torch.Tensor:
        """round to the nearest E2M1 value, same dtype (quant_args.py:55-67 -> fp4_utils.py:77-98); one CUDA kernel"""
        from ..ops import cast_to_fp4

        return cast_to_fp4(x)


class QuantizationType(str, Enum):
    INT = "int"
    FLOAT = "float"


class QuantizationStrategy(str, Enum):
    TENSOR = "tensor"
    CHANNEL = "channel"
    GROUP = "group"
    BLOCK = "block"
    TOKEN = "token"
    TENSOR_GROUP = "tensor_group"
    ATTN_HEAD = "attn_head"


class DynamicType(str, Enum):
    LOCAL = "local"


class ActivationOrdering(str, Enum):
    """GROUP: columns grouped by g_idx (weights stay in their original order); WEIGHT: calibration-only reorder.
    DYNAMIC / STATIC are the reference's aliases (quant_args.py:138-166): distinct members that compare and hash equal to
    GROUP / WEIGHT."""

    GROUP = "group"
    WEIGHT = "weight"
    DYNAMIC = "dynamic"
    STATIC = "static"

    @staticmethod
    def get_aliases() -> dict:
        return {"dynamic": "group", "static": "weight"}

    def _canon(self, v):
        v = v.value if isinstance(v, Enum) else v
        return self.get_aliases().get(v, v)

    def __eq__(self, other):
        return self._canon(self) == self._canon(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._canon(self))


class TorchDtype:
    """pydantic field type for torch.dtype, (de)serialised as 'torch.int8' style strings"""

    @classmethod
    def __get_pydantic_core_schema__(cls, _source, _handler):
        def parse(v):
            if isinstance(v, torch.dtype):
                return v
            if isinstance(v, str):
                dt = getattr(torch, v.replace("torch.", ""), None)
                if isinstance(dt, torch.dtype):
                    return dt
            raise ValueError(f"{v!r} is not a torch dtype")

        return core_schema.no_info_plain_validator_function(
            parse, serialization=core_schema.plain_serializer_function_ser_schema(str)  # in python mode too (utils/type.py:50-54)
        )


class QuantizationArgs(BaseModel, use_enum_values=True):
    """see reference quantization/quant_args.py:169-429 for field semantics"""

    num_bits: int = 8
    type: QuantizationType = QuantizationType.INT
    symmetric: bool = True
    group_size: int | None = None
    strategy: QuantizationStrategy | None = None
    block_structure: list[int] | None = None
    dynamic: DynamicType | bool = False
    actorder: ActivationOrdering | bool | None = None
    scale_dtype: TorchDtype | None = None
    zp_dtype: TorchDtype | None = None
    observer: str | None = None
    observer_kwargs: dict[str, Any] = Field(default_factory=dict)

    model_config = ConfigDict(extra="forbid", arbitrary_types_allowed=True)

    @field_serializer("zp_dtype")
    def _ser_zp(self, dtype):
        return None if self.symmetric or dtype is None else str(dtype)

    @field_serializer("scale_dtype")
    def _ser_scale(self, dtype):
        return None if dtype is None else str(dtype)

    @field_validator("type", mode="before")
    def _v_type(cls, v):
        return QuantizationType(v.lower()) if isinstance(v, str) else v

    @field_validator("strategy", mode="before")
    def _v_strategy(cls, v):
        return QuantizationStrategy(v.lower()) if isinstance(v, str) else v

    @field_validator("dynamic", mode="before")
    def _v_dynamic(cls, v):
        return DynamicType(v.lower()) if isinstance(v, str) else v

    @field_validator("group_size", mode="before")
    def _v_group(cls, v):
        if v is not None and v < -1:
            raise ValueError(
                f"Invalid group size {v}. Use group_size > 0 for strategy='group' and group_size = -1 for 'channel'"
            )
        return v

    @field_validator("block_structure", mode="before")
    def _v_block(cls, v):
        if v is None:
            return v
        err = ValueError(f"Invalid block_structure '{v}'. Must be a list of positive ints [rows, cols].")
        if isinstance(v, str):  # legacy "128x128"
            try:
                v = [int(t) for t in v.split("x")]
            except Exception:
                raise err
        if isinstance(v, (list, tuple)) and len(v) == 2 and all(isinstance(t, int) and t > 0 for t in v):
            return list(v)
        raise err

    @field_validator("actorder", mode="before")
    def _v_actorder(cls, v):
        if isinstance(v, bool):
            return ActivationOrdering.GROUP if v else None
        if isinstance(v, str):
            return ActivationOrdering(v.lower())
        return v

    @model_validator(mode="after")
    def _finish(self):
        strategy, group_size, dynamic, observer, zp_dtype = self.strategy, self.group_size, self.dynamic, self.observer, self.zp_dtype
        grouped = (QuantizationStrategy.GROUP, QuantizationStrategy.TENSOR_GROUP)

        if strategy is None:  # infer from group_size
            if group_size is None:
                strategy = QuantizationStrategy.TENSOR
            elif group_size > 0:
                strategy = QuantizationStrategy.GROUP
            elif group_size == -1:
                strategy = QuantizationStrategy.CHANNEL
            else:
                raise ValueError(
                    f"Invalid group size {group_size}. Use group_size > 0 for strategy='group' and group_size = -1 for 'channel'"
                )
        if strategy == QuantizationStrategy.TOKEN and not dynamic:
            raise ValueError("Cannot perform static token quantization, please use `dynamic=True`")
        if strategy in grouped and (group_size is None or group_size <= 0):
            raise ValueError(f"strategy {strategy} requires group_size to be set to a positive value")
        if group_size is not None and group_size > 0 and strategy not in grouped:
            raise ValueError("group_size requires strategy to be set to 'group'")
        if (strategy == QuantizationStrategy.BLOCK) != (self.block_structure is not None):
            raise ValueError(
                f"Block strategy requires block structure\n{self}" if strategy == QuantizationStrategy.BLOCK
                else f"Block structure requires block strategy\n{self}"
            )
        if self.actorder == ActivationOrdering.GROUP and strategy not in grouped:
            raise ValueError("Must use group or tensor_group quantization strategy in order to apply group activation ordering")

        if dynamic:
            allowed = (QuantizationStrategy.TOKEN, QuantizationStrategy.TENSOR, QuantizationStrategy.TENSOR_GROUP, QuantizationStrategy.GROUP)
            if strategy not in allowed:
                raise ValueError(f"One of {allowed} must be used for dynamic quant.")
            if dynamic == DynamicType.LOCAL and strategy != QuantizationStrategy.TENSOR_GROUP:
                raise ValueError("local is only supported for strategy tensor_group")
            if observer is not None:
                if dynamic is True:
                    if observer != "memoryless":
                        warnings.warn("No observer is used for dynamic quant., setting to None")
                    observer = None
            elif dynamic == DynamicType.LOCAL:
                observer = "minmax"
        elif observer is None:
            observer = "memoryless_minmax"

        if zp_dtype is None:
            if self.num_bits == 4 and self.type == QuantizationType.FLOAT:
                zp_dtype = FP8_E4M3_DATA.dtype
            else:
                zp_dtype = self.pytorch_dtype()

        self.__dict__["strategy"] = QuantizationStrategy(strategy).value
        self.__dict__["observer"] = observer
        self.__dict__["zp_dtype"] = zp_dtype
        return self

    def pytorch_dtype(self) -> torch.dtype:
        """dtype the quantized values are stored in (quant_args.py:413-427)"""
        if self.type == QuantizationType.FLOAT:
            if self.num_bits == 8:
                return FP8_E4M3_DATA.dtype
            raise NotImplementedError("Only num_bits in (8) are supported")
        if self.type == QuantizationType.INT:
            if self.num_bits <= 8:
                return torch.int8
            return torch.int16 if self.num_bits <= 16 else torch.int32
        raise ValueError(f"Invalid quantization type {self.type}")


def round_to_quantized_type_dtype(tensor: torch.Tensor, dtype: torch.dtype, cast_to_original_dtype: bool = True) -> torch.Tensor:
    """host-side qparam helper (tiny tensors: scales / zero points), quant_args.py:432-457"""
    original = tensor.dtype
    if dtype.is_floating_point:
        info = torch.finfo(dtype)
        rounded = torch.clamp(tensor, info.min, info.max).to(dtype)
    else:
        info = torch.iinfo(dtype)
        rounded = torch.round(torch.clamp(tensor, info.min, info.max)).to(dtype)
    return rounded.to(original) if cast_to_original_dtype else rounded


def round_to_quantized_type_args(tensor: torch.Tensor, args: QuantizationArgs, min: torch.Tensor, max: torch.Tensor,
                                 cast_to_original_dtype: bool = True) -> torch.Tensor:
    """clamp + round to the quantized grid (quant_args.py:460-496).  On the hot path this is fused
    into the CUDA kernels; this entry exists for API parity and runs the same kernel: it is
    quantize() with scale 1 and no zero point."""
    from ..ops import quantize as _q

    one = torch.ones((), dtype=tensor.dtype, device=tensor.device)
    out = _q(tensor, one, None, _TensorWide(args), dtype=None)
    if not cast_to_original_dtype and args.type == QuantizationType.FLOAT and args.num_bits == 8:
        return out.to(FP8_E4M3_DATA.dtype)
    return out


class _TensorWide:
    """view of QuantizationArgs that forces the per-tensor strategy"""

    def __init__(self, args):
        self.num_bits, self.type = args.num_bits, args.type
        self.strategy, self.group_size, self.block_structure = "tensor", None, None

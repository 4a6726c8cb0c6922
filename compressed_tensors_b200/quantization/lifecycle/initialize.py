"""
Attach quantization parameters to a module (mirror of quantization/lifecycle/initialize.py:46-276
for Linear / Embedding modules): `weight_scale`, `weight_zero_point`, `weight_g_idx`,
`input_scale`, ... with the shapes each strategy implies and the dtypes the kernels expect
(scale in the weight dtype, zero point in args.zp_dtype).
"""
from __future__ import annotations

import math

import torch
from torch.nn import Module, Parameter

from ..quant_args import ActivationOrdering, DynamicType, QuantizationArgs, QuantizationStrategy
from ..quant_config import QuantizationStatus
from ..quant_scheme import QuantizationScheme
from ..utils.helpers import strategy_cdiv
from .forward import set_forward_quantized

__all__ = ["initialize_module_for_quantization", "initialize_qparams", "QPARAM_SUFFIXES"]

QPARAM_SUFFIXES = ("_scale", "_zero_point", "_global_scale", "_g_idx")
_BASE_NAMES = ("weight", "input", "output", "q", "k", "v")


def _clear_qparams(module: Module):
    for base in _BASE_NAMES:
        for suffix in QPARAM_SUFFIXES:
            name = base + suffix
            if name in module._parameters or name in module._buffers:
                delattr(module, name)


def initialize_module_for_quantization(module: Module, scheme: QuantizationScheme | None = None, force_zero_point: bool = True):
    scheme = scheme or getattr(module, "quantization_scheme", None)
    if scheme is None:
        return
    _clear_qparams(module)
    if not isinstance(module, (torch.nn.Linear, torch.nn.Embedding)):
        raise ValueError(f"Quantization of module type {type(module)} is not supported")
    weight = module.weight
    if scheme.input_activations is not None:
        initialize_qparams(module, "input", scheme.input_activations, weight.shape[-1:], weight.dtype, force_zero_point)
    if scheme.weights is not None:
        initialize_qparams(module, "weight", scheme.weights, weight.shape, weight.dtype, force_zero_point)
    if scheme.output_activations is not None:
        initialize_qparams(module, "output", scheme.output_activations, weight.shape[:-1], weight.dtype, force_zero_point)
    set_forward_quantized(module)
    module.quantization_scheme = scheme
    module.quantization_status = QuantizationStatus.INITIALIZED


def initialize_qparams(module: Module, base_name: str, quantization_args: QuantizationArgs, observed_shape, observed_dtype: torch.dtype,
                       force_zero_point: bool = True):
    strategy, dynamic = quantization_args.strategy, quantization_args.dynamic
    device = module.weight.device if getattr(module, "weight", None) is not None else None
    if dynamic is True:
        return
    if strategy == QuantizationStrategy.TENSOR_GROUP:
        module.register_parameter(f"{base_name}_global_scale", Parameter(torch.empty(1, dtype=torch.float32, device=device), requires_grad=False))
    if dynamic == DynamicType.LOCAL:
        return

    if strategy == QuantizationStrategy.TENSOR:
        shape = (1,)
    elif strategy == QuantizationStrategy.TOKEN:
        raise ValueError("Cannot perform static token quantization")
    elif strategy == QuantizationStrategy.CHANNEL:
        if len(observed_shape) < 2:
            raise ValueError("Channel quant requires at least 2 observed dimensions")
        shape = (observed_shape[-2], 1)
    elif strategy in (QuantizationStrategy.GROUP, QuantizationStrategy.TENSOR_GROUP):
        if len(observed_shape) < 1:
            raise ValueError("Group quant requires at least 1 observed dimension")
        groups = strategy_cdiv(observed_shape[-1], quantization_args.group_size, strategy)
        shape = (*observed_shape[:-1], groups)
        if quantization_args.actorder == ActivationOrdering.GROUP:
            module.register_parameter(f"{base_name}_g_idx", Parameter(
                torch.full((observed_shape[-1],), -1, device=device, dtype=torch.int), requires_grad=False))
    elif strategy == QuantizationStrategy.BLOCK:
        if len(observed_shape) < 2:
            raise ValueError("Block quant requires at least 2 observed dimensions")
        bh, bw = quantization_args.block_structure
        shape = (math.ceil(observed_shape[-2] / bh), strategy_cdiv(observed_shape[-1], bw, strategy))
    elif strategy == QuantizationStrategy.ATTN_HEAD:
        if len(observed_shape) < 3:
            raise ValueError("Attention quant requires at least 3 observed dimensions")
        shape = (observed_shape[-3], 1, 1)
    else:
        raise AssertionError(f"Unknown strategy {strategy}")

    scale_dtype = observed_dtype if observed_dtype in (torch.float16, torch.bfloat16, torch.float32, torch.float64) else torch.float16
    module.register_parameter(f"{base_name}_scale", Parameter(torch.empty(shape, dtype=scale_dtype, device=device), requires_grad=False))
    if force_zero_point or not quantization_args.symmetric:
        module.register_parameter(f"{base_name}_zero_point", Parameter(
            torch.zeros(shape, device=device, dtype=quantization_args.zp_dtype), requires_grad=False))

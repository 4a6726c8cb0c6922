"""
apply_quantization_config: walk a model, match module names / class names against the config's
targets and ignore list, attach schemes and qparams (mirror of quantization/lifecycle/apply.py:100-169
restricted to Linear / Embedding targets -- attention / kv-cache hooks are outside this engine's path).

Target syntax (utils/match.py of the reference): a target matches a module when it equals the
module name, equals the class name, or is "re:<regex>" matching either.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from copy import deepcopy

import torch
from torch.nn import Module

from ..quant_config import QuantizationConfig, QuantizationStatus
from .initialize import initialize_module_for_quantization

__all__ = ["apply_quantization_config", "is_match", "match_named_modules", "load_pretrained_quantization_parameters"]


def _one_match(value: str, target: str) -> bool:
    if target.startswith("re:"):
        return re.match(target[3:], value) is not None
    return value == target


def is_match(name: str, module: Module, targets, ignore=()) -> bool:
    """True when any target (and no ignore entry) matches the module's name or class name"""
    if isinstance(targets, str):
        targets = [targets]
    cls_names = [c.__name__ for c in type(module).__mro__ if c is not object]

    def hit(t: str) -> bool:
        return _one_match(name, t) or any(_one_match(c, t) for c in cls_names)

    if any(hit(t) for t in (ignore or ())):
        return False
    return any(hit(t) for t in targets)


def match_named_modules(model: Module, targets, ignore=()):
    for name, module in model.named_modules():
        if is_match(name, module, targets, ignore):
            yield name, module


def apply_quantization_config(model: Module, config: QuantizationConfig | None, run_compressed: bool = False, show_progress: bool = False):
    config = deepcopy(config)
    if config is None:
        return dict()
    force_zero_point = config.quantization_status < QuantizationStatus.COMPRESSED
    target_to_scheme = OrderedDict()
    for scheme in config.config_groups.values():
        for target in scheme.targets:
            target_to_scheme[target] = scheme
    for name, module in match_named_modules(model, list(target_to_scheme), config.ignore):
        if not isinstance(module, (torch.nn.Linear, torch.nn.Embedding)):
            continue
        first = next(t for t in target_to_scheme if is_match(name, module, [t]))
        module.quantization_scheme = target_to_scheme[first]
        initialize_module_for_quantization(module, force_zero_point=force_zero_point)
        module.quantization_status = config.quantization_status


def load_pretrained_quantization_parameters(model: Module, model_name_or_path: str | None = None, load_weight_qparams: bool = False) -> None:
    """Copy the quantization parameters stored in a saved checkpoint into a model that `apply_quantization_config` has already
    initialised (apply.py:49-97, :195-236): input / output scales, zero points and g_idx always, the weight ones when asked.  A scale
    without a stored zero point (symmetric schemes) gets zeros.  `model_name_or_path` is a local checkpoint directory (or one
    .safetensors file): this engine has no download path."""
    from safetensors import safe_open

    from ...offload import update_offload_parameter
    from ...utils.safetensors_load import get_quantization_parameter_to_path_mapping
    from ..utils.helpers import is_module_quantized

    mapping = get_quantization_parameter_to_path_mapping(str(model_name_or_path))

    def read(full_name: str):
        path = mapping.get(full_name)
        if path is None:
            return None
        with safe_open(path, framework="pt", device="cpu") as f:
            return f.get_tensor(full_name)

    for name, module in model.named_modules():
        if not is_module_quantized(module):
            continue
        scheme = module.quantization_scheme
        bases = [b for b, on in (("input", scheme.input_activations is not None), ("output", scheme.output_activations is not None),
                                 ("weight", load_weight_qparams and scheme.weights is not None)) if on]
        for base in bases:
            g_idx = read(f"{name}.{base}_g_idx")
            if g_idx is not None:
                update_offload_parameter(module, f"{base}_g_idx", g_idx)
            scale = read(f"{name}.{base}_scale")
            if scale is None:
                continue
            update_offload_parameter(module, f"{base}_scale", scale)
            zp = read(f"{name}.{base}_zero_point")
            update_offload_parameter(module, f"{base}_zero_point", zp if zp is not None else torch.zeros_like(scale, device="cpu"))

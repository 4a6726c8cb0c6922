"""
apply_quantization_config: walk a model, match module names / class names against the config's
targets and ignore list, attach schemes and qparams (mirror of quantization/lifecycle/apply.py:100-169
restricted to Linear / Embedding targets -- attention / kv-cache hooks are outside this engine's path).

Target syntax (utils/match.py of the reference): a target matches a module when it equals the
module name, equals the class name, or is "re:<regex>" matching either.
"""
from __future__ import annotations

import logging
from collections import OrderedDict
from copy import deepcopy

import torch
from torch.nn import Module

from ...utils.match import is_match, match_named_modules, match_targets
from ..quant_config import QuantizationConfig, QuantizationStatus
from .initialize import initialize_module_for_quantization

__all__ = ["apply_quantization_config", "is_match", "match_named_modules", "load_pretrained_quantization_parameters"]

_LOGGER = logging.getLogger(__name__)


def _looks_like_attention(module: Module) -> bool:
    """the reference's `_is_attention_module` test (quantization/lifecycle/initialize.py:123-130): class name contains "attention"
    and the module owns k_proj / v_proj / qkv_proj"""
    return "attention" in module.__class__.__name__.lower() and any(hasattr(module, n) for n in ("k_proj", "v_proj", "qkv_proj"))


def apply_quantization_config(model: Module, config: QuantizationConfig | None, run_compressed: bool = False, show_progress: bool = False):
    """quantization/lifecycle/apply.py:100-169.  A module that matches several targets takes the scheme of the most specific one
    (`match_targets`: exact name, then regex on the name, then class name -- apply.py:149-151, :258-265), not the first in config
    order.  KV-cache schemes and attention-module targets are outside this engine's path (OUT OF SCOPE): they raise /
    warn instead of being dropped silently."""
    config = deepcopy(config)
    if config is None:
        return dict()
    force_zero_point = config.quantization_status < QuantizationStatus.COMPRESSED
    if getattr(config, "kv_cache_scheme", None) is not None:
        raise NotImplementedError(
            "compressed_tensors_b200 does not implement KV-cache quantization (the reference's _apply_kv_cache_scheme / "
            "initialize_hooked_kv_cache, apply.py:172-192): a checkpoint with `kv_cache_scheme` cannot be loaded by this engine")
    target_to_scheme = OrderedDict()
    for scheme in config.config_groups.values():
        for target in scheme.targets:
            target_to_scheme[target] = scheme
    for name, module in match_named_modules(model, target_to_scheme, config.ignore, warn_on_fail=True):
        scheme = target_to_scheme[match_targets(name, module, target_to_scheme)[0]]
        if isinstance(module, (torch.nn.Linear, torch.nn.Embedding)):
            module.quantization_scheme = scheme
            initialize_module_for_quantization(module, force_zero_point=force_zero_point)
            module.quantization_status = config.quantization_status
        elif _looks_like_attention(module):
            _LOGGER.warning(f"{name}: attention-module quantization targets are not supported by compressed_tensors_b200; "
                            "no q/k/v scales are attached (the reference would run initialize_hooked_attention here)")


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

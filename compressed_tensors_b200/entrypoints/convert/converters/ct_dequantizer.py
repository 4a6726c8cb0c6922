"""
CompressedTensorsDequantizer (mirror of entrypoints/convert/converters/ct_dequantizer.py:21-171): turn a compressed-tensors
checkpoint back into dense weights.  Each matched module is one fused unpack+dequantize (or dequantize) kernel on the B200;
tensors that arrive on the CPU are staged through the GPU by the ops layer.
"""
from __future__ import annotations

import os
from typing import Iterable

import pydantic
import torch

from ....compressors.base import BaseCompressor
from ....compressors.format import infer_module_format
from ....config import CompressionFormat
from ....quantization import QuantizationConfig
from ....utils.match import match_name, match_quantizable_tensors
from ....utils.safetensors_load import CONFIG_NAME, get_checkpoint_files, get_quantization_config

__all__ = ["CompressedTensorsDequantizer"]

_KV_CACHE_PARAMS = ("k_scale", "v_scale", "q_scale")   # quantization/quant_args KVCacheScaleType values


class CompressedTensorsDequantizer:
    def __init__(self, model_stub: str | os.PathLike, ignore: Iterable[str] = tuple(), dtype=torch.bfloat16):
        self.dtype = dtype
        files = get_checkpoint_files(model_stub)
        cfg = files.get(CONFIG_NAME) or files.get("params.json")
        if cfg is None:
            raise ValueError("Could not find config.json file")
        data = get_quantization_config(cfg)
        if data is None:
            raise ValueError("Could not find quantization_config in config.json")
        try:
            self.quant_config = QuantizationConfig.model_validate(data)
        except pydantic.ValidationError as e:
            raise ValueError("Model quantization config was found, but it does not match expected compressed-tensors quantization format") from e
        self.quant_config.ignore = list(self.quant_config.ignore or []) + list(ignore)
        for scheme in self.quant_config.config_groups.values():
            scheme.format = CompressionFormat(infer_module_format(torch.nn.Linear, scheme))

    def _schemes(self):
        for scheme in self.quant_config.config_groups.values():
            comp = BaseCompressor.get_value_from_registry(scheme.format.value if hasattr(scheme.format, "value") else scheme.format)
            yield scheme, comp, comp.compression_param_names(scheme)

    def process(self, tensors: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        out = {}
        for scheme, comp, names in self._schemes():
            for module_name, _ in match_quantizable_tensors(tensors, ignore=self.quant_config.ignore, targets=scheme.targets, param_targets=[names[0]]):
                state = {n: tensors.pop(f"{module_name}.{n}") for n in names}
                out[f"{module_name}.weight"] = comp.decompress(state, scheme)["weight"].to(self.dtype)
        for name, t in tensors.items():   # everything untargeted is copied, kv-cache qparams are dropped
            if not name.endswith(_KV_CACHE_PARAMS):
                out[name] = t
        return out

    def validate(self, tensors: dict[str, torch.Tensor]):
        consumed, matched = set(), set()
        for scheme, _, names in self._schemes():
            for module_name, _ in match_quantizable_tensors(tensors, self.quant_config.ignore, scheme.targets, param_targets=[names[0]]):
                matched.add(module_name)
                for n in names:
                    key = f"{module_name}.{n}"
                    if key not in tensors:
                        raise ValueError(f"Expected key {key} not found")
                    consumed.add(key)
        left = [n for n in tensors if n not in consumed and n.rpartition(".")[0] in matched]
        if left:
            raise ValueError(f"Found {len(left)} unconsumed keys -- {left}")

    def create_config(self):
        return None

    def get_dependencies(self, weight_name: str) -> set[str]:
        module_name, _, param_name = weight_name.rpartition(".")
        if any(match_name(module_name, i) for i in self.quant_config.ignore):
            return set()
        for scheme, _, names in self._schemes():
            if "Linear" in scheme.targets or any(match_name(module_name, t) for t in scheme.targets):
                return {f"{module_name}.{n}" for n in names[1:]} if param_name == names[0] else set()
        return set()

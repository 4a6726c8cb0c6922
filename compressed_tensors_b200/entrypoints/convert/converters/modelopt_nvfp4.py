"""
ModelOptNvfp4Converter (mirror of entrypoints/convert/converters/modelopt_nvfp4.py:20-155): NVIDIA ModelOpt NVFP4 checkpoints
-> compressed-tensors nvfp4-pack-quantized.  Pure renaming plus reciprocals of the per-tensor scales (qparam-sized); the
packed weights are byte-identical in both conventions, so no weight-sized kernel runs here.
"""
from __future__ import annotations

from typing import Iterable

import torch

from ....config import CompressionFormat
from ....quantization import QuantizationArgs, QuantizationConfig, QuantizationScheme, QuantizationStatus, preset_name_to_scheme
from ....utils.match import match_name, match_quantizable_tensors

__all__ = ["ModelOptNvfp4Converter"]


class ModelOptNvfp4Converter:
    def __init__(self, ignore: Iterable[str] = tuple(), targets: Iterable[str] = tuple(), kv_cache_scheme: QuantizationArgs | None = None):
        self.ignore, self.targets, self.kv_cache_scheme = list(ignore), list(targets), kv_cache_scheme
        self.param_names = ["input_scale", "weight", "weight_scale", "weight_scale_2"]
        if kv_cache_scheme is not None:
            self.param_names += ["k_scale", "v_scale"]

    def process(self, tensors: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        for module_name, name in match_quantizable_tensors(tensors, self.ignore, self.targets, param_targets=self.param_names):
            param = name.rpartition(".")[-1]
            if param == "input_scale":                     # modelopt stores 1 / global scale
                tensors[f"{module_name}.input_global_scale"] = 1 / tensors.pop(name)
            elif param == "weight":                        # uint8 nibbles, same layout
                tensors[f"{module_name}.weight_packed"] = tensors.pop(name)
            elif param == "weight_scale_2":
                tensors[f"{module_name}.weight_global_scale"] = 1 / tensors.pop(name)
            elif param in ("k_scale", "v_scale"):
                tensors[name] = tensors[name].to(self.kv_cache_scheme.scale_dtype or torch.bfloat16)
        return tensors

    def validate(self, tensors: dict[str, torch.Tensor]):
        targeted = {n for _, n in match_quantizable_tensors(tensors, self.ignore, self.targets, param_targets=self.param_names)}
        banned = ("input_scale", "weight_scale", "weight_scale_2", "k_scale", "v_scale")
        for name in tensors:
            if name not in targeted and not any(match_name(name, i) for i in self.ignore) and name.rpartition(".")[-1] in banned:
                raise ValueError(f"Hit unexpected non-targeted tensor {name}")

    def get_dependencies(self, weight_name: str) -> set[str]:
        module_name, _, param = weight_name.rpartition(".")
        if (param == "weight" and any(match_name(module_name, t) for t in self.targets)
                and not any(match_name(module_name, i) for i in self.ignore)):
            deps = {f"{module_name}.input_scale", f"{module_name}.weight_scale", f"{module_name}.weight_scale_2"}
            if self.kv_cache_scheme:
                if module_name.endswith("k_proj"):
                    deps.add(f"{module_name}.k_scale")
                if module_name.endswith("v_proj"):
                    deps.add(f"{module_name}.v_scale")
            return deps
        return set()

    def create_config(self) -> QuantizationConfig:
        base = preset_name_to_scheme("NVFP4", self.targets)
        scheme = QuantizationScheme(targets=self.targets, weights=base.weights, input_activations=base.input_activations,
                                    format=CompressionFormat.nvfp4_pack_quantized.value)
        return QuantizationConfig(config_groups={"config_group_0": scheme}, ignore=self.ignore, kv_cache_scheme=self.kv_cache_scheme,
                                  format=CompressionFormat.nvfp4_pack_quantized.value, quantization_status=QuantizationStatus.COMPRESSED.value)

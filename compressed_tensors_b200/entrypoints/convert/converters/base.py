"""
Converter protocol and the shard planner (mirror of entrypoints/convert/converters/base.py:19-144).
A converter rewrites the tensors of one safetensors shard; `get_dependencies` names the partner tensors a weight needs
(scales, zero points, ...) so that every job can load them even when they live in another shard.
"""
from __future__ import annotations

from collections import defaultdict
from typing import TYPE_CHECKING, Protocol

import torch

__all__ = ["Converter", "build_inverse_weight_maps"]

if TYPE_CHECKING:
    from ....quantization import QuantizationConfig


class Converter(Protocol):
    def process(self, tensors: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        """name -> tensor of one shard in, name -> tensor to be written out"""
        raise NotImplementedError()

    def validate(self, tensors: dict[str, torch.Tensor]):
        """raise early if the shard does not look like what the converter expects"""
        raise NotImplementedError()

    def create_config(self) -> "QuantizationConfig | None":
        """quantization_config of the converted checkpoint (None: it becomes full precision)"""
        raise NotImplementedError()

    def get_dependencies(self, weight_name: str) -> set[str]:
        raise NotImplementedError()


def build_inverse_weight_maps(weight_map: dict[str, str], model_files: dict[str, str], converters: list) -> dict[str, dict[str, list[str]]]:
    """
    For every output shard: {resolved source path: [tensor names]} = its own primary tensors plus the transitive
    dependencies the converters declare, wherever those are stored.  A tensor that is somebody's dependency is never a
    primary of its own shard (it is written with the tensor that needs it).
    """
    def closure(name: str, seen: set[str]) -> set[str]:
        for conv in converters:
            for dep in conv.get_dependencies(name):
                if dep not in seen:
                    seen.add(dep)
                    closure(dep, seen)
        return seen

    deps = {}
    for name in weight_map:
        deps[name] = closure(name, set())
        assert name not in deps[name], f"{name} found in dependencies {deps[name]}"
    partners = set().union(*deps.values()) if deps else set()

    plans: dict[str, dict[str, list[str]]] = defaultdict(lambda: defaultdict(list))
    for name, shard in weight_map.items():
        if name in partners:
            continue
        for needed in (name, *deps[name]):
            if needed not in weight_map:
                raise ValueError(f"Dependency weight {needed} not found in weight map")
            plans[shard][model_files[weight_map[needed]]].append(needed)
    return {shard: dict(v) for shard, v in plans.items()}

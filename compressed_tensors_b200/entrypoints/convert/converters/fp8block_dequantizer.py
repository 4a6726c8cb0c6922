"""
FP8BlockDequantizer (mirror of entrypoints/convert/converters/fp8block_dequantizer.py:14-158): checkpoints whose Linear weights
are float8_e4m3fn with one float32 `weight_scale_inv` per (128 x 128) block -> dense weights.
The reference pads to whole blocks, reshapes to 4-D, multiplies in float32 and casts; here it is ONE streaming kernel
(DequantF32ScaleOp: fp8 -> float32, x scale, round to the output dtype) addressing the ceil-div scale grid directly.
"""
from __future__ import annotations

from typing import Iterable

import torch

from .... import ops
from ....utils.match import match_name, match_quantizable_tensors

__all__ = ["FP8BlockDequantizer"]


class FP8BlockDequantizer:
    def __init__(self, ignore: Iterable[str] = tuple(), targets: Iterable[str] = tuple(), weight_block_size=(128, 128), dtype=torch.bfloat16):
        self.ignore, self.targets = list(ignore), list(targets)
        self.weight_block_size = tuple(weight_block_size)
        self.dtype = dtype
        self.param_names = ["weight", "weight_scale_inv"]

    def process(self, tensors: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        for module_name, name in match_quantizable_tensors(tensors, self.ignore, self.targets, param_targets=self.param_names):
            if name.rpartition(".")[-1] == "weight":
                tensors[f"{module_name}.weight"] = self._create_dequantized_weight(tensors[f"{module_name}.weight"], tensors[f"{module_name}.weight_scale_inv"])
                del tensors[f"{module_name}.weight_scale_inv"]
        return tensors

    def validate(self, tensors: dict[str, torch.Tensor]):
        targeted = [n for _, n in match_quantizable_tensors(tensors, self.ignore, self.targets, param_targets=self.param_names)]
        for name in targeted:
            module_name, _, param = name.rpartition(".")
            if param == "weight" and f"{module_name}.weight_scale_inv" not in tensors:
                raise ValueError(f"Found weight without corresponding weight_scale_inv {name}")
            if param == "weight_scale_inv" and f"{module_name}.weight" not in tensors:
                raise ValueError(f"Found weight_scale_inv without corresponding weight {name}")
        for name in tensors:
            if name not in targeted and not any(match_name(name, i) for i in self.ignore) and name.rsplit(".", 1)[-1] == "weight_scale_inv":
                raise ValueError(f"Found unexpected non-targeted tensor {name}")

    def create_config(self):
        return None

    def get_dependencies(self, weight_name: str) -> set[str]:
        module_name, _, param = weight_name.rpartition(".")
        if (param == "weight" and any(match_name(module_name, t) for t in self.targets)
                and not any(match_name(module_name, i) for i in self.ignore)):
            return {f"{module_name}.weight_scale_inv"}
        return set()

    def _create_dequantized_weight(self, weight: torch.Tensor, weight_scale_inv: torch.Tensor) -> torch.Tensor:
        """(weight.to(float32) * scale_inv.to(float32) per block).to(self.dtype), shape of `weight`"""
        return ops.dequantize_block_fp8(weight, weight_scale_inv, self.weight_block_size, self.dtype)

"""
AutoAWQConverter (mirror of entrypoints/convert/converters/autoawq.py:27-262): AutoAWQ GEMM checkpoints
(qweight / qzeros / scales) -> compressed-tensors pack-quantized W4A16 (weight_packed / weight_zero_point / weight_scale /
weight_shape).  The reference unpacks to int8, gathers, masks, transposes and re-packs; that chain is a fixed permutation
of nibbles, done here by one kernel per tensor (`ct_awq_repack_int4`, csrc/convert.cu) without ever widening the codes.
"""
from __future__ import annotations

import re
from typing import Any, Iterable, cast

import torch

from .... import ops
from ....config import CompressionFormat
from ....quantization import QuantizationArgs, QuantizationConfig, QuantizationScheme, QuantizationStatus, QuantizationStrategy, QuantizationType
from ....utils.match import match_name

__all__ = ["AutoAWQConverter"]


class AutoAWQConverter:
    AWQ_REVERSE_ORDER = [0, 4, 1, 5, 2, 6, 3, 7]

    def __init__(self, bits: int = 4, group_size: int = 128, zero_point: bool = True, version: str = "gemm",
                 ignore: Iterable[str] = ("lm_head",), targets: Iterable[str] = ("Linear",)):
        if bits != 4:
            raise ValueError("AutoAWQConverter currently supports only 4-bit weights")
        if version != "gemm":
            raise ValueError(f"Unsupported AutoAWQ version: {version}")
        self.bits, self.group_size, self.zero_point, self.version = bits, group_size, zero_point, version
        self.ignore, self.targets = list(ignore), list(targets)

    @classmethod
    def from_pretrained(cls, model_name_or_path: str, targets: Iterable[str] = ("Linear",), trust_remote_code: bool = False) -> "AutoAWQConverter":
        from transformers import AutoConfig

        config = AutoConfig.from_pretrained(model_name_or_path, trust_remote_code=trust_remote_code)
        awq = getattr(config, "quantization_config", None)
        if awq is None:
            raise ValueError("Model config does not contain quantization_config")
        awq = cast(dict[str, Any], awq)
        if awq.get("quant_method") != "awq":
            raise ValueError("Model config is not an AutoAWQ config")
        return cls.from_autoawq_config(awq, targets=targets)

    @classmethod
    def from_autoawq_config(cls, autoawq_config: dict[str, Any], targets: Iterable[str] = ("Linear",)) -> "AutoAWQConverter":
        ignore = ["lm_head"] + [f"re:.*{re.escape(m)}.*" for m in (autoawq_config.get("modules_to_not_convert") or [])]
        return cls(bits=autoawq_config.get("bits", 4), group_size=autoawq_config.get("group_size", 128),
                   zero_point=autoawq_config.get("zero_point", True), version=autoawq_config.get("version", "gemm"), ignore=ignore, targets=targets)

    def process(self, tensors: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        for name in list(tensors):
            if not name.endswith(".qweight"):
                continue
            module_name = name.removesuffix(".qweight")
            if not self._is_targeted(module_name):
                continue
            qweight = tensors.pop(f"{module_name}.qweight")
            qzeros = tensors.pop(f"{module_name}.qzeros", None)
            scales = tensors.pop(f"{module_name}.scales")
            if self.zero_point and qzeros is None:
                raise ValueError("Found qweight without corresponding qzeros")
            in_features, out_features = qweight.shape[0], qweight.shape[1] * (32 // self.bits)
            tensors[f"{module_name}.weight_scale"] = scales.T.contiguous()
            tensors[f"{module_name}.weight_packed"] = ops.awq_repack(qweight)
            tensors[f"{module_name}.weight_shape"] = torch.tensor([out_features, in_features])
            if self.zero_point:
                tensors[f"{module_name}.weight_zero_point"] = ops.awq_repack_zeros(qzeros)
        return tensors

    def validate(self, tensors: dict[str, torch.Tensor]):
        for name in tensors:
            module_name, _, param = name.rpartition(".")
            if param in {"qweight", "qzeros", "scales"} and not self._is_targeted(module_name):
                raise ValueError(f"Found unexpected non-targeted tensor {name}")
            if param != "qweight" or not self._is_targeted(module_name):
                continue
            for dep in self.get_dependencies(name):
                if dep not in tensors:
                    raise ValueError(f"Found qweight without corresponding {dep}")

    def create_config(self) -> QuantizationConfig:
        weights = QuantizationArgs(num_bits=self.bits, type=QuantizationType.INT, symmetric=not self.zero_point, group_size=self.group_size,
                                   strategy=QuantizationStrategy.GROUP)
        return QuantizationConfig(
            config_groups={"config_group_0": QuantizationScheme(targets=self.targets, weights=weights, format=CompressionFormat.pack_quantized.value)},
            ignore=self.ignore, format=CompressionFormat.pack_quantized.value, quantization_status=QuantizationStatus.COMPRESSED.value)

    def get_dependencies(self, weight_name: str) -> set[str]:
        module_name, _, suffix = weight_name.rpartition(".")
        if suffix == "qweight" and self._is_targeted(module_name):
            deps = {f"{module_name}.scales"}
            if self.zero_point:
                deps.add(f"{module_name}.qzeros")
            return deps
        return set()

    def _is_targeted(self, module_name: str) -> bool:
        if any(match_name(module_name, i) for i in self.ignore):
            return False
        if len(self.targets) == 0 or "Linear" in self.targets:
            return True
        return any(match_name(module_name, t) for t in self.targets)

    # kept for API parity with the reference (its tests call them); the converter itself uses the fused kernel
    @staticmethod
    def unpack_awq(qweight: torch.Tensor, qzeros: torch.Tensor | None, bits: int):
        shifts = torch.arange(0, 32, bits, device=qweight.device)
        iw = torch.bitwise_right_shift(qweight[:, :, None], shifts[None, None, :]).to(torch.int8)
        iw = iw.view(iw.shape[0], -1)
        if qzeros is None:
            return iw, None
        iz = torch.bitwise_right_shift(qzeros[:, :, None], shifts[None, None, :]).to(torch.int8)
        return iw, iz.view(iz.shape[0], -1)

    @staticmethod
    def reverse_awq_order(iweights: torch.Tensor, izeros: torch.Tensor | None, bits: int):
        order = torch.arange(iweights.shape[-1], dtype=torch.int32, device=iweights.device).view(-1, 32 // bits)
        order = order[:, AutoAWQConverter.AWQ_REVERSE_ORDER].view(-1)
        return iweights[:, order], (izeros[:, order] if izeros is not None else None)

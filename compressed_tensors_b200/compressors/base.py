"""
The compressor plugin contract -- mirror of compressors/base.py:34-219 of the reference.

A compressor is a class registered under a CompressionFormat name; it is never instantiated.
All entry points are classmethods over *local-name* state dicts ("weight", "weight_scale", ...):

    compress(state_dict, scheme) -> state_dict      (does not mutate its input)
    decompress(state_dict, scheme) -> state_dict
    can_compress(module_type, scheme) -> bool
    compression_param_names(scheme) -> tuple        (first element = the root parameter)

`compress_module` / `decompress_module` apply a compressor to one nn.Module in place.
"""
from __future__ import annotations

from abc import ABC
from typing import Optional

import torch

from ..config import CompressionFormat
from ..quantization import QuantizationScheme, QuantizationStatus
from ..registry import RegistryMixin
from ..utils.module import get_direct_state_dict, replace_direct_state_dict

__all__ = ["BaseCompressor", "compress_module", "decompress_module", "COMPRESSIBLE_MODULE_TYPES"]

COMPRESSIBLE_MODULE_TYPES = (torch.nn.Linear, torch.nn.Embedding)
TensorStateDict = dict


class BaseCompressor(RegistryMixin, ABC):
    @classmethod
    def compression_param_names(cls, scheme: QuantizationScheme) -> tuple:
        raise NotImplementedError(f"{cls.__name__} does not implement the classmethod compression_param_names interface")

    @classmethod
    def compress(cls, state_dict: TensorStateDict, scheme: QuantizationScheme) -> TensorStateDict:
        raise NotImplementedError(f"{cls.__name__} does not implement the classmethod compress interface")

    @classmethod
    def decompress(cls, state_dict: TensorStateDict, scheme: QuantizationScheme) -> TensorStateDict:
        raise NotImplementedError(f"{cls.__name__} does not implement the classmethod decompress interface")

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        raise NotImplementedError(f"{cls.__name__} does not implement match")

    @classmethod
    def compress_module(cls, module: torch.nn.Module) -> None:
        scheme = getattr(module, "quantization_scheme")
        replace_direct_state_dict(module, cls.compress(get_direct_state_dict(module), scheme))
        module.quantization_status = QuantizationStatus.COMPRESSED

    @classmethod
    def decompress_module(cls, module: torch.nn.Module) -> None:
        scheme = getattr(module, "quantization_scheme")
        replace_direct_state_dict(module, cls.decompress(get_direct_state_dict(module), scheme))
        module.quantization_status = QuantizationStatus.DECOMPRESSED

    @classmethod
    def _remove_symmetric_zp(cls, state_dict: TensorStateDict, scheme: QuantizationScheme) -> TensorStateDict:
        """zero points of symmetric schemes are not stored (vLLM refuses them), base.py:148-167"""
        for args, key in ((scheme.input_activations, "input_zero_point"), (scheme.weights, "weight_zero_point"),
                          (scheme.output_activations, "output_zero_point")):
            if args and args.symmetric:
                state_dict.pop(key, None)
        return state_dict


def _resolve_format(module: torch.nn.Module, scheme: QuantizationScheme, format) -> CompressionFormat:
    from .format import infer_module_format

    # precedence: explicit argument > scheme.format > inferred (base.py:189-191)
    return CompressionFormat(format or scheme.format or infer_module_format(type(module), scheme))


def compress_module(module: torch.nn.Module, format: Optional[CompressionFormat] = None):
    scheme = getattr(module, "quantization_scheme", None)
    if not isinstance(scheme, QuantizationScheme):
        return
    scheme.format = _resolve_format(module, scheme, format)
    BaseCompressor.get_value_from_registry(scheme.format.value).compress_module(module)


def decompress_module(module: torch.nn.Module, format: Optional[CompressionFormat] = None):
    scheme = getattr(module, "quantization_scheme", None)
    if not isinstance(scheme, QuantizationScheme):
        return
    scheme.format = _resolve_format(module, scheme, format)
    BaseCompressor.get_value_from_registry(scheme.format.value).decompress_module(module)

"""
naive-quantized / int-quantized / float-quantized: the weight is replaced by its int8 or
float8_e4m3fn codes of the same shape (mirror of compressors/naive_quantized/base.py:27-164).
compress = one quantize kernel (bf16 -> 1-byte codes), decompress = one dequantize kernel.
"""
from __future__ import annotations

from ... import ops
from ...config import CompressionFormat
from ...quantization import ActivationOrdering, QuantizationScheme, QuantizationStrategy, QuantizationType
from ...utils.helpers import getattr_chain
from ..base import COMPRESSIBLE_MODULE_TYPES, BaseCompressor

__all__ = ["NaiveQuantizationCompressor", "IntQuantizationCompressor", "FloatQuantizationCompressor"]


@BaseCompressor.register(name=CompressionFormat.naive_quantized.value)
class NaiveQuantizationCompressor(BaseCompressor):
    @classmethod
    def compression_param_names(cls, scheme: QuantizationScheme) -> tuple:
        names = ("weight", "weight_scale")
        if not getattr_chain(scheme, "weights.symmetric", True):
            names += ("weight_zero_point",)
        if getattr_chain(scheme, "weights.actorder", None) == ActivationOrdering.GROUP:
            names += ("weight_g_idx",)
        return names

    @classmethod
    def compress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        args = scheme.weights
        # block strategy: the reference zero-pads to whole blocks, quantizes, slices back
        # (naive_quantized/base.py:70-94); the kernel addresses the ceil-div scale grid directly,
        # padded elements never reach the output, so no padded copy is made.
        state_dict["weight"] = ops.quantize(
            weight, state_dict.get("weight_scale"), state_dict.get("weight_zero_point", None), args,
            dtype=args.pytorch_dtype(), g_idx=state_dict.get("weight_g_idx", None),
        )
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme: QuantizationScheme):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        state_dict["weight"] = ops.dequantize(
            weight, state_dict.get("weight_scale"), state_dict.get("weight_zero_point", None),
            g_idx=state_dict.get("weight_g_idx", None),
        )
        return state_dict

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        return module_type in COMPRESSIBLE_MODULE_TYPES and scheme.weights is not None


@BaseCompressor.register(name=CompressionFormat.int_quantized.value)
class IntQuantizationCompressor(NaiveQuantizationCompressor):
    """W8A8-style integer weights with quantized input activations"""

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        return (module_type in COMPRESSIBLE_MODULE_TYPES and scheme.input_activations is not None
                and scheme.weights is not None and scheme.weights.type == QuantizationType.INT.value)


@BaseCompressor.register(name=CompressionFormat.float_quantized.value)
class FloatQuantizationCompressor(NaiveQuantizationCompressor):
    """FP8 weights with quantized input activations"""

    @classmethod
    def can_compress(cls, module_type: type, scheme: QuantizationScheme) -> bool:
        return (module_type in COMPRESSIBLE_MODULE_TYPES and scheme.input_activations is not None
                and scheme.weights is not None and scheme.weights.type == QuantizationType.FLOAT.value)

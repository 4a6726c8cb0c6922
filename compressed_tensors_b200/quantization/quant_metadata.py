"""names of the quantization parameters a module can carry (mirror of quantization/quant_metadata.py:13-80)"""
from enum import Enum

from torch.nn import Module

__all__ = ["QuantizationMetadata", "KVCacheScaleType"]


class KVCacheScaleType(Enum):
    KEY = "k_scale"
    VALUE = "v_scale"
    QUERY = "q_scale"


class QuantizationMetadata:
    @staticmethod
    def all_qparam_names():
        """every qparam name the lifecycle may register on a module (serialized parameters excluded)"""
        return [KVCacheScaleType.KEY.value, KVCacheScaleType.VALUE.value] + [
            f"{base}_{suffix}" for base in ("input", "weight", "output") for suffix in ("global_scale", "scale", "shape", "zero_point", "g_idx")]

    @classmethod
    def clear_all_qparams(cls, module: Module):
        for key in cls.all_qparam_names():
            if hasattr(module, key):
                delattr(module, key)

    @classmethod
    def clear_quantization(cls, module: Module):
        """remove qparams, the scheme and the wrapped forward; quantization_status / quantization_enabled stay"""
        if hasattr(module.forward, "__wrapped__"):
            module.forward = module.forward.__wrapped__.__get__(module)
        cls.clear_all_qparams(module)
        if hasattr(module, "quantization_scheme"):
            delattr(module, "quantization_scheme")

"""identity compressor for unquantized modules (compressors/dense/base.py:13-60)"""
from ..base import BaseCompressor
from ...config import CompressionFormat

__all__ = ["DenseCompressor"]


@BaseCompressor.register(name=CompressionFormat.dense.value)
class DenseCompressor(BaseCompressor):
    @classmethod
    def compression_param_names(cls, scheme) -> tuple:
        return ("weight",)

    @classmethod
    def compress(cls, state_dict, scheme):
        return state_dict

    @classmethod
    def decompress(cls, state_dict, scheme):
        return state_dict

    @classmethod
    def can_compress(cls, module_type, scheme) -> bool:
        return True

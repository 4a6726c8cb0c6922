"""
Compression format names and the (legacy) sparsity config schema -- mirror of config/base.py:15-104,
config/dense.py, config/sparse_24_bitmask.py:16-29 and config/sparse_bitmask.py:12-25 of the
reference.  String values are the on-disk / config.json vocabulary and must not change.
"""
from __future__ import annotations

from enum import Enum, unique
from typing import List, Optional

from pydantic import BaseModel

from ..registry import RegistryMixin

__all__ = [
    "CompressionFormat",
    "SparsityStructure",
    "SparsityCompressionConfig",
    "DenseSparsityConfig",
    "Sparse24BitMaskConfig",
    "BitmaskConfig",
]


@unique
class CompressionFormat(str, Enum):
    dense = "dense"
    sparse_bitmask = "sparse-bitmask"
    sparse_24_bitmask = "sparse-24-bitmask"
    int_quantized = "int-quantized"
    float_quantized = "float-quantized"
    naive_quantized = "naive-quantized"
    pack_quantized = "pack-quantized"
    marlin_24 = "marlin-24"
    mixed_precision = "mixed-precision"
    nvfp4_pack_quantized = "nvfp4-pack-quantized"
    mxfp4_pack_quantized = "mxfp4-pack-quantized"
    mxfp8_quantized = "mxfp8-quantized"


@unique
class SparsityStructure(Enum):
    """'2:4', 'unstructured' (also None) or '0:0'; lookups are case-insensitive"""

    TWO_FOUR = "2:4"
    UNSTRUCTURED = "unstructured"
    ZERO_ZERO = "0:0"

    @classmethod
    def _missing_(cls, value):
        if value is None:
            return cls.UNSTRUCTURED
        if isinstance(value, str):
            for member in cls:
                if member.value == value.lower():
                    return member
        raise ValueError(f"{value} is not a valid {cls.__name__}")


class SparsityCompressionConfig(RegistryMixin, BaseModel):
    """parameters of a sparsity compressor (format name, targets, ignore list, statistics)"""

    format: str
    targets: Optional[List[str]] = None
    ignore: Optional[List[str]] = None
    global_sparsity: Optional[float] = 0.0
    sparsity_structure: Optional[str] = "unstructured"


@SparsityCompressionConfig.register(name=CompressionFormat.dense.value)
class DenseSparsityConfig(SparsityCompressionConfig):
    format: str = CompressionFormat.dense.value
    global_sparsity: Optional[float] = 0.0
    sparsity_structure: Optional[str] = "unstructured"


@SparsityCompressionConfig.register(name=CompressionFormat.sparse_24_bitmask.value)
class Sparse24BitMaskConfig(SparsityCompressionConfig):
    format: str = CompressionFormat.sparse_24_bitmask.value
    global_sparsity: Optional[float] = 0.0
    sparsity_structure: Optional[str] = SparsityStructure.TWO_FOUR.value


@SparsityCompressionConfig.register(name=CompressionFormat.sparse_bitmask.value)
class BitmaskConfig(SparsityCompressionConfig):
    format: str = CompressionFormat.sparse_bitmask.value
    global_sparsity: Optional[float] = 0.0
    sparsity_structure: Optional[str] = "unstructured"

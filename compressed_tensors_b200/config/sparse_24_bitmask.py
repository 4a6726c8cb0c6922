"""module path of the reference (config/sparse_24_bitmask.py)"""
from .base import Sparse24BitMaskConfig  # noqa: F401

__all__ = ["Sparse24BitMaskConfig"]

"""module path of the reference (config/sparse_bitmask.py)"""
from .base import BitmaskConfig  # noqa: F401

__all__ = ["BitmaskConfig"]

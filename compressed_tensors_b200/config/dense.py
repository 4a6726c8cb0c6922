"""module path of the reference (config/dense.py): the dense (no sparsity) config class lives in config/base.py here"""
from .base import DenseSparsityConfig  # noqa: F401

__all__ = ["DenseSparsityConfig"]

"""per-module on/off switch of the fake-quantization wrapper (mirror of quantization/lifecycle/helpers.py:17-22)"""
from torch.nn import Module

__all__ = ["enable_quantization", "disable_quantization"]


def enable_quantization(module: Module):
    module.quantization_enabled = True


def disable_quantization(module: Module):
    module.quantization_enabled = False

"""marker base class for helper modules that configs must never target (mirror of utils/internal.py:9-18)"""
import torch

__all__ = ["InternalModule"]


class InternalModule(torch.nn.Module):
    pass

"""
Direct (non-recursive) state-dict access used by compress_module / decompress_module
(mirror of utils/module.py:13-65).
"""
from __future__ import annotations

from itertools import chain

import torch

__all__ = ["get_direct_state_dict", "replace_direct_state_dict"]


def get_direct_state_dict(module: torch.nn.Module) -> dict[str, torch.Tensor]:
    """parameters and buffers of this module only, as plain tensors"""
    out = {}
    for name, t in chain(module._parameters.items(), module._buffers.items()):
        out[name] = t.data if isinstance(t, (torch.nn.Parameter, torch.nn.Buffer)) else t
    return out


def replace_direct_state_dict(module: torch.nn.Module, new_state_dict: dict[str, torch.Tensor]):
    """drop names that disappeared, (re)register everything new as frozen Parameters, keep identical tensors"""
    old = get_direct_state_dict(module)
    for name in old:
        if name not in new_state_dict:
            delattr(module, name)
    for name, value in new_state_dict.items():
        if name in old:
            if old[name] is value:
                continue
            delattr(module, name)
        setattr(module, name, torch.nn.Parameter(value, requires_grad=False))

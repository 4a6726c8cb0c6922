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
    """drop names that disappeared, (re)register everything else as frozen Parameters (utils/module.py:34-65).

    Works on the module's `_parameters` / `_buffers` dicts directly: `delattr` / `setattr` go through nn.Module's attribute
    machinery (type checks, hook lookups, three dict probes per call), which made this function the largest share of the per-module
    host time of a whole-model compress -- at 8 ranks every rank re-registers ~1500 tensors of the modules it only mirrors on meta."""
    params, bufs = module._parameters, module._buffers
    for name in [n for n in chain(params, bufs) if n not in new_state_dict]:
        params.pop(name, None)
        bufs.pop(name, None)
    for name, value in new_state_dict.items():
        bufs.pop(name, None)
        module.__dict__.pop(name, None)
        params[name] = torch.nn.Parameter(value, requires_grad=False) if value is not None else None

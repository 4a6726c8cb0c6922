"""
Small host utilities of the path (mirror of the pieces of utils/helpers.py the compressors use):
getattr_chain (:…), patch_attr, tensor_follows_mask_structure (:87-109) and the bitmask pair
pack_bitmasks / unpack_bitmasks (:306-343) -- the latter run as CUDA kernels, not numpy.
"""
from __future__ import annotations

import contextlib
from typing import Any

import torch

from ..ops import pack_bitmasks, unpack_bitmasks

__all__ = ["getattr_chain", "patch_attr", "tensor_follows_mask_structure", "pack_bitmasks", "unpack_bitmasks", "TensorStateDict", "find_unique_name"]

TensorStateDict = dict[str, torch.Tensor]
_MISSING = object()


def getattr_chain(obj: Any, chain_str: str, *args, **kwargs) -> Any:
    """getattr along a dotted path with an optional default (positional or default=...)"""
    if len(args) >= 1:
        default, has_default = args[0], True
    elif "default" in kwargs:
        default, has_default = kwargs["default"], True
    else:
        default, has_default = _MISSING, False
    cur = obj
    for name in chain_str.split("."):
        nxt = getattr(cur, name, _MISSING)
        if nxt is _MISSING or (cur is None):
            if has_default:
                return default
            raise AttributeError(f"{chain_str} not found on {obj}")
        cur = nxt
    return cur


@contextlib.contextmanager
def patch_attr(base: object, attr: str, value: Any):
    """temporarily set base.attr = value"""
    sentinel = object()
    original = getattr(base, attr, sentinel)
    setattr(base, attr, value)
    try:
        yield
    finally:
        if original is sentinel:
            delattr(base, attr)
        else:
            setattr(base, attr, original)


def tensor_follows_mask_structure(tensor: torch.Tensor, mask: str = "2:4") -> bool:
    """at least n zeros in every chunk of m elements; raises ValueError otherwise (helpers.py:87-109)"""
    n, m = (int(v) for v in mask.split(":"))
    zeros = (tensor.reshape(-1, m) == 0).sum(dim=1)
    if not bool(torch.all(zeros >= n)):
        raise ValueError()
    return True


def find_unique_name(name: str, existing_names) -> str:
    """`name` if it is free, else its base (without a trailing _N) with the next free counter: group_0 -> group_1 (helpers.py:506-535)"""
    import re

    used = set(existing_names)
    if name not in used:
        return name
    m = re.match(r"^(.+?)_(\d+)$", name)
    base, n = (m.group(1), int(m.group(2))) if m else (name, 0)
    n += 1
    while f"{base}_{n}" in used:
        n += 1
    return f"{base}_{n}"

"""
Name matching for checkpoint tensors (mirror of utils/match.py:422-445 `match_name` and :469-528 `match_quantizable_tensors`).
Only the name-based matchers the checkpoint converters need live here; module/class matching for live models is in
quantization/lifecycle/apply.py.
"""
from __future__ import annotations

import re
from typing import Iterable, Iterator, Mapping, Optional

import torch

__all__ = ["match_name", "match_quantizable_tensors"]


def match_name(name: str, target: str, fused: Optional[Mapping[str, Iterable[str]]] = None) -> bool:
    """`target` matches `name` exactly, or as a regex when it starts with "re:"; `fused` maps the suffix of a fused module
    (e.g. qkv_proj) to the suffixes of its shards, any of which may match"""
    if fused is not None:
        for fused_suffix, shard_suffixes in fused.items():
            if name.endswith(fused_suffix):
                stem = name.removesuffix(fused_suffix)
                return any(match_name(stem + s, target) for s in shard_suffixes)
    if target.startswith("re:"):
        return re.match(target.removeprefix("re:"), name) is not None
    return target == name


def match_quantizable_tensors(tensors: Mapping[str, torch.Tensor], ignore: Iterable[str], targets: Iterable[str] = tuple(),
                              param_targets: Iterable[str] = ("weight",), allow_nonquantizable: bool = False) -> Iterator[tuple[str, str]]:
    """yield (module name, tensor name) for every tensor whose parameter name is in `param_targets` and whose module is targeted
    (no targets, or "Linear" among them, means every module) and not ignored; modules ending in "norm" are skipped"""
    ignore, targets, param_targets = list(ignore), list(targets), list(param_targets)
    for name in list(tensors.keys()):
        module_name, _, param_name = name.rpartition(".")
        if not allow_nonquantizable and module_name.endswith("norm"):
            continue
        if not any(match_name(param_name, t) for t in param_targets):
            continue
        if not (len(targets) == 0 or "Linear" in targets or any(match_name(module_name, t) for t in targets)):
            continue
        if any(match_name(module_name, i) for i in ignore):
            continue
        yield module_name, name

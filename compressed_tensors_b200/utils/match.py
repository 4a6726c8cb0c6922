"""
Name matching for checkpoint tensors (mirror of utils/match.py:422-445 `match_name` and :469-528 `match_quantizable_tensors`).
Only the name-based matchers the checkpoint converters need live here; module/class matching for live models is in
quantization/lifecycle/apply.py.
"""
from __future__ import annotations

import re
from typing import Iterable, Iterator, Mapping, Optional

import torch

__all__ = ["match_name", "match_quantizable_tensors", "is_match", "match_named_modules", "match_named_parameters", "match_targets"]


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


def _match_class(module: torch.nn.Module, target: str) -> bool:
    """any torch parent class is named `target` (vLLM's LinearBase counts as Linear); utils/match.py:448-466"""
    return any(issubclass(c, torch.nn.Module) and (c.__name__ == target or (c.__name__ == "LinearBase" and target == "Linear"))
               for c in module.__class__.__mro__)


def is_match(name: str, module: torch.nn.Module, targets, ignore=tuple(), fused: Optional[Mapping[str, Iterable[str]]] = None) -> bool:
    """module name or one of its classes matches a target and none of `ignore` (utils/match.py:344-381)"""
    targets = [targets] if isinstance(targets, str) else targets
    ignore = [ignore] if isinstance(ignore, str) else ignore
    from .internal import InternalModule

    return (not isinstance(module, InternalModule) and any(match_name(name, t, fused) or _match_class(module, t) for t in targets)
            and not any(match_name(name, i, fused) or _match_class(module, i) for i in ignore))


def match_named_modules(model: torch.nn.Module, targets, ignore=None, fused: Optional[Mapping[str, Iterable[str]]] = None, warn_on_fail: bool = False):
    """(name, module) of every submodule matching a target and none of `ignore`, in named_modules() order (utils/match.py:34-70)"""
    targets, ignore = list(targets or []), list(ignore or [])
    unmatched = set(targets)
    for name, module in model.named_modules():
        for t in targets:
            if is_match(name, module, t, fused=fused):
                unmatched.discard(t)
                if not is_match(name, module, ignore, fused=fused):
                    yield name, module
                break
    if warn_on_fail:
        import logging

        for t in unmatched:
            logging.getLogger(__name__).warning(f"Could not match `{t}` in instance of {model.__class__.__name__}")


def match_named_parameters(model: torch.nn.Module, targets, ignore=None, fused: Optional[Mapping[str, Iterable[str]]] = None, warn_on_fail: bool = False):
    """(qualified name, parent module, parameter) for parameters whose qualified name matches (utils/match.py:73-113)"""
    targets, ignore = list(targets or []), list(ignore or [])
    unmatched = set(targets)
    for module_name, module in model.named_modules():
        for pname, param in module.named_parameters(recurse=False):
            fqn = f"{module_name}.{pname}"
            for t in targets:
                if match_name(fqn, t, fused):
                    unmatched.discard(t)
                    if not any(match_name(fqn, i, fused) for i in ignore):
                        yield fqn, module, param
    if warn_on_fail:
        import logging

        for t in unmatched:
            logging.getLogger(__name__).warning(f"Could not match `{t}` in instance of {model.__class__.__name__}")


def match_targets(name: str, module: torch.nn.Module, targets) -> list:
    """the targets matching (name, module), most specific first: exact names, regexes, class names (utils/match.py:116-151)"""
    targets = sorted(targets or [], key=lambda x: ("re:" in x, x))
    out = [t for t in targets if match_name(name, t)]
    out += [t for t in targets if _match_class(module, t) and t not in out]
    return out

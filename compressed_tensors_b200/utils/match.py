"""
Target matching of the reference's utils/match.py: names and regexes against checkpoint tensor names (`match_name`,
`match_quantizable_tensors`, what the model-free converters use) and against the modules / parameters of a live model (`is_match`,
`match_named_modules`, `match_named_parameters`, `match_targets`, `match_modules_set`, `is_narrow_match`), which
`apply_quantization_config` and llm-compressor's modifiers use to decide where a scheme applies.
"""
from __future__ import annotations

import logging
import re
from typing import Iterable, Iterator, Mapping, Optional

import torch

_LOGGER = logging.getLogger(__name__)

__all__ = ["match_name", "match_quantizable_tensors", "is_match", "match_named_modules", "match_named_parameters", "match_targets",
           "get_lowest_common_ancestor_name", "match_modules_set", "is_narrow_match"]


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
        for t in unmatched:
            _LOGGER.warning(f"Could not match `{t}` in instance of {model.__class__.__name__}")


def match_named_parameters(model: torch.nn.Module, targets, ignore=None, fused: Optional[Mapping[str, Iterable[str]]] = None, warn_on_fail: bool = False):
    """(qualified name, parent module, parameter) for parameters whose qualified name matches (utils/match.py:73-113)"""
    targets, ignore = list(targets or []), list(ignore or [])
    unmatched = set(targets)
    from .internal import InternalModule

    for module_name, module in model.named_modules():
        if isinstance(module, InternalModule):      # helper modules (observers, transforms) never take part in matching
            continue
        for pname, param in module.named_parameters(recurse=False):
            fqn = f"{module_name}.{pname}"
            for t in targets:
                if match_name(fqn, t, fused):
                    unmatched.discard(t)
                    if not any(match_name(fqn, i, fused) for i in ignore):
                        yield fqn, module, param
    if warn_on_fail:
        for t in unmatched:
            _LOGGER.warning(f"Could not match `{t}` in instance of {model.__class__.__name__}")


def match_targets(name: str, module: torch.nn.Module, targets) -> list:
    """the targets matching (name, module), most specific first: exact names, regexes, class names (utils/match.py:116-151)"""
    targets = sorted(targets or [], key=lambda x: ("re:" in x, x))
    out = [t for t in targets if match_name(name, t)]
    out += [t for t in targets if _match_class(module, t) and t not in out]
    return out


def get_lowest_common_ancestor_name(names) -> str:
    """dotted name of the deepest module that contains every named module (None entries are skipped; "" is the root).  A module is
    its own ancestor: ["a.b", "a.b.c"] -> "a.b", but ["abc", "ab"] -> "" because components, not characters, are compared
    (utils/match.py:154-178)"""
    paths = [n.split(".") if n else [] for n in names if n is not None]
    if not paths:
        return ""
    shared = []
    for parts in zip(*paths):
        if any(p != parts[0] for p in parts):
            break
        shared.append(parts[0])
    return ".".join(shared)


def match_modules_set(model: torch.nn.Module, targets, ignore=None, error_on_module_rematch: bool = True):
    """Walk named_modules() once and yield one group per "parent context": a list with one entry per target, each entry the list of
    modules that matched that target inside the context -- e.g. (q_proj, k_proj, v_proj) per attention block, or one layernorm plus
    ALL experts' up_proj per MoE layer.  A group is complete once every target has at least one match; it is closed (yielded) when the
    next match would move the common ancestor of the group's members.  Leftover matches that never complete a group are an error
    (utils/match.py:181-341)."""
    targets, ignore = list(targets or []), list(ignore or [])
    groups: dict = {t: [] for t in targets}
    missing = set(targets)
    context = None
    for name, module in model.named_modules():
        hits = [t for t in targets if is_match(name, module, t, ignore)]
        if len(hits) > 1 and error_on_module_rematch:
            raise ValueError(f"module: {name} was matched with multiple targets: {set(hits)} which is unexpected "
                             "disable this check by setting `error_on_module_rematch = False`")
        for t in hits:
            widened = get_lowest_common_ancestor_name([name, context])
            if not missing and widened != context:
                # the group is complete and this match lives outside its context: hand the group out and start the next one here
                yield [groups[x] for x in targets]
                groups, missing, widened = {x: [] for x in targets}, set(targets), name
            groups[t].append(module)
            missing.discard(t)
            context = widened
    if len(missing) == len(set(targets)):
        return                                   # nothing matched at all
    if missing:
        raise ValueError(f"Found a final incomplete set with matches found for keys: {set(targets) - missing} "
                         f"but no matches found for keys: {missing}")
    yield [groups[x] for x in targets]


def is_narrow_match(model: torch.nn.Module, targets, name: str, module: Optional[torch.nn.Module] = None) -> bool:
    """some target matches the module itself but neither its parent nor any of its descendants (utils/match.py:384-419)"""
    targets = [targets] if isinstance(targets, str) else list(targets)
    module = module if module is not None else model.get_submodule(name)
    parent_name = name.rsplit(".", 1)[0]
    parent = model.get_submodule(parent_name)

    def below(target: str) -> bool:
        return any(is_match(f"{name}.{child_name}", child, target) for child_name, child in module.named_modules() if child_name)

    return any(is_match(name, module, t) and not is_match(parent_name, parent, t) and not below(t) for t in targets)

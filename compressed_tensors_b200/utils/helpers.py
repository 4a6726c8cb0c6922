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

__all__ = ["getattr_chain", "patch_attr", "patch_attrs", "tensor_follows_mask_structure", "pack_bitmasks", "unpack_bitmasks", "TensorStateDict",
           "find_unique_name", "get_nested_value", "ParameterizedDefaultDict", "fix_fsdp_module_name", "replace_module",
           "is_compressed_tensors_config", "deprecated", "shard_tensor", "combine_shards", "Aliasable", "get_num_attn_heads",
           "get_num_kv_heads", "get_head_dim", "is_accelerator_type"]

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


# ---------------------------------------------------------------------------------------------------------------------------
# small host-side utilities of the reference's utils/helpers.py that its callers (llm-compressor, transformers, vLLM) import
# ---------------------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def patch_attrs(bases, attr: str, values):
    """patch_attr over parallel sequences of objects and values; every original is restored on exit (helpers.py:376-400)"""
    with contextlib.ExitStack() as stack:
        for base, value in zip(bases, values):
            stack.enter_context(patch_attr(base, attr, value))
        yield


def get_nested_value(data_dict, path: str, default=None):
    """data_dict["a"]["b"] for path "a.b"; `default` when a level is missing or not subscriptable (helpers.py:139-146)"""
    node = data_dict
    for key in path.split("."):
        try:
            node = node[key]
        except (KeyError, TypeError):
            return default
    return node


class ParameterizedDefaultDict(dict):
    """a defaultdict whose factory receives the missing key (a tuple key is splatted into positional arguments);
    `get(*key, factory_kwargs=...)` additionally forwards keyword arguments to the factory (helpers.py:403-433)"""

    def __init__(self, default_factory):
        super().__init__()
        self.default_factory = default_factory
        self._factory_kwargs = {}

    def __missing__(self, key):
        args = key if isinstance(key, tuple) else (key,)
        value = self[key] = self.default_factory(*args, **self._factory_kwargs)
        return value

    def get(self, *args, factory_kwargs=None):
        with patch_attr(self, "_factory_kwargs", dict(factory_kwargs or {})):
            return self[args]


_FSDP_WRAPPER = "_fsdp_wrapped_module"


def fix_fsdp_module_name(name: str) -> str:
    """module name without FSDP's wrapper component, wherever it sits (helpers.py:74-84)"""
    return ".".join(part for part in name.split(".") if part != _FSDP_WRAPPER)


def replace_module(model: torch.nn.Module, name: str, new_module: torch.nn.Module) -> None:
    """model.<name> = new_module for a dotted submodule name (helpers.py:112-121)"""
    parent_name, _, child = name.rpartition(".")
    setattr(model.get_submodule(parent_name) if parent_name else model, child, new_module)


def is_compressed_tensors_config(compression_config) -> bool:
    """True for an instance of transformers' CompressedTensorsConfig, False when transformers is absent (helpers.py:124-136)"""
    try:
        from transformers.utils.quantization_config import CompressedTensorsConfig
    except ImportError:
        return False
    return isinstance(compression_config, CompressedTensorsConfig)


def deprecated(future_name: str | None = None, message: str | None = None):
    """decorator: DeprecationWarning on every call, pointing at `future_name` unless `message` replaces the text (helpers.py:180-207)"""
    import functools
    import warnings

    def decorator(func):
        text = message
        if text is None:
            text = f"{func.__name__} is deprecated and will be removed in a future release"
            if future_name is not None:
                text += f". Please use {future_name} instead."

        @functools.wraps(func)
        def wrapped(*args, **kwargs):
            warnings.warn(text, DeprecationWarning, stacklevel=2)
            return func(*args, **kwargs)

        return wrapped

    return decorator


def shard_tensor(tensor: torch.Tensor, shard_sizes, dim: int = 0) -> list:
    """views of `tensor` of the given sizes along `dim`; the sizes must add up (helpers.py:241-270)"""
    if sum(shard_sizes) != tensor.size(dim):
        raise ValueError("Sum of shard_sizes must equal the size of the tensor along the specified dimension.")
    return list(torch.split(tensor, list(shard_sizes), dim=dim))


def combine_shards(shards, dim: int = 0) -> torch.Tensor:
    """the shards of one dtype joined along `dim` into a new tensor (helpers.py:273-303)"""
    if not shards:
        raise ValueError("The list of shards is empty.")
    if len({shard.dtype for shard in shards}) > 1:
        raise ValueError("All shards must have the same dtype.")
    return torch.cat(list(shards), dim=dim)


class Aliasable:
    """mixin for enums whose members have alternative spellings: equality and hashing go through `get_aliases()` (alias -> canonical
    value), so a member equals its alias and plain strings compare by canonical value (helpers.py:210-238)"""

    @staticmethod
    def get_aliases() -> dict:
        raise NotImplementedError()

    def _canonical(self, value):
        return self.get_aliases().get(value, value)

    def __eq__(self, other):
        other_value = other.value if isinstance(other, self.__class__) else other
        return self._canonical(self.value) == self._canonical(other_value)

    def __hash__(self):
        return hash(self._canonical(self.value))


def _config_value(config, what: str, direct: str, numerator: str, denominator: str) -> int:
    if hasattr(config, direct):
        return getattr(config, direct)
    if numerator and hasattr(config, numerator) and hasattr(config, denominator):
        return getattr(config, numerator) // getattr(config, denominator)
    need = f"either `{direct}` or both `{numerator}` and `{denominator}`" if numerator else f"`{direct}`"
    raise ValueError(f"Cannot determine {what} from config. Config must define {need}. {config}")


def get_num_attn_heads(config) -> int:
    """num_attention_heads, or hidden_size // head_dim (helpers.py:436-454)"""
    return _config_value(config, "num_attention_heads", "num_attention_heads", "hidden_size", "head_dim")


def get_num_kv_heads(config) -> int:
    """num_key_value_heads (helpers.py:457-471)"""
    return _config_value(config, "num_key_value_heads", "num_key_value_heads", "", "")


def get_head_dim(config) -> int:
    """head_dim, or hidden_size // num_attention_heads (helpers.py:474-492)"""
    return _config_value(config, "head_dim", "head_dim", "hidden_size", "num_attention_heads")


def is_accelerator_type(device_type: str) -> bool:
    """the device type names the accelerator this process has; False without one (helpers.py:495-503)"""
    if not torch.accelerator.is_available():
        return False
    return device_type == torch.accelerator.current_accelerator().type

"""
Name -> value registry with aliases and "path/to/file.py:Name" plugin loading.

Host-side mirror of the reference's plugin registry (registry/registry.py:56-336): the
compressor plugin API hangs off `BaseCompressor.register(name=...)` /
`BaseCompressor.get_value_from_registry(name)`.  Same lookup rules:
  * names are standardised: '_' and ' ' -> '-', lower case (registry.py:28-42)
  * registering a different value under a taken name raises RuntimeError (registry.py:215-223)
  * "file.py:Name" imports Name from the file (registry.py:241-244, :318-336)
"""
from __future__ import annotations

import importlib.util
import os
from collections import defaultdict
from typing import Any

__all__ = [
    "RegistryMixin",
    "register",
    "get_from_registry",
    "registered_names",
    "registered_aliases",
    "standardize_lookup_name",
    "standardize_alias_name",
    "register_alias",
]

_VALUES: dict[type, dict[str, Any]] = defaultdict(dict)
_ALIASES: dict[type, dict[str, str]] = defaultdict(dict)


def standardize_lookup_name(name: str) -> str:
    return name.replace("_", "-").replace(" ", "-").lower()


def _as_alias_list(alias) -> list[str]:
    if alias is None:
        return []
    if isinstance(alias, str):
        return [standardize_lookup_name(alias)]
    return [standardize_lookup_name(a) for a in alias]


def standardize_alias_name(name):
    """standardize_lookup_name over None / one name / a list of names (registry.py:45-53)"""
    if name is None:
        return None
    return standardize_lookup_name(name) if isinstance(name, str) else [standardize_lookup_name(n) for n in name]


def register_alias(name: str, parent_class: type, alias=None):
    """map the alias(es), and the name itself, to `name` in the parent class's alias table (registry.py:285-318); an alias equal to
    the name, or one that is already taken, is a KeyError"""
    aliases = [] if alias is None else (list(alias) if isinstance(alias, (list, tuple)) else [alias])
    if name in aliases:
        raise KeyError(f"Attempting to register alias {name}, that is identical to the standardized name: {name}.")
    table = _ALIASES[parent_class]
    for a in aliases + [name]:
        if a in table:
            raise KeyError(f"Attempting to register alias {a} as {name} however {a} has already been registered as {table[a]}")
    for a in aliases + [name]:
        table[a] = name


def _check_subclass(parent: type, value: Any):
    if not (isinstance(value, type) and issubclass(value, parent)):
        raise ValueError(f"{value} must be a subclass of {parent} to live in its registry")


def register(parent_class: type, value: Any, name: str | None = None, alias=None, require_subclass: bool = False):
    key = standardize_lookup_name(name if name is not None else value.__name__)
    aliases = _as_alias_list(alias)
    if key in aliases:
        raise KeyError(f"Attempting to register alias {key}, that is identical to the standardized name: {key}.")
    table = _ALIASES[parent_class]
    for a in aliases + [key]:
        if a in table:
            raise KeyError(f"Attempting to register alias {a} as {key} however {a} has already been registered as {table[a]}")
    if require_subclass:
        _check_subclass(parent_class, value)
    existing = _VALUES[parent_class].get(key)
    if existing is not None and existing is not value:
        raise RuntimeError(
            f"Attempting to register name {key} as {value} however {key} has already been registered as {existing}"
        )
    for a in aliases + [key]:
        table[a] = key
    _VALUES[parent_class][key] = value


def _load_from_file(path: str, attr: str) -> Any:
    # the lookup name was lower-cased; recover the attribute case-insensitively
    if not os.path.exists(path):
        candidates = [p for p in (path, path.replace("-", "_")) if os.path.exists(p)]
        if not candidates:
            raise FileNotFoundError(path)
        path = candidates[0]
    spec = importlib.util.spec_from_file_location("ct_b200_plugin_" + str(abs(hash(path))), path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    for cand in dir(module):
        if cand.lower().replace("_", "-") == attr.lower().replace("_", "-"):
            return getattr(module, cand)
    raise AttributeError(f"{attr} not found in {path}")


def get_from_registry(parent_class: type, name: str, require_subclass: bool = False) -> Any:
    if ":" in name:
        path, attr = name.rsplit(":", 1)
        value = _load_from_file(path, attr)
    else:
        key = standardize_lookup_name(name)
        key = _ALIASES[parent_class].get(key, key)
        value = _VALUES[parent_class].get(key)
        if value is None:
            raise KeyError(
                f"Unable to find {key} registered under type {parent_class}.\n"
                f"Registered values for {parent_class}: {registered_names(parent_class)}\n"
                f"Registered aliases for {parent_class}: {registered_aliases(parent_class)}"
            )
    if require_subclass:
        _check_subclass(parent_class, value)
    return value


def registered_names(parent_class: type) -> list[str]:
    return list(_VALUES[parent_class].keys())


def registered_aliases(parent_class: type) -> list[str]:
    return sorted(set(_ALIASES[parent_class]) - set(_VALUES[parent_class]))


class RegistryMixin:
    """mix-in giving a class its own registry of named values (usually subclasses)"""

    registry_requires_subclass: bool = False

    @classmethod
    def register(cls, name: str | None = None, alias=None):
        def decorator(value):
            cls.register_value(value, name=name, alias=alias)
            return value

        return decorator

    @classmethod
    def register_value(cls, value: Any, name: str | None = None, alias=None):
        register(cls, value, name=name, alias=alias, require_subclass=cls.registry_requires_subclass)

    @classmethod
    def get_value_from_registry(cls, name: str):
        return get_from_registry(cls, name, require_subclass=cls.registry_requires_subclass)

    @classmethod
    def load_from_registry(cls, name: str, **constructor_kwargs):
        return cls.get_value_from_registry(name)(**constructor_kwargs)

    @classmethod
    def registered_names(cls) -> list[str]:
        return registered_names(cls)

    @classmethod
    def registered_aliases(cls) -> list[str]:
        return registered_aliases(cls)

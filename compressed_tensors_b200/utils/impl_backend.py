"""
Op-level backend hook, same surface as the reference's ImplBackend (utils/impl_backend.py:23-134):
`ImplBackend.register(name, req, priority)`, `ImplBackend.entrypoint(name)`, `ImplBackend.call`.

In the reference the entrypoint body is the torch-eager fallback.  Here the sm_100a kernels ARE
the implementation, registered at priority 0; an entrypoint whose registered backends all decline
raises instead of silently running something slower (there is no eager path in this engine).
`CT_ENFORCE_EAGER` is therefore rejected loudly.
"""
from __future__ import annotations

import os
from functools import wraps
from typing import Callable

__all__ = ["ImplBackend"]


class ImplBackend:
    _backends: dict[str, list[tuple[int, Callable, Callable]]] = {}
    _entrypoints: dict[str, Callable] = {}

    @classmethod
    def register(cls, fn_name: str, req: Callable[..., bool] = lambda *a, **k: True, priority="0"):
        """register `fn` as an implementation of `fn_name`; priority 'disable' skips registration"""
        def decorator(fn):
            if priority == "disable":
                return fn
            cls._backends.setdefault(fn_name, []).append((int(priority), req, fn))
            cls._backends[fn_name].sort(key=lambda t: t[0])
            return fn
        return decorator

    @classmethod
    def entrypoint(cls, fn_name: str):
        def decorator(body):
            if fn_name in cls._entrypoints:
                raise ValueError(f"entrypoint {fn_name} already defined")

            @wraps(body)
            def wrapper(*args, **kwargs):
                if os.environ.get("CT_ENFORCE_EAGER", "0") not in ("", "0"):
                    raise RuntimeError("CT_ENFORCE_EAGER is set but compressed_tensors_b200 has no eager implementation")
                for _, req, fn in cls._backends.get(fn_name, []):
                    if req(*args, **kwargs):
                        return fn(*args, **kwargs)
                return body(*args, **kwargs)

            cls._entrypoints[fn_name] = wrapper
            return wrapper
        return decorator

    @classmethod
    def call(cls, fn_name: str, *args, **kwargs):
        return cls._entrypoints[fn_name](*args, **kwargs)

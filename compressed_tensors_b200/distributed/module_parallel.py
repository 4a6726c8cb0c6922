"""
replace_module_parallel: apply `apply_fn` (compress_module) to a list of modules across ranks
(mirror of distributed/module_parallel.py:23-90).

  1. modules are dealt to ranks by greedy_bin_packing on their byte size
  2. non-owner ranks run apply_fn on a META copy of the module (shape-only compressor path) so every
     rank ends up with the same parameter names / shapes / dtypes
  3. the owner runs apply_fn for real (one GPU, no collective inside)
  4. recouple: for every resulting tensor the owner broadcasts the data (dist.broadcast, NCCL over
     NVLink on GPUs; float8 viewed as uint8); the reference pickles state dicts through the CPU
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from ..utils.module import get_direct_state_dict, replace_direct_state_dict
from .assign import greedy_bin_packing
from .utils import as_broadcastable, module_size

__all__ = ["replace_module_parallel"]


def _to_meta(module: torch.nn.Module):
    """data -> meta; `*_shape` bookkeeping tensors hold VALUES the shape-only path reads (decompress), they stay real"""
    sd = get_direct_state_dict(module)
    replace_direct_state_dict(module, {k: (v if v is None or k.endswith("shape") else torch.empty_like(v, device="meta")) for k, v in sd.items()})


def _wire_device(dev: torch.device) -> torch.device:
    """where a tensor has to live to be broadcast: NCCL only moves device memory, so a host-resident module goes through this rank's GPU"""
    if dev.type != "cuda" and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return dev


def replace_module_parallel(modules: list, apply_fn: Callable, weight_fn: Callable = module_size, desc: Optional[str] = None,
                            apply_many_fn: Optional[Callable] = None):
    """`apply_many_fn(list_of_modules)` (extension): when given, the owner's modules are processed in one batched call"""
    rank, world = dist.get_rank(), dist.get_world_size()
    devices = {}
    for m in modules:
        tensors = [t for t in get_direct_state_dict(m).values() if t is not None]
        if tensors:
            devices[id(m)] = tensors[0].device
    _, _, owner = greedy_bin_packing(modules, world, weight_fn)

    for m in modules:
        if owner[m] != rank:
            _to_meta(m)
            apply_fn(m)
    mine = [m for m in modules if owner[m] == rank]
    if apply_many_fn is not None:
        apply_many_fn(mine)
    else:
        for m in mine:
            apply_fn(m)

    for m in modules:  # same (sorted) order on every rank
        sd = get_direct_state_dict(m)
        home = devices.get(id(m), torch.device("cpu"))
        dev = _wire_device(home)
        new = {}
        for name in sd:  # identical key order on all ranks (same compressor code path)
            t = sd[name]
            if t is None:
                new[name] = None
                continue
            if owner[m] == rank:
                buf = as_broadcastable(t.contiguous().to(dev))
                dist.broadcast(buf, src=owner[m])
                new[name] = t                                   # the owner keeps its own tensors
            else:
                buf = as_broadcastable(torch.empty(t.shape, dtype=t.dtype, device=dev))
                dist.broadcast(buf, src=owner[m])
                got = buf.view(t.dtype) if buf.dtype != t.dtype else buf
                # tensors the shape-only path already produced for real (e.g. weight_shape, on the CPU) stay where they were
                new[name] = got.to(home) if t.device.type == "meta" else got.to(t.device)
        replace_direct_state_dict(m, new)

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
    # torch.empty(shape, ...) rather than empty_like(device="meta"): the latter goes through torch._refs and drags sympy in (2 s on first use)
    replace_direct_state_dict(module, {k: (v if v is None or k.endswith("shape") else torch.empty(v.shape, dtype=v.dtype, device="meta"))
                                       for k, v in sd.items()})


def _wire_device(dev: torch.device) -> torch.device:
    """where a tensor has to live to be broadcast: NCCL only moves device memory, so a host-resident module goes through this rank's GPU"""
    if dev.type != "cuda" and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu") if dev.type == "meta" else dev


_ALIGN = 256
_BUCKET = 256 << 20      # bytes one owner contributes to one gathered buffer (a tensor larger than this gets a bucket of its own)


def _layout(modules, owner, world):
    """byte layout of every owner's results inside the gathered buffers; identical on every rank (the meta mirror gave every rank the
    same names / shapes / dtypes).  The results travel in BUCKETS: bucket k is one flat buffer of `world` equal slots, slot r holding the
    k-th <= 256 MB run of owner r's tensors in module order.  One buffer for everything would be one collective fewer, but every
    received tensor is a view of its buffer, and a single surviving view (a caller that keeps one weight) would pin all of it; with
    buckets a later decompress_model also returns the memory progressively, bucket by bucket, as it walks the modules.
    Returns (buckets, total_bytes): buckets[k] = (slot_bytes, {rank: [(module, name, offset, nbytes, shape, dtype), ...]})"""
    cur = [0] * world          # bucket an owner is filling
    fill = [0] * world         # bytes it has put there
    slots, entries = [], []

    def bucket(k):
        while len(slots) <= k:
            slots.append(0)
            entries.append({r: [] for r in range(world)})
        return entries[k]

    for m in modules:
        o = owner[m]
        for name, t in get_direct_state_dict(m).items():
            if t is None:
                continue
            nbytes = t.numel() * t.element_size()
            if fill[o] and fill[o] + nbytes > _BUCKET:
                cur[o] += 1
                fill[o] = 0
            bucket(cur[o])[o].append((m, name, fill[o], nbytes, tuple(t.shape), t.dtype))
            fill[o] = (fill[o] + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
            slots[cur[o]] = max(slots[cur[o]], fill[o])
    buckets = [(slots[k], entries[k]) for k in range(len(slots)) if slots[k]]
    return buckets, sum(sl for sl, _ in buckets) * world


def _allgather_fits(modules, owner, world) -> bool:
    """the gathered buffers (every owner's results, once) fit comfortably on EVERY rank -- agreed collectively, or some ranks would
    enter an all_gather and others a broadcast"""
    _, total = _layout(modules, owner, world)
    nccl = dist.get_backend() == "nccl" and torch.cuda.is_available()
    if nccl:
        free_b, _ = torch.cuda.mem_get_info()
        free_b += torch.cuda.memory_reserved() - torch.cuda.memory_allocated()      # blocks the caching allocator holds but does not use
        ok = total <= free_b // 2
    else:
        ok = total <= (4 << 30)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=_wire_device(torch.device("meta")))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def _recouple_allgather(modules, owner, rank, world, devices) -> int:
    buckets, _ = _layout(modules, owner, world)
    if not buckets:
        return 0
    wire = _wire_device(torch.device("meta"))
    received = {}
    moved = 0
    for slot, entries in buckets:
        out = torch.empty(world * slot, dtype=torch.uint8, device=wire)
        mine = out[rank * slot:(rank + 1) * slot]
        for m, name, off, nbytes, _, _ in entries[rank]:
            if nbytes:
                t = get_direct_state_dict(m)[name]
                mine[off:off + nbytes].copy_(as_broadcastable(t.contiguous()).reshape(-1).view(torch.uint8), non_blocking=True)
        dist.all_gather_into_tensor(out, mine)          # in place: this rank's slot is its own contribution
        for r in range(world):
            moved += sum(e[3] for e in entries[r])      # bytes that went through the collective (like the broadcast path counts)
            if r == rank:
                continue                                 # the owner keeps its own tensors
            base = r * slot
            for m, name, off, nbytes, shape, dtype in entries[r]:
                v = out[base + off: base + off + nbytes]
                v = v.view(dtype).reshape(shape) if nbytes else torch.empty(shape, dtype=dtype, device=wire)   # offsets are 256-byte aligned
                received.setdefault(id(m), (m, {}))[1][name] = v
    for m, got in received.values():
        sd = get_direct_state_dict(m)
        home = devices.get(id(m), torch.device("cpu"))
        if home.type == "meta":
            home = wire
        new = {}
        for name, t in sd.items():
            if t is None:
                new[name] = None
                continue
            v = got[name]
            # tensors the shape-only path already produced for real (e.g. weight_shape, on the CPU) stay where they were
            new[name] = v.to(home) if t.device.type == "meta" else v.to(t.device)
        replace_direct_state_dict(m, new)
    return moved


def _recouple_broadcast(modules, owner, rank, devices) -> int:
    moved = 0
    for m in modules:  # same (sorted) order on every rank
        sd = get_direct_state_dict(m)
        home = devices.get(id(m), torch.device("cpu"))
        dev = _wire_device(home)
        if home.type == "meta":
            home = dev
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
            moved += t.numel() * t.element_size()
        replace_direct_state_dict(m, new)
    return moved


def _sync(stats):
    if stats is not None and torch.cuda.is_available():
        torch.cuda.synchronize()


def replace_module_parallel(modules: list, apply_fn: Callable, weight_fn: Callable = module_size, desc: Optional[str] = None,
                            apply_many_fn: Optional[Callable] = None, recouple: bool = True, stats: Optional[dict] = None,
                            mirror_cache: bool = True):
    """Extensions over the reference's signature (all optional):
      apply_many_fn(list_of_modules): the owner's modules are processed in one batched call
      recouple=False: skip step 4 -- every rank keeps only the results of the modules it owns (the others stay on meta); the flow for
                      "each owner writes its own checkpoint shard".  True (default) picks between "allgather" -- every owner packs its
                      results into one flat byte buffer and ONE all_gather_into_tensor moves everything (all owners send at once: the
                      NVSwitch is used from every GPU, where per-tensor broadcasts have one sender at a time) -- and "broadcast" (one
                      dist.broadcast per tensor, frees memory incrementally; chosen when the gathered buffer would not fit comfortably);
                      either can be forced by name
      stats: a dict that receives `apply_s` (this rank's own work + the meta mirror of the others', device-synchronised), its parts
             `owner_host_s` (incl. the final device wait) / `mirror_host_s` / `device_ms` (CUDA events around the owner's launches), `recouple_s`, `recouple_bytes`,
             `owned_modules`, `owned_bytes`; asking for them adds two device synchronisations
    A module whose tensors are on META on a non-owner rank from the start (a model sharded tensor-per-GPU: only the owner ever
    materialised the weight) receives the owner's result on this rank's wire device (its GPU under NCCL)."""
    import time

    rank, world = dist.get_rank(), dist.get_world_size()
    devices = {}
    for m in modules:
        tensors = [t for t in get_direct_state_dict(m).values() if t is not None]
        if tensors:
            devices[id(m)] = tensors[0].device
    _, _, owner = greedy_bin_packing(modules, world, weight_fn)

    _sync(stats)
    t0 = time.perf_counter()
    mine = [m for m in modules if owner[m] == rank]
    others = [m for m in modules if owner[m] != rank]
    owned_bytes = sum(weight_fn(m) for m in mine)
    ev = None
    t_mirror = 0.0

    def mirror():
        """the other ranks' modules, shape-only.  `apply_fn` runs for real on the first module of every signature (class, scheme
        object, names / shapes / dtypes of its tensors, values of the `*_shape` bookkeeping tensors the shape-only path reads); its
        effect -- the resulting meta state dict and quantization_status -- is replayed for the other modules of that signature
        (a 70B model has 560 modules and 4 signatures).  Valid for compress_module / decompress_module, whose meta path is a pure
        function of exactly that signature."""
        nonlocal t_mirror
        t = time.perf_counter()
        plans = {}
        for m in others:
            sd = get_direct_state_dict(m)
            key = (type(m), id(getattr(m, "quantization_scheme", None)), getattr(m, "quantization_status", None),
                   tuple((k, tuple(v.shape), v.dtype, tuple(v.reshape(-1).tolist()) if k.endswith("shape") else None) for k, v in sd.items() if v is not None))
            plan = plans.get(key) if mirror_cache else None
            if plan is None:
                _to_meta(m)
                apply_fn(m)
                if mirror_cache:
                    out = get_direct_state_dict(m)
                    plans[key] = ({k: (None if v is None else ((tuple(v.shape), v.dtype) if v.device.type == "meta" else v)) for k, v in out.items()},
                                  getattr(m, "quantization_status", None))
            else:
                spec, status = plan
                replace_direct_state_dict(m, {k: (None if v is None else (torch.empty(v[0], dtype=v[1], device="meta") if isinstance(v, tuple) else v.clone()))
                                              for k, v in spec.items()})
                if status is not None:
                    m.quantization_status = status
        t_mirror = time.perf_counter() - t

    # Mirroring the other ranks' modules on meta is host-only bookkeeping (the larger share of the host time at 8 ranks).  When those
    # modules hold real tensors here (every rank loaded the whole model) it comes FIRST: it releases their memory before this rank's
    # outputs are allocated.  When they are on meta already (a model sharded tensor-per-rank) the owner's kernels are enqueued first
    # and the mirror runs while the GPU is busy.
    mirror_first = any(devices.get(id(m), torch.device("meta")).type != "meta" for m in others)
    if mirror_first:
        mirror()
    if stats is not None and torch.cuda.is_available():
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if apply_many_fn is not None:
        apply_many_fn(mine)
    else:
        for m in mine:
            apply_fn(m)
    if ev is not None:
        ev[1].record()
    if not mirror_first:
        mirror()
    _sync(stats)
    t1 = time.perf_counter()

    moved, how = 0, "none"
    if recouple:
        how = recouple if isinstance(recouple, str) else "auto"
        if how == "auto":
            how = "allgather" if _allgather_fits(modules, owner, world) else "broadcast"
        if how == "allgather":
            moved = _recouple_allgather(modules, owner, rank, world, devices)
        else:
            moved = _recouple_broadcast(modules, owner, rank, devices)
    _sync(stats)
    if stats is not None:
        stats.update(apply_s=t1 - t0, recouple_s=time.perf_counter() - t1, recouple_bytes=moved, owned_modules=len(mine),
                     owned_bytes=int(owned_bytes), world_size=world, mirror_host_s=t_mirror, owner_host_s=t1 - t0 - t_mirror, recouple_how=how,
                     device_ms=(ev[0].elapsed_time(ev[1]) if ev is not None else None))

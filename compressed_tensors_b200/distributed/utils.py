"""process-group helpers (mirror of distributed/utils.py:56-110)"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

__all__ = ["is_distributed", "init_dist", "as_broadcastable", "module_size", "set_source_process", "get_source_rank", "is_source_process",
           "wait_for_comms"]

import contextlib

_SRC_RANK = 0


@contextlib.contextmanager
def set_source_process(src_rank: int):
    """temporarily make `src_rank` the rank that broadcasts (mirror of distributed/utils.py:33-49)"""
    global _SRC_RANK
    keep, _SRC_RANK = _SRC_RANK, src_rank
    try:
        yield
    finally:
        _SRC_RANK = keep


def get_source_rank() -> int:
    return _SRC_RANK


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def init_dist():
    """one rank per GPU under torchrun: NCCL on CUDA, gloo otherwise"""
    if "TORCHELASTIC_RUN_ID" not in os.environ and "RANK" not in os.environ:
        raise ValueError("Trying to initialize distributed without running under torchrun (no RANK in the environment)")
    rank, local, world = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ["WORLD_SIZE"])
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", init_method="env://", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend="gloo", init_method="env://", rank=rank, world_size=world)
    dist.barrier()


def as_broadcastable(tensor: torch.Tensor) -> torch.Tensor:
    """NCCL cannot move float8: view such tensors as uint8 (utils.py:96-110)"""
    if tensor.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
        return tensor.view(torch.uint8)
    return tensor


def module_size(module: torch.nn.Module) -> int:
    """bytes of the module's own parameters and buffers (load-balancing weight)"""
    total = 0
    for t in list(module._parameters.values()) + list(module._buffers.values()):
        if t is not None:
            total += t.numel() * t.element_size()
    return total


def is_source_process() -> bool:
    """this rank is the one that broadcasts (rank 0 unless `set_source_process` says otherwise); always true outside torch.distributed
    (distributed/utils.py:29-30)"""
    return not is_distributed() or dist.get_rank() == _SRC_RANK


def wait_for_comms(pending_comms: list) -> None:
    """wait() on every async work handle, then empty the list in place so that it can collect the next batch (distributed/utils.py:113-127)"""
    for work in pending_comms:
        work.wait()
    pending_comms.clear()

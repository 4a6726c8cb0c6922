"""
convert_checkpoint (mirror of entrypoints/convert/convert_checkpoint.py:32-134): rewrite a local safetensors checkpoint shard
by shard, without instantiating the model.  B200 flow: `max_workers` threads, each with its own CUDA stream, read a shard
into HBM, run the converter's kernels and write the result, so disk reads, H2D/D2H copies and kernels of different shards
overlap (the C ABI is reentrant and every launch owns its scratch).
"""
from __future__ import annotations

import os
import shutil
from concurrent.futures import ThreadPoolExecutor, as_completed
from pathlib import Path
from typing import Callable, Optional, Union

import torch

from ...utils.safetensors_load import get_checkpoint_files, get_weight_map, is_weights_file, update_safetensors_index
from .convert_file import convert_file, validate_file, write_checkpoint_quantization_config
from .converters import build_inverse_weight_maps

__all__ = ["convert_checkpoint", "exec_jobs"]


def convert_checkpoint(model_stub: Union[str, os.PathLike], save_directory: Union[str, os.PathLike], converter, max_workers: int = 1,
                       device: Optional[Union[str, torch.device]] = None) -> None:
    model_files = get_checkpoint_files(model_stub)
    weight_map = get_weight_map(model_files)
    plans = build_inverse_weight_maps(weight_map=weight_map, model_files=model_files, converters=[converter])

    validate_jobs, convert_jobs = [], []
    for shard_name, resolved in model_files.items():
        save_path = Path(save_directory) / shard_name
        if shard_name.endswith("safetensors"):
            if shard_name not in plans:
                raise ValueError(f"Could not find inverse_weight_map for shard {shard_name}")
            validate_jobs.append((validate_file, plans[shard_name], converter))
            convert_jobs.append((convert_file, plans[shard_name], save_path, converter, device))
        elif str(resolved) != str(save_path):
            save_path.parent.mkdir(parents=True, exist_ok=True)   # configs, tokenizers, ...: copied as they are
            shutil.copyfile(resolved, save_path)

    exec_jobs(validate_jobs, max_workers, desc="Validating")
    total_size, new_map = 0, {}
    for size, part in exec_jobs(convert_jobs, max_workers, desc="Converting"):
        total_size += size
        new_map.update(part)
    write_checkpoint_quantization_config(save_directory, converter)
    update_safetensors_index(save_directory, total_size, new_map)


def exec_jobs(jobs: list, max_workers: int = 1, desc: str = "Executing Jobs") -> list:
    """run (callable, *args) tuples, in a thread pool when max_workers > 1"""
    if max_workers == 1:
        return [job[0](*job[1:]) for job in jobs]
    results = []
    with ThreadPoolExecutor(max_workers) as pool:
        for fut in as_completed([pool.submit(*job) for job in jobs]):
            results.append(fut.result())
    return results

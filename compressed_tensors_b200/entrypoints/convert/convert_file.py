"""
Per-shard jobs of the model-free conversion (mirror of entrypoints/convert/convert_file.py:27-121), with one addition:
`device` -- the shard is read straight into that device's memory and the converter's kernels run there; results are
brought back to the host only to be written.
"""
from __future__ import annotations

import json
import os
from typing import Optional, Union

import torch
from safetensors.torch import save_file

from ... import __version__ as ct_version
from ...utils.safetensors_load import QUANTIZATION_CONFIG_NAME, find_config_path, load_tensors_from_inverse_weight_map

__all__ = ["validate_file", "convert_file", "write_checkpoint_quantization_config"]

COMPRESSION_VERSION_NAME = "version"


def write_checkpoint_quantization_config(save_directory: Union[str, os.PathLike], converter) -> None:
    """replace (or remove, for a converter that dequantizes) the quantization_config of config.json / params.json"""
    data = None
    qc = converter.create_config()
    if qc is not None:
        data = qc.model_dump()
        data[COMPRESSION_VERSION_NAME] = ct_version
    path = find_config_path(save_directory)
    if path is None:
        return
    with open(path, "r") as f:
        config = json.load(f)
    if data is None:
        if QUANTIZATION_CONFIG_NAME in config:
            del config[QUANTIZATION_CONFIG_NAME]
        elif QUANTIZATION_CONFIG_NAME in config.get("text_config", {}):
            del config["text_config"][QUANTIZATION_CONFIG_NAME]
    else:
        config[QUANTIZATION_CONFIG_NAME] = data
    with open(path, "w") as f:
        json.dump(config, f, indent=2, sort_keys=True)


def validate_file(inverse_weight_map: dict, converter) -> None:
    """shape-only pass (meta tensors): cheap enough to run over the whole checkpoint before any kernel is launched"""
    converter.validate(load_tensors_from_inverse_weight_map(inverse_weight_map, device="meta"))


def convert_file(inverse_weight_map: dict, save_path: Union[str, os.PathLike], converter, device: Optional[Union[str, torch.device]] = None):
    """load -> converter.process -> save_file; returns (bytes written, {tensor name: shard file name})"""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if dev.type == "cuda":
        stream = torch.cuda.Stream(device=dev)   # own stream per job: concurrent jobs overlap their copies and kernels
        with torch.cuda.stream(stream):
            tensors = converter.process(load_tensors_from_inverse_weight_map(inverse_weight_map, device=dev))
            tensors = {k: v.contiguous().to("cpu", non_blocking=False) for k, v in tensors.items()}
        stream.synchronize()
    else:
        tensors = converter.process(load_tensors_from_inverse_weight_map(inverse_weight_map, device=dev))
        tensors = {k: v.contiguous() for k, v in tensors.items()}
    os.makedirs(os.path.dirname(os.fspath(save_path)) or ".", exist_ok=True)
    save_file(tensors, os.fspath(save_path))
    total = sum(t.nbytes for t in tensors.values())
    return total, {k: os.path.basename(os.fspath(save_path)) for k in tensors}

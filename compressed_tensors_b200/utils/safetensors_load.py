"""
Checkpoint-directory helpers for the model-free converters (mirror of the parts of utils/safetensors_load.py the convert
entrypoint uses: :61-258, :470-521).  Local directories only -- this engine has no network path; Hub stubs must be
downloaded first.  `load_tensors_from_inverse_weight_map(device="cuda:N")` reads shards straight into device memory
(safetensors' own device loader), which is what the GPU converters want.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Optional, Union

import torch
from safetensors import safe_open

__all__ = [
    "InverseWeightMap", "load_tensors_from_inverse_weight_map", "find_config_path", "get_quantization_config",
    "find_safetensors_index_path", "find_safetensors_index_file", "get_weight_map", "update_safetensors_index",
    "is_weights_file", "get_checkpoint_files", "get_safetensors_header", "match_param_name", "get_weight_mappings",
    "get_nested_weight_mappings", "get_quantization_parameter_to_path_mapping", "is_quantization_param", "get_safetensors_folder",
]

CONFIG_NAME = "config.json"
SAFE_WEIGHTS_NAME = "model.safetensors"
SAFE_WEIGHTS_INDEX_NAME = "model.safetensors.index.json"
QUANTIZATION_CONFIG_NAME = "quantization_config"

_ST_DTYPES = {"BOOL": torch.bool, "U8": torch.uint8, "I8": torch.int8, "I16": torch.int16, "U16": torch.uint16, "F16": torch.float16,
              "BF16": torch.bfloat16, "I32": torch.int32, "U32": torch.uint32, "F32": torch.float32, "F64": torch.float64,
              "I64": torch.int64, "U64": torch.uint64, "F8_E4M3": torch.float8_e4m3fn, "F8_E5M2": torch.float8_e5m2}

InverseWeightMap = dict  # resolved shard path -> list of tensor names to read from it (None / empty = all)


def is_weights_file(file_name: str) -> bool:
    return file_name.endswith((".bin", ".safetensors", ".pth", ".msgpack", ".pt"))


def get_checkpoint_files(model_stub: Union[str, os.PathLike]) -> dict[str, str]:
    """relative path -> absolute path of every file under a LOCAL checkpoint directory (hidden cache files skipped)"""
    root = os.fspath(model_stub)
    if not os.path.isdir(root):
        raise ValueError(f"{root} is not a local directory: compressed_tensors_b200 converts local checkpoints only (no network)")
    out = {}
    for dirpath, _, filenames in os.walk(root):
        for fn in filenames:
            rel = os.path.relpath(os.path.join(dirpath, fn), root)
            if not rel.startswith((".cache", ".gitattributes")):
                out[rel] = os.path.join(root, rel)
    return out


def find_safetensors_index_path(save_directory: Union[str, os.PathLike]) -> Optional[str]:
    for fn in os.listdir(save_directory):
        if fn.endswith("safetensors.index.json"):
            return os.path.join(save_directory, fn)
    return None


def find_config_path(save_directory: Union[str, os.PathLike]) -> Optional[str]:
    names = os.listdir(save_directory)
    for candidate in (CONFIG_NAME, "params.json"):
        if candidate in names:
            return os.path.join(save_directory, candidate)
    return None


def get_quantization_config(config_path: str) -> Optional[dict]:
    """quantization_config, else text_config.quantization_config, else compression_config (the cascade vLLM uses)"""
    with open(config_path, "r") as f:
        config = json.load(f)
    if QUANTIZATION_CONFIG_NAME in config:
        return config[QUANTIZATION_CONFIG_NAME]
    if QUANTIZATION_CONFIG_NAME in config.get("text_config", {}):
        return config["text_config"][QUANTIZATION_CONFIG_NAME]
    return config.get("compression_config")


def find_safetensors_index_file(model_files: dict[str, str]) -> Optional[str]:
    for suffix in (SAFE_WEIGHTS_INDEX_NAME, ".safetensors.index.json"):
        for rel, path in model_files.items():
            if rel.endswith(suffix):
                return path
    return None


def get_weight_map(model_files: dict[str, str]) -> dict[str, str]:
    """tensor name -> shard file name, from the index json or, for a single-file checkpoint, from model.safetensors itself"""
    index = find_safetensors_index_file(model_files)
    if index is not None:
        with open(index, "r") as f:
            return json.load(f)["weight_map"]
    if SAFE_WEIGHTS_NAME not in model_files:
        raise ValueError(f"File {SAFE_WEIGHTS_NAME} expected but not found in {model_files.keys()}")
    with safe_open(model_files[SAFE_WEIGHTS_NAME], framework="pt") as f:
        return {name: SAFE_WEIGHTS_NAME for name in f.keys()}


def update_safetensors_index(save_directory: Union[str, os.PathLike], total_size: int, weight_map: dict[str, str]) -> None:
    path = find_safetensors_index_path(save_directory) or os.path.join(save_directory, SAFE_WEIGHTS_INDEX_NAME)
    with open(path, "w") as f:
        json.dump({"metadata": {"total_size": total_size}, "weight_map": weight_map}, f, indent=2, sort_keys=True)


def load_tensors_from_inverse_weight_map(inverse_weight_map: InverseWeightMap,
                                         device: Union[str, torch.device] = torch.device("cpu")) -> dict[str, torch.Tensor]:
    """read the listed tensors of every shard; `device` may be "meta" (shapes only), "cpu" or a CUDA device (read straight to HBM)"""
    device = torch.device(device)
    tensors: dict[str, torch.Tensor] = {}
    for source_file, names in inverse_weight_map.items():
        st_device = "cpu" if device.type == "meta" else str(device)
        with safe_open(source_file, framework="pt", device=st_device) as f:
            keys = set(f.keys())
            for name in (names if names else sorted(keys)):
                if name not in keys:
                    raise ValueError(f"Expected to find tensor {name} in {source_file}, but tensor was not found.")
                if device.type == "meta":
                    sl = f.get_slice(name)
                    tensors[name] = torch.empty(size=sl.get_shape(), dtype=_ST_DTYPES[sl.get_dtype()], device="meta")
                else:
                    tensors[name] = f.get_tensor(name)
    return tensors


# ---------------------------------------------------------------------------------------------------------------------------
# name -> file maps of a saved (possibly compressed) checkpoint, as the reference's loaders use them (safetensors_load.py:302-538)
# ---------------------------------------------------------------------------------------------------------------------------
def get_safetensors_header(safetensors_path: str) -> dict:
    """the JSON header of one .safetensors file: 8-byte little-endian length, then that many bytes of JSON"""
    import struct

    with open(safetensors_path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        return json.loads(f.read(n))


def match_param_name(full_name: str, param_name: str) -> Optional[str]:
    """"model.layers.0.q_proj" for ("model.layers.0.q_proj.weight_packed", "weight_packed"); None when the suffix differs"""
    suffix = "." + param_name
    return full_name[: -len(suffix)] if full_name.endswith(suffix) and len(full_name) > len(suffix) else None


def get_weight_mappings(path_to_model_or_tensors: str) -> dict[str, str]:
    """tensor name -> file holding it, for one .safetensors file, a directory with model.safetensors, or a sharded directory with
    an index (paths joined onto the directory)"""
    path = str(path_to_model_or_tensors)
    if os.path.isfile(path):
        return {name: path for name in get_safetensors_header(path) if name != "__metadata__"}
    single, index = os.path.join(path, SAFE_WEIGHTS_NAME), os.path.join(path, SAFE_WEIGHTS_INDEX_NAME)
    if os.path.exists(single):
        return {name: single for name in get_safetensors_header(single) if name != "__metadata__"}
    if os.path.exists(index):
        with open(index, "r", encoding="utf-8") as f:
            return {name: os.path.join(path, shard) for name, shard in json.load(f)["weight_map"].items()}
    raise ValueError(f"Could not find a safetensors weight or index file at {path}")


def get_nested_weight_mappings(model_path: str, params_to_nest: Iterable[str], return_unmatched_params: bool = False):
    """{module: {param: file}} for the tensor names that end in one of `params_to_nest`; optionally also the flat map of the names
    that matched none (what a second, stacked compressor still needs)"""
    params_to_nest = list(params_to_nest)
    nested: dict[str, dict[str, str]] = {}
    unmatched: dict[str, str] = {}
    for name, location in get_weight_mappings(model_path).items():
        hit = False
        for param in params_to_nest:
            module = match_param_name(name, param)
            if module:
                nested.setdefault(module, {})[param] = location
                hit = True
        if not hit and return_unmatched_params:
            unmatched[name] = location
    return (nested, unmatched) if return_unmatched_params else nested


def is_quantization_param(name: str) -> bool:
    """a tensor name that ends in "_scale" (global scales included), "zero_point" or "g_idx" (safetensors_load.py:524-538)"""
    return name.endswith(("_scale", "zero_point", "g_idx"))


def get_quantization_parameter_to_path_mapping(model_path: str) -> dict[str, str]:
    """the quantization parameters of a checkpoint and the files they live in"""
    return {name: path for name, path in get_weight_mappings(model_path).items() if is_quantization_param(name)}


def get_safetensors_folder(pretrained_model_name_or_path: str, cache_dir: Optional[str] = None) -> str:
    """the local folder that holds a model's safetensors files (safetensors_load.py:260-299).  A path that exists is returned as is
    (absolute); a Hub id is looked up in the local Hugging Face cache only -- this engine never downloads."""
    path = str(pretrained_model_name_or_path)
    if os.path.exists(path):
        return os.path.abspath(path)
    try:
        from transformers.utils import cached_file

        for name in (SAFE_WEIGHTS_NAME, SAFE_WEIGHTS_INDEX_NAME):
            found = cached_file(path, name, cache_dir=cache_dir, local_files_only=True, _raise_exceptions_for_missing_entries=False,
                                _raise_exceptions_for_connection_errors=False)
            if found is not None:
                return os.path.split(found)[0]
    except Exception:  # noqa: BLE001  (no transformers, malformed id, offline cache miss)
        pass
    raise ValueError(f"Could not locate safetensors weight or index file from {pretrained_model_name_or_path}.")

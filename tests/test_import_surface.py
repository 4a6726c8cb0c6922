"""
The module paths and names the reference's hot-path tests and its external callers import, on this package
directly and through the drop-in alias `compat/compressed_tensors` (checked in a fresh interpreter, because the alias must not share
a process with `compressed_tensors_b200`).
"""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SURFACE = {
    "": ["ModelCompressor", "PackedQuantizationCompressor", "IntQuantizationCompressor", "FloatQuantizationCompressor", "CompressionFormat", "__version__"],
    "compressors": ["BaseCompressor", "ModelCompressor", "compress_module", "decompress_module", "COMPRESSIBLE_MODULE_TYPES"],
    "compressors.base": ["BaseCompressor", "compress_module", "decompress_module"],
    "compressors.model_compressors.model_compressor": ["ModelCompressor"],
    "compressors.pack_quantized.helpers": ["pack_to_int32", "unpack_from_int32"],
    "compressors.nvfp4.helpers": ["pack_fp4_to_uint8", "unpack_fp4_from_uint8"],
    "compressors.nvfp4.base": ["NVFP4PackedCompressor"],
    "compressors.mxfp4.base": ["MXFP4PackedCompressor"],
    "compressors.mxfp8.base": ["MXFP8QuantizationCompressor"],
    "compressors.mx_utils": ["compress_mx_scale", "decompress_mx_scale"],
    "compressors.format": ["infer_module_format", "infer_model_format"],
    "config": ["CompressionFormat", "SparsityCompressionConfig", "SparsityStructure", "DenseSparsityConfig", "Sparse24BitMaskConfig", "BitmaskConfig"],
    "quantization": ["QuantizationArgs", "QuantizationScheme", "QuantizationConfig", "QuantizationStatus", "QuantizationStrategy", "QuantizationType",
                     "ActivationOrdering", "DynamicType", "FP8_E4M3_DATA", "FP4_E2M1_DATA", "apply_quantization_config",
                     "initialize_module_for_quantization", "preset_name_to_scheme", "is_preset_scheme", "KVCacheScaleType", "QuantizationMetadata"],
    "quantization.lifecycle.forward": ["quantize", "dequantize", "fake_quantize", "forward_quantize", "set_forward_quantized", "_process_quantization"],
    "quantization.lifecycle.forward_helpers": ["_quantize", "_dequantize", "_quantize_dequantize", "_is_fp8_supported", "adapt_scale_and_zp_for_triton"],
    "quantization.lifecycle.initialize": ["initialize_module_for_quantization"],
    "quantization.quant_args": ["round_to_quantized_type_args", "round_to_quantized_type_dtype"],
    "quantization.utils": ["calculate_qparams", "calculate_range", "is_module_quantized", "compute_dynamic_scales_and_zp", "generate_gparam", "strategy_cdiv",
                           "maybe_pad_tensor_for_block_quant", "calculate_block_padding", "generate_mx_scales", "round_to_power_2"],
    "quantization.utils.helpers": ["calculate_qparams"],
    "utils": ["get_direct_state_dict", "replace_direct_state_dict", "getattr_chain", "patch_attr", "TensorStateDict", "is_match", "pack_bitmasks", "unpack_bitmasks",
              "match_named_modules", "match_named_parameters", "match_targets", "InternalModule"],
    "utils.impl_backend": ["ImplBackend"],
    "utils.match": ["match_name", "match_quantizable_tensors", "is_match"],
    "utils.safetensors_load": ["get_checkpoint_files", "get_weight_map", "update_safetensors_index", "load_tensors_from_inverse_weight_map", "find_config_path"],
    "distributed": ["greedy_bin_packing", "replace_module_parallel", "init_dist", "is_distributed", "set_source_process", "as_broadcastable"],
    "offload": ["update_offload_parameter", "disable_onloading", "get_execution_device", "is_distributed", "offload_module", "set_onload_device",
                "OffloadCache", "to_meta", "as_single_threaded", "module_size"],
    "compressors.model_compressors": ["ModelCompressor"],
    "config.dense": ["DenseSparsityConfig"],
    "config.sparse_24_bitmask": ["Sparse24BitMaskConfig"],
    "config.sparse_bitmask": ["BitmaskConfig"],
    "quantization.utils.fp4_utils": ["cast_to_fp4"],
    "quantization.lifecycle.apply": ["apply_quantization_config", "load_pretrained_quantization_parameters"],
    "distributed.utils": ["is_source_process", "wait_for_comms", "set_source_process", "as_broadcastable"],
    "utils.helpers": ["Aliasable", "get_num_attn_heads", "get_num_kv_heads", "get_head_dim", "is_accelerator_type", "ParameterizedDefaultDict",
                      "patch_attrs", "get_nested_value", "deprecated", "shard_tensor", "combine_shards", "replace_module", "fix_fsdp_module_name"],
    "utils.type": ["TorchDtype", "TensorStateDict"],
    "registry.registry": ["RegistryMixin", "register_alias", "standardize_alias_name", "standardize_lookup_name"],
    "transform": ["TransformConfig", "TransformArgs", "TransformScheme", "TransformLocation"],
    "quantization.quant_metadata": ["KVCacheScaleType", "QuantizationMetadata"],
    "quantization.lifecycle.helpers": ["enable_quantization", "disable_quantization"],
    "entrypoints.convert": ["convert_checkpoint", "AutoAWQConverter", "FP8BlockDequantizer", "CompressedTensorsDequantizer", "ModelOptNvfp4Converter",
                            "Converter", "build_inverse_weight_maps", "convert_file", "validate_file"],
}


@pytest.mark.parametrize("sub", sorted(SURFACE))
def test_surface_on_the_package(sub):
    mod = importlib.import_module("compressed_tensors_b200" + ("." + sub if sub else ""))
    missing = [n for n in SURFACE[sub] if not hasattr(mod, n)]
    assert not missing, f"compressed_tensors_b200.{sub} lacks {missing}"


def test_surface_through_the_drop_in_alias():
    code = (
        "import importlib, json, sys\n"
        f"surface = {SURFACE!r}\n"
        "import compressed_tensors as ct\n"
        "assert 'compressed_tensors_b200' not in sys.modules\n"
        "bad = {}\n"
        "for sub, names in surface.items():\n"
        "    m = importlib.import_module('compressed_tensors' + ('.' + sub if sub else ''))\n"
        "    miss = [n for n in names if not hasattr(m, n)]\n"
        "    if miss: bad[sub] = miss\n"
        "from compressed_tensors.compressors import BaseCompressor\n"
        "from compressed_tensors.quantization import preset_name_to_scheme\n"
        "from compressed_tensors.compressors.format import infer_module_format\n"
        "import torch\n"
        "fmt = infer_module_format(torch.nn.Linear, preset_name_to_scheme('W4A16', ['Linear']))\n"
        "assert BaseCompressor.get_value_from_registry(fmt.value).__module__.startswith('compressed_tensors.compressors.pack_quantized')\n"
        "assert 'compressed_tensors_b200' not in sys.modules, 'the alias must load ONE namespace'\n"
        "print(json.dumps(bad))\n"
    )
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.strip().splitlines()[-1] == "{}", r.stdout

"""
Host side of the checkpoint converters (no GPU): the oracle's AWQ / FP8-block restatement against golden vectors produced by
the reference, and the planning / validation / config logic of the converters, written after the reference's own tests
(tests/test_entrypoints/convert/converters/*.py).
"""
import json

import pytest
import torch
from safetensors.torch import save_file

import oracle
from compressed_tensors_b200.config import CompressionFormat
from compressed_tensors_b200.entrypoints.convert import (AutoAWQConverter, CompressedTensorsDequantizer, FP8BlockDequantizer, ModelOptNvfp4Converter,
                                                         build_inverse_weight_maps)
from compressed_tensors_b200.quantization import QuantizationArgs, QuantizationConfig, QuantizationScheme, QuantizationStatus
from compressed_tensors_b200.utils.match import match_name, match_quantizable_tensors
from compressed_tensors_b200.utils.safetensors_load import get_checkpoint_files, get_weight_map, load_tensors_from_inverse_weight_map
from tests.golden import load
from tests.util import same

G = load("convert")


@pytest.mark.parametrize("i", range(len(G["awq"])))
def test_oracle_awq_repack_golden(i):
    c = G["awq"][i]
    same(oracle.awq_repack(c["qweight"]), c["result"]["m.q_proj.weight_packed"], "awq weight_packed")
    if c["zero_point"]:
        same(oracle.awq_repack_zeros(c["qzeros"]), c["result"]["m.q_proj.weight_zero_point"], "awq weight_zero_point")


@pytest.mark.parametrize("i", range(len(G["fp8block"])))
def test_oracle_fp8_block_golden(i):
    c = G["fp8block"][i]
    same(oracle.dequantize_block_fp8(c["weight"], c["scale_inv"], c["block"], c["out"].dtype), c["out"], "fp8 block dequantize")


def test_match_name_and_tensors():
    assert match_name("model.layers.0.mlp.up_proj", "re:.*up_proj$") and not match_name("a.b", "re:b")
    assert match_name("lm_head", "lm_head") and not match_name("lm_head2", "lm_head")
    assert match_name("model.qkv_proj", "re:.*q_proj$", fused={"qkv_proj": ["q_proj", "k_proj", "v_proj"]})
    t = {"a.mlp.weight": 1, "a.mlp.weight_scale": 2, "a.input_layernorm.weight": 3, "lm_head.weight": 4, "b.attn.weight": 5}
    got = list(match_quantizable_tensors(t, ignore=["lm_head"], targets=["re:.*mlp$"], param_targets=["weight"]))
    assert got == [("a.mlp", "a.mlp.weight")]
    assert {n for _, n in match_quantizable_tensors(t, ignore=[], targets=[])} == {"a.mlp.weight", "lm_head.weight", "b.attn.weight"}


def _fp8_checkpoint(tmp_path):
    d = tmp_path / "model"
    d.mkdir()
    f1 = {"embed_tokens.weight": torch.randn(128, 128), "layer0.weight": torch.randn(128, 128).to(torch.float8_e4m3fn),
          "layer1.weight_scale_inv": torch.rand(1, 1) + 0.5}
    f2 = {"layer0.weight_scale_inv": torch.rand(1, 1) + 0.5, "layer1.weight": torch.randn(128, 128).to(torch.float8_e4m3fn),
          "layer2.weight": torch.randn(128, 128).to(torch.float8_e4m3fn), "layer2.weight_scale_inv": torch.rand(1, 1) + 0.5,
          "lm_head.weight": torch.randn(128, 128)}
    save_file(f1, str(d / "model-00001-of-00002.safetensors"))
    save_file(f2, str(d / "model-00002-of-00002.safetensors"))
    wm = {k: "model-00001-of-00002.safetensors" for k in f1}
    wm.update({k: "model-00002-of-00002.safetensors" for k in f2})
    (d / "model.safetensors.index.json").write_text(json.dumps({"metadata": {"total_size": 0}, "weight_map": wm}))
    (d / "config.json").write_text(json.dumps({"model_type": "test", "quantization_config": {"quant_method": "fp8"}}))
    return d, wm, {**f1, **f2}


def test_build_inverse_weight_maps_moves_partners_together(tmp_path):
    d, wm, _ = _fp8_checkpoint(tmp_path)
    files = get_checkpoint_files(d)
    assert get_weight_map(files) == wm
    conv = FP8BlockDequantizer(targets=[r"re:.*layer\d.*"])
    plans = build_inverse_weight_maps(weight_map=wm, model_files=files, converters=[conv])
    assert set(plans) == {"model-00001-of-00002.safetensors", "model-00002-of-00002.safetensors"}
    seen = [n for plan in plans.values() for names in plan.values() for n in names]
    assert sorted(seen) == sorted(wm), "every tensor exactly once"
    shard1 = {n for names in plans["model-00001-of-00002.safetensors"].values() for n in names}
    assert {"layer0.weight", "layer0.weight_scale_inv"} <= shard1 and "layer1.weight_scale_inv" not in shard1
    meta = load_tensors_from_inverse_weight_map(plans["model-00001-of-00002.safetensors"], device="meta")
    assert meta["layer0.weight"].device.type == "meta" and meta["layer0.weight"].dtype == torch.float8_e4m3fn
    conv.validate(meta)
    with pytest.raises(ValueError, match="not found in weight map"):
        build_inverse_weight_maps({k: v for k, v in wm.items() if k != "layer0.weight_scale_inv"}, files, [conv])


def test_fp8block_validate_and_dependencies():
    conv = FP8BlockDequantizer(targets=[r"re:.*proj$"])
    assert conv.get_dependencies("m.q_proj.weight") == {"m.q_proj.weight_scale_inv"}
    assert conv.get_dependencies("m.norm.weight") == set() and conv.create_config() is None
    with pytest.raises(ValueError, match="without corresponding weight_scale_inv"):
        conv.validate({"m.q_proj.weight": torch.empty(1, device="meta")})
    with pytest.raises(ValueError, match="unexpected non-targeted"):
        conv.validate({"m.other.weight_scale_inv": torch.empty(1, device="meta")})


def test_autoawq_config_dependencies_validate():
    conv = AutoAWQConverter.from_autoawq_config({"bits": 4, "group_size": 64, "zero_point": True, "version": "gemm", "modules_to_not_convert": ["vision_tower"]})
    cfg = conv.create_config()
    scheme = cfg.config_groups["config_group_0"]
    assert cfg.format == CompressionFormat.pack_quantized.value and cfg.quantization_status == QuantizationStatus.COMPRESSED
    assert cfg.ignore == ["lm_head", "re:.*vision_tower.*"] and scheme.format == CompressionFormat.pack_quantized.value
    assert scheme.weights.num_bits == 4 and scheme.weights.group_size == 64 and scheme.weights.symmetric is False
    c2 = AutoAWQConverter(targets=[r"re:.*down_proj$"])
    assert c2.get_dependencies("m.mlp.down_proj.qweight") == {"m.mlp.down_proj.qzeros", "m.mlp.down_proj.scales"}
    assert c2.get_dependencies("m.mlp.up_proj.qweight") == set()
    assert AutoAWQConverter(targets=[r"re:.*down_proj$"], zero_point=False).get_dependencies("m.mlp.down_proj.qweight") == {"m.mlp.down_proj.scales"}
    with pytest.raises(ValueError, match="without corresponding"):
        AutoAWQConverter().validate({"m.mlp.down_proj.qweight": torch.zeros(1, 1, device="meta")})
    with pytest.raises(ValueError):
        AutoAWQConverter(bits=8)
    # the reference's helper statics keep working on the CPU (plain torch indexing, no kernel)
    vals = torch.tensor([[0, 1, 2, 3, 4, 5, 6, 7]], dtype=torch.int32)
    packed = torch.zeros(1, 1, dtype=torch.int32)
    for off in range(8):
        packed |= vals[:, off::8] << (off * 4)
    un, _ = AutoAWQConverter.unpack_awq(packed, None, bits=4)
    ro, _ = AutoAWQConverter.reverse_awq_order(un, None, bits=4)
    assert torch.equal(un & 15, vals.to(torch.int8)) and torch.equal(ro & 15, torch.tensor([[0, 4, 1, 5, 2, 6, 3, 7]], dtype=torch.int8))


def _ct_dequantizer(ignore=None):
    dq = object.__new__(CompressedTensorsDequantizer)
    dq.dtype = torch.bfloat16
    scheme = QuantizationScheme(targets=["re:.*mlp.*"], weights=QuantizationArgs(num_bits=8, type="int", strategy="channel", symmetric=True, dynamic=False),
                                format=CompressionFormat.naive_quantized)
    dq.quant_config = QuantizationConfig(config_groups={"group_0": scheme}, ignore=ignore or [])
    return dq


def _ct_tensors(device="meta"):
    return {"model.layers.0.mlp.up_proj.weight": torch.empty(64, 64, dtype=torch.int8, device=device),
            "model.layers.0.mlp.up_proj.weight_scale": torch.empty(64, 1, device=device),
            "model.layers.0.mlp.down_proj.weight": torch.empty(64, 64, dtype=torch.int8, device=device),
            "model.layers.0.mlp.down_proj.weight_scale": torch.empty(64, 1, device=device),
            "model.language_model.layers.0.input_layernorm.weight": torch.empty(64, 1, dtype=torch.bfloat16, device=device),
            "model.layers.0.self_attn.q_proj.weight": torch.empty(128, 64, dtype=torch.bfloat16, device=device),
            "model.embed_tokens.weight": torch.empty(128, 64, dtype=torch.bfloat16, device=device)}


def test_ct_dequantizer_validate_and_dependencies():
    dq = _ct_dequantizer(ignore=["model.embed_tokens"])
    dq.validate(_ct_tensors())
    t = _ct_tensors()
    del t["model.layers.0.mlp.up_proj.weight_scale"]
    with pytest.raises(ValueError, match="Expected key"):
        dq.validate(t)
    t = _ct_tensors()
    t["model.layers.0.mlp.up_proj.extra_param"] = torch.empty(64, device="meta")
    with pytest.raises(ValueError, match="unconsumed keys"):
        dq.validate(t)
    assert dq.get_dependencies("model.layers.0.mlp.up_proj.weight") == {"model.layers.0.mlp.up_proj.weight_scale"}
    assert dq.get_dependencies("model.layers.0.mlp.up_proj.weight_scale") == set()
    assert dq.get_dependencies("model.embed_tokens.weight") == set()
    assert dq.create_config() is None


def test_modelopt_nvfp4_renames_and_inverts():
    conv = ModelOptNvfp4Converter(targets=[r"re:.*proj$"], ignore=["lm_head"])
    t = {"m.q_proj.weight": torch.zeros(4, 8, dtype=torch.uint8), "m.q_proj.weight_scale": torch.ones(4, 1).to(torch.float8_e4m3fn),
         "m.q_proj.weight_scale_2": torch.tensor(0.25), "m.q_proj.input_scale": torch.tensor(4.0), "m.norm.weight": torch.ones(4)}
    conv.validate(t)
    out = conv.process(dict(t))
    assert set(out) == {"m.q_proj.weight_packed", "m.q_proj.weight_scale", "m.q_proj.weight_global_scale", "m.q_proj.input_global_scale", "m.norm.weight"}
    assert out["m.q_proj.weight_global_scale"].item() == 4.0 and out["m.q_proj.input_global_scale"].item() == 0.25
    assert conv.get_dependencies("m.q_proj.weight") == {"m.q_proj.input_scale", "m.q_proj.weight_scale", "m.q_proj.weight_scale_2"}
    cfg = conv.create_config()
    assert cfg.format == CompressionFormat.nvfp4_pack_quantized.value and cfg.config_groups["config_group_0"].weights.group_size == 16
    with pytest.raises(ValueError, match="unexpected non-targeted"):
        conv.validate({"m.other.weight_scale_2": torch.ones(1)})

"""
Checkpoint converters on the B200: the one-pass AWQ repack and FP8-block dequantize kernels against
golden vectors produced by the reference and against the CPU oracle, the converters' process() on device tensors, and
convert_checkpoint end to end on small on-disk checkpoints with a thread pool.
"""
import json

import pytest
import torch
from safetensors import safe_open
from safetensors.torch import save_file

import oracle
from compressed_tensors_b200 import ops
from compressed_tensors_b200.config import CompressionFormat
from compressed_tensors_b200.entrypoints.convert import (AutoAWQConverter, CompressedTensorsDequantizer, FP8BlockDequantizer, convert_checkpoint)
from compressed_tensors_b200.quantization import QuantizationArgs, QuantizationConfig, QuantizationScheme
from tests.golden import load
from tests.util import same

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = load("convert")


@pytest.mark.parametrize("i", range(len(G["awq"])))
def test_awq_repack_golden_and_converter(i):
    c = G["awq"][i]
    want = c["result"]
    same(ops.awq_repack(c["qweight"].to(DEV)).cpu(), want["m.q_proj.weight_packed"], "awq_repack")
    if c["zero_point"]:
        same(ops.awq_repack_zeros(c["qzeros"].to(DEV)).cpu(), want["m.q_proj.weight_zero_point"], "awq_repack_zeros")
    conv = AutoAWQConverter(group_size=c["group_size"], zero_point=c["zero_point"])
    t = {"m.q_proj.qweight": c["qweight"].to(DEV), "m.q_proj.scales": c["scales"].to(DEV), "m.embed.weight": torch.ones(2, 2, device=DEV)}
    if c["zero_point"]:
        t["m.q_proj.qzeros"] = c["qzeros"].to(DEV)
    conv.validate(t)
    got = conv.process(t)
    assert set(got) == set(want) | {"m.embed.weight"}
    for k, v in want.items():
        same(got[k].cpu(), v, f"AutoAWQConverter.process[{k}]")
        assert got[k].is_contiguous()


@pytest.mark.parametrize("k,n", [(4096, 4096), (4096, 11008), (1000, 264), (72, 8)])
def test_awq_repack_vs_oracle(k, n):
    g = torch.Generator().manual_seed(k + n)
    q = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n // 8), generator=g, dtype=torch.int64).to(torch.int32)
    same(ops.awq_repack(q.to(DEV)).cpu(), oracle.awq_repack(q), "awq_repack")
    z = torch.randint(-2 ** 31, 2 ** 31 - 1, (max(k // 128, 1), n // 8), generator=g, dtype=torch.int64).to(torch.int32)
    same(ops.awq_repack_zeros(z.to(DEV)).cpu(), oracle.awq_repack_zeros(z), "awq_repack_zeros")
    same(ops.awq_repack(q), oracle.awq_repack(q), "awq_repack (cpu tensor in)")


def test_awq_repack_then_decompress_equals_awq_dequantization():
    """size-independent property: converting and then decompressing gives (code - zero) * scale of the AWQ checkpoint"""
    from compressed_tensors_b200.compressors import PackedQuantizationCompressor

    g = torch.Generator().manual_seed(1)
    k, n, gs = 512, 1024, 128
    codes = torch.randint(0, 16, (k, n), generator=g)
    zeros = torch.randint(0, 16, (k // gs, n), generator=g)
    order = [0, 2, 4, 6, 1, 3, 5, 7]   # AutoAWQ's packing order

    def pack(v):
        out = torch.zeros(v.shape[0], v.shape[1] // 8, dtype=torch.int64)
        for j, o in enumerate(order):
            out |= v[:, o::8].to(torch.int64) << (4 * j)
        return out.to(torch.int32)   # wraps like the int32 storage

    scales = (torch.rand(k // gs, n, generator=g) * 0.01 + 0.001).to(torch.float16)
    conv = AutoAWQConverter(group_size=gs)
    t = conv.process({"m.q.qweight": pack(codes).to(DEV), "m.q.qzeros": pack(zeros).to(DEV), "m.q.scales": scales.to(DEV)})
    scheme = conv.create_config().config_groups["config_group_0"]
    state = {p: t[f"m.q.{p}"] for p in ("weight_packed", "weight_scale", "weight_zero_point", "weight_shape")}
    w = PackedQuantizationCompressor.decompress(state, scheme)["weight"]
    want = ((codes - zeros.repeat_interleave(gs, 0)).to(torch.float16) * scales.repeat_interleave(gs, 0)).T
    assert w.shape == (n, k) and torch.equal(w.cpu(), want.contiguous())


@pytest.mark.parametrize("i", range(len(G["fp8block"])))
def test_fp8_block_dequant_golden(i):
    c = G["fp8block"][i]
    conv = FP8BlockDequantizer(weight_block_size=c["block"], dtype=c["out"].dtype)
    same(conv._create_dequantized_weight(c["weight"].to(DEV), c["scale_inv"].to(DEV)).cpu(), c["out"], "FP8BlockDequantizer")


@pytest.mark.parametrize("shape,block", [((4096, 7168), (128, 128)), ((2112, 7168), (128, 128)), ((200, 300), (128, 128)), ((129, 257), (64, 64))])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_fp8_block_dequant_vs_oracle(shape, block, dt):
    g = torch.Generator().manual_seed(shape[0])
    w = (torch.randn(shape, generator=g) * 2).to(torch.float8_e4m3fn)
    s = torch.rand(-(-shape[0] // block[0]), -(-shape[1] // block[1]), generator=g) * 0.02 + 1e-4
    same(ops.dequantize_block_fp8(w.to(DEV), s.to(DEV), block, dt).cpu(), oracle.dequantize_block_fp8(w, s, block, dt), "dequantize_block_fp8")


def test_ct_dequantizer_process_on_device():
    dq = object.__new__(CompressedTensorsDequantizer)
    dq.dtype = torch.bfloat16
    scheme = QuantizationScheme(targets=["re:.*mlp.*"], weights=QuantizationArgs(num_bits=8, type="int", strategy="channel", symmetric=True, dynamic=False),
                                format=CompressionFormat.naive_quantized)
    dq.quant_config = QuantizationConfig(config_groups={"group_0": scheme}, ignore=["model.embed_tokens"])
    t = {"model.layers.0.mlp.up_proj.weight": torch.randint(-128, 127, (64, 64), dtype=torch.int8, device=DEV),
         "model.layers.0.mlp.up_proj.weight_scale": torch.rand(64, 1, device=DEV),
         "model.layers.0.mlp.down_proj.weight": torch.randint(-128, 127, (64, 64), dtype=torch.int8, device=DEV),
         "model.layers.0.mlp.down_proj.weight_scale": torch.rand(64, 1, device=DEV),
         "model.layers.0.input_layernorm.weight": torch.randn(64, 1, dtype=torch.bfloat16, device=DEV),
         "model.layers.0.self_attn.q_proj.weight": torch.randn(128, 64, dtype=torch.bfloat16, device=DEV),
         "model.layers.0.self_attn.k_scale": torch.ones(1, device=DEV),
         "model.embed_tokens.weight": torch.randn(128, 64, dtype=torch.bfloat16, device=DEV)}
    keep = {k: v.clone() for k, v in t.items()}
    out = dq.process(t)
    for proj in ("up_proj", "down_proj"):
        w = out[f"model.layers.0.mlp.{proj}.weight"]
        want = (keep[f"model.layers.0.mlp.{proj}.weight"].float() * keep[f"model.layers.0.mlp.{proj}.weight_scale"]).to(torch.bfloat16)
        assert w.dtype == torch.bfloat16 and torch.equal(w, want)
        assert f"model.layers.0.mlp.{proj}.weight_scale" not in out
    assert torch.equal(out["model.layers.0.self_attn.q_proj.weight"], keep["model.layers.0.self_attn.q_proj.weight"])
    assert torch.equal(out["model.embed_tokens.weight"], keep["model.embed_tokens.weight"])
    assert "model.layers.0.self_attn.k_scale" not in out, "kv-cache qparams are dropped"


def _read_all(directory):
    out = {}
    for p in sorted(directory.glob("*.safetensors")):
        with safe_open(str(p), framework="pt") as f:
            for k in f.keys():
                out[k] = f.get_tensor(k)
    return out


@pytest.mark.parametrize("workers", [1, 3])
def test_convert_checkpoint_fp8_block_end_to_end(tmp_path, workers):
    src, dst = tmp_path / "src", tmp_path / "dst"
    src.mkdir()
    g = torch.Generator().manual_seed(0)
    wm, ref = {}, {}
    shards = {f"model-0000{s + 1}-of-00003.safetensors": {} for s in range(3)}
    for s in range(3):
        name = f"model-0000{s + 1}-of-00003.safetensors"
        for l in range(2):
            w = (torch.randn(256, 384, generator=g) * 2).to(torch.float8_e4m3fn)
            sc = torch.rand(2, 3, generator=g) * 0.02 + 1e-3
            shards[name][f"model.layers.{s}.mlp.p{l}_proj.weight"] = w
            # the partner scale lives in the NEXT shard: the planner has to bring it over
            ref[f"model.layers.{s}.mlp.p{l}_proj.weight"] = oracle.dequantize_block_fp8(w, sc, (128, 128), torch.bfloat16)
            wm[f"model.layers.{s}.mlp.p{l}_proj.weight"] = name
            other = f"model-0000{(s + 1) % 3 + 1}-of-00003.safetensors"
            shards[other][f"model.layers.{s}.mlp.p{l}_proj.weight_scale_inv"] = sc
            wm[f"model.layers.{s}.mlp.p{l}_proj.weight_scale_inv"] = other
        shards[name][f"model.layers.{s}.norm.weight"] = torch.randn(384, generator=g).to(torch.bfloat16)
        ref[f"model.layers.{s}.norm.weight"] = shards[name][f"model.layers.{s}.norm.weight"]
        wm[f"model.layers.{s}.norm.weight"] = name
    for name, t in shards.items():
        save_file(t, str(src / name))
    (src / "model.safetensors.index.json").write_text(json.dumps({"metadata": {"total_size": 0}, "weight_map": wm}))
    (src / "config.json").write_text(json.dumps({"model_type": "test", "quantization_config": {"quant_method": "fp8", "weight_block_size": [128, 128]}}))
    (src / "tokenizer.json").write_text("{}")

    convert_checkpoint(src, dst, FP8BlockDequantizer(targets=[r"re:.*proj$"]), max_workers=workers, device=DEV)

    got = _read_all(dst)
    assert set(got) == set(ref)
    for k, v in ref.items():
        same(got[k], v, f"converted {k}")
    cfg = json.loads((dst / "config.json").read_text())
    assert "quantization_config" not in cfg and (dst / "tokenizer.json").exists()
    idx = json.loads((dst / "model.safetensors.index.json").read_text())
    assert set(idx["weight_map"]) == set(ref) and idx["metadata"]["total_size"] == sum(v.nbytes for v in ref.values())


def test_convert_checkpoint_awq_then_dequantize_round_trip(tmp_path):
    """AutoAWQ checkpoint -> compressed-tensors (AutoAWQConverter) -> dense (CompressedTensorsDequantizer), two passes over disk"""
    src, mid, dst = tmp_path / "awq", tmp_path / "ct", tmp_path / "dense"
    src.mkdir()
    g = torch.Generator().manual_seed(5)
    k, n, gs = 256, 512, 128
    t, dense = {}, {}
    for name in ("model.layers.0.mlp.up_proj", "model.layers.0.mlp.down_proj"):
        q = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n // 8), generator=g, dtype=torch.int64).to(torch.int32)
        z = torch.randint(-2 ** 31, 2 ** 31 - 1, (k // gs, n // 8), generator=g, dtype=torch.int64).to(torch.int32)
        s = (torch.rand(k // gs, n, generator=g) * 0.01 + 1e-3).to(torch.float16)
        t[f"{name}.qweight"], t[f"{name}.qzeros"], t[f"{name}.scales"] = q, z, s
        iw, iz = AutoAWQConverter.unpack_awq(q, z, 4)
        iw, iz = AutoAWQConverter.reverse_awq_order(iw, iz, 4)
        dense[f"{name}.weight"] = (((iw & 15) - (iz & 15).repeat_interleave(gs, 0)).to(torch.float16) * s.repeat_interleave(gs, 0)).T.contiguous()
    t["model.embed_tokens.weight"] = torch.randn(64, 32, generator=g).to(torch.float16)
    dense["model.embed_tokens.weight"] = t["model.embed_tokens.weight"]
    save_file(t, str(src / "model.safetensors"))
    (src / "config.json").write_text(json.dumps({"model_type": "test", "quantization_config": {"quant_method": "awq", "bits": 4, "group_size": gs, "zero_point": True, "version": "gemm"}}))

    conv = AutoAWQConverter.from_autoawq_config({"bits": 4, "group_size": gs, "zero_point": True, "version": "gemm"})
    convert_checkpoint(src, mid, conv, max_workers=1, device=DEV)
    cfg = json.loads((mid / "config.json").read_text())
    assert cfg["quantization_config"]["format"] == "pack-quantized" and cfg["quantization_config"]["quantization_status"] == "compressed"
    ct = _read_all(mid)
    assert ct["model.layers.0.mlp.up_proj.weight_packed"].shape == (n, k // 8) and ct["model.layers.0.mlp.up_proj.weight_shape"].tolist() == [n, k]

    convert_checkpoint(mid, dst, CompressedTensorsDequantizer(mid, dtype=torch.float16), max_workers=2, device=DEV)
    got = _read_all(dst)
    assert set(got) == set(dense)
    for name, want in dense.items():
        assert got[name].dtype == torch.float16 and torch.equal(got[name], want), name
    assert "quantization_config" not in json.loads((dst / "config.json").read_text())

"""
FP4 (E2M1) / MX ops on the B200 through the C ABI: golden vectors produced by the reference
(tests/golden/fp4.pt.gz), the CPU oracle on larger seeded inputs, and size-independent properties.
Bit-exact (sign of zero included); NaN payloads are not compared.
"""
import pytest
import torch

import oracle
from compressed_tensors_b200 import ops
from compressed_tensors_b200.quantization import QuantizationArgs
from tests.golden import load
from tests.util import same, same_nan

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = load("fp4")
DTS = (torch.bfloat16, torch.float16, torch.float32)
FP4 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
NV = dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16)
MX4 = dict(num_bits=4, type="float", symmetric=True, strategy="group", group_size=32)


def qa(d):
    return QuantizationArgs(**{k: v for k, v in d.items() if k in ("num_bits", "type", "symmetric", "strategy", "group_size", "block_structure")})


def okw(a):
    return dict(strategy=a.strategy, group_size=a.group_size, block_structure=a.block_structure, num_bits=a.num_bits, qtype=a.type)


@pytest.mark.parametrize("i", range(len(G["cast"])))
def test_cast_to_fp4_golden(i):
    c = G["cast"][i]
    same_nan(ops.cast_to_fp4(c["x"].to(DEV)).cpu(), c["y"], "cast_to_fp4")
    same_nan(ops.cast_to_fp4(c["x"]), c["y"], "cast_to_fp4 (cpu tensor in)")


@pytest.mark.parametrize("i", range(len(G["pack"])))
def test_pack_unpack_fp4_golden(i):
    c = G["pack"][i]
    same(ops.pack_fp4_to_uint8(c["x"].to(DEV)).cpu(), c["packed"], "pack_fp4_to_uint8")
    m, n = c["x"].shape
    for name, want in c["unpacked"].items():
        dt = getattr(torch, name.split(".")[1])
        same(ops.unpack_fp4_from_uint8(c["packed"].to(DEV), m, n, dt).cpu(), want, f"unpack -> {name}")


def test_pack_fp4_errors():
    with pytest.raises(ValueError):
        ops.pack_fp4_to_uint8(torch.zeros(3, 7, device=DEV))


@pytest.mark.parametrize("i", range(len(G["nvfp4"])))
def test_nvfp4_golden(i):
    c = G["nvfp4"][i]
    a, gs = qa(c["args"]), c["global_scale"].to(DEV)
    x, s, zp = c["x"].to(DEV), c["scale"].to(DEV), c["qparams_zp"].to(DEV)
    same(ops.quantize(x, s, zp, a, global_scale=gs).cpu(), c["q"], "nvfp4 quantize")
    same(ops.fake_quantize(x, s, zp, a, global_scale=gs).cpu(), c["fq"], "nvfp4 fake_quantize")
    same(ops.dequantize(c["q"].to(DEV), s, global_scale=gs, dtype=c["x"].dtype).cpu(), c["dq"], "nvfp4 dequantize")
    # fused: quantize + nibble pack == pack(quantize); unpack + dequantize == dequantize(unpack)
    packed = ops.quantize_pack_fp4(x, s, zp, a, global_scale=gs)
    same(packed.cpu(), oracle.pack_fp4_to_uint8(c["q"]), "quantize_pack_fp4")
    back = ops.unpack_dequantize_fp4(packed, s.to(torch.bfloat16), gs, dtype=torch.bfloat16)
    want = oracle.dequantize(oracle.unpack_fp4_from_uint8(packed.cpu(), *c["x"].shape, torch.bfloat16), c["scale"].to(torch.bfloat16), None,
                             global_scale=c["global_scale"], dtype=torch.bfloat16)
    same(back.cpu(), want, "unpack_dequantize_fp4")


@pytest.mark.parametrize("i", range(len(G["mx"])))
def test_mx_golden(i):
    c = G["mx"][i]
    a = qa(c["args"])
    x, s, zp = c["x"].to(DEV), c["scale"].to(DEV), c["qparams_zp"].to(DEV)
    dt = torch.float8_e4m3fn if a.num_bits == 8 else None
    same(ops.quantize(x, s, zp, a, dtype=dt).cpu(), c["q"], "mx quantize")
    same(ops.fake_quantize(x, s, zp, a).cpu(), c["fq"], "mx fake_quantize")
    same(ops.dequantize(c["q"].to(DEV), s, dtype=c["x"].dtype).cpu(), c["dq"], "mx dequantize")


@pytest.mark.parametrize("i", range(len(G["e8m0"])))
def test_e8m0_golden(i):
    c = G["e8m0"][i]
    same(ops.compress_mx_scale(c["scale"].to(DEV)).cpu(), c["enc"], "compress_mx_scale")
    same(ops.decompress_mx_scale(c["enc"].to(DEV)).cpu(), c["dec"], "decompress_mx_scale")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("scale_dt", [torch.float32, None])
@pytest.mark.parametrize("shape", [(64, 256), (37, 48), (128, 2048)])
def test_nvfp4_vs_oracle(dt, scale_dt, shape):
    g = torch.Generator().manual_seed(hash((str(dt), shape)) % 1000)
    x = (torch.randn(shape, generator=g) * torch.exp2(torch.randint(-6, 3, (shape[0], 1), generator=g).float())).to(dt)
    gs = (448.0 * 6.0 / x.float().abs().max()).reshape(1)
    s = (x.float().unflatten(-1, (-1, 16)).abs().amax(-1) / 6.0 * gs).to(torch.float8_e4m3fn).to(scale_dt or dt)
    s = torch.where(s == 0, torch.tensor(2.0 ** -9, dtype=s.dtype), s)
    a = QuantizationArgs(**NV)
    same(ops.quantize(x.to(DEV), s.to(DEV), None, a, global_scale=gs.to(DEV)).cpu(), oracle.quantize(x, s, None, global_scale=gs, **okw(a)), "quantize")
    same(ops.fake_quantize(x.to(DEV), s.to(DEV), None, a, global_scale=gs.to(DEV)).cpu(), oracle.fake_quantize(x, s, None, global_scale=gs, **okw(a)), "fake_quantize")
    packed = ops.quantize_pack_fp4(x.to(DEV), s.to(DEV), None, a, global_scale=gs.to(DEV))
    same(packed.cpu(), oracle.pack_fp4_to_uint8(oracle.quantize(x, s, None, global_scale=gs, **okw(a))), "quantize_pack_fp4")
    # decompress from the STORED fp8 scale == the reference's decompress (scale.to(bf16), bf16 output)
    s8 = s.to(torch.float8_e4m3fn)
    want = oracle.dequantize(oracle.unpack_fp4_from_uint8(packed.cpu(), *shape, torch.bfloat16), s8.to(torch.bfloat16), None, global_scale=gs, dtype=torch.bfloat16)
    same(ops.unpack_dequantize_fp4(packed, s8.to(DEV), gs.to(DEV), stored_scale="fp8").cpu(), want, "unpack_dequantize_fp4 (stored fp8 scale)")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("shape", [(64, 256), (5, 96), (128, 4096)])
def test_mxfp4_vs_oracle(dt, shape):
    g = torch.Generator().manual_seed(shape[1])
    x = (torch.randn(shape, generator=g) * 0.3).to(dt)
    e = torch.floor(torch.log2(x.float().unflatten(-1, (-1, 32)).abs().amax(-1).clamp_min(1e-20))) - 2
    s = torch.exp2(e).to(dt)
    a = QuantizationArgs(**MX4)
    q = oracle.quantize(x, s, None, **okw(a))
    same(ops.quantize(x.to(DEV), s.to(DEV), None, a).cpu(), q, "quantize")
    packed = ops.quantize_pack_fp4(x.to(DEV), s.to(DEV), None, a)
    same(packed.cpu(), oracle.pack_fp4_to_uint8(q), "quantize_pack_fp4")
    enc = ops.compress_mx_scale(s.to(DEV))
    same(enc.cpu(), oracle.compress_mx_scale(s), "compress_mx_scale")
    sb = oracle.decompress_mx_scale(enc.cpu())
    want = oracle.dequantize(oracle.unpack_fp4_from_uint8(packed.cpu(), *shape, torch.bfloat16), sb, None, dtype=torch.bfloat16)
    same(ops.unpack_dequantize_fp4(packed, enc, None, stored_scale="e8m0").cpu(), want, "unpack_dequantize_fp4 (stored e8m0 scale)")
    same(ops.unpack_dequantize_fp4(packed, sb.to(DEV)).cpu(), want, "unpack_dequantize_fp4 (float scale)")


def test_fp4_pack_round_trip_full_size():
    """size-independent property at a Llama-3-8B shape: unpack(pack(v)) == v for every valid fp4 value incl. -0.0"""
    idx = torch.randint(0, 16, (4096, 14336), device=DEV)
    v = torch.where(idx >= 8, -FP4.to(DEV)[idx % 8], FP4.to(DEV)[idx % 8]).to(torch.bfloat16)
    p = ops.pack_fp4_to_uint8(v)
    assert p.shape == (4096, 7168) and p.dtype == torch.uint8
    back = ops.unpack_fp4_from_uint8(p, 4096, 14336, torch.bfloat16)
    assert torch.equal(back.view(torch.int16), v.view(torch.int16))
    # idempotence of the cast on its own outputs
    assert torch.equal(ops.cast_to_fp4(v.abs()).view(torch.int16), v.abs().view(torch.int16))

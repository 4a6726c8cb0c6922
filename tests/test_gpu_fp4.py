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


# ---- compressors: state dict in -> state dict out, against the reference's own outputs --------------------------
def _cls(fmt):
    from compressed_tensors_b200.compressors import MXFP4PackedCompressor, MXFP8QuantizationCompressor, NVFP4PackedCompressor

    return {"nvfp4": NVFP4PackedCompressor, "mxfp4": MXFP4PackedCompressor, "mxfp8": MXFP8QuantizationCompressor}[fmt]


def _same_state(got, want, what):
    got = {k: v for k, v in got.items() if v is not None}
    assert set(got) == set(want), (what, sorted(got), sorted(want))
    for k, w in want.items():
        g = got[k].data if isinstance(got[k], torch.nn.Parameter) else got[k]
        if w.dtype == torch.float8_e4m3fn:
            g, w = g.view(torch.uint8), w.view(torch.uint8)
        same(g.cpu(), w, f"{what}[{k}]")


@pytest.mark.parametrize("i", range(len(G["compressors"])))
@pytest.mark.parametrize("where", [DEV, "cpu"])
def test_fp4_mx_compressors_golden(i, where):
    from compressed_tensors_b200.quantization import QuantizationScheme

    c = G["compressors"][i]
    cls = _cls(c["format"])
    extra = {"zp_dtype": torch.uint8} if c["format"].startswith("mx") else {}
    scheme = QuantizationScheme(targets=["Linear"], weights=QuantizationArgs(**{**c["args"], **extra}))
    state = {k: v.to(where) for k, v in c["state"].items()}
    comp = cls.compress(state, scheme)
    assert set(state) == set(c["state"]), "input must not be mutated"
    _same_state(comp, c["compressed"], f"{c['format']} compress")
    assert all(v.device.type == torch.device(where).type for v in comp.values() if v is not None)
    back = cls.decompress(comp, scheme)
    _same_state(back, c["decompressed"], f"{c['format']} decompress")


# ---- the reference's own tests for these formats (tests/test_compressors/test_fp4_quant.py, test_fp4_optimizations.py,
# ---- test_mxfp4_quant.py, test_mxfp8_quant.py), re-expressed against this package -------------------------------------
def test_ref_pack_unpack_preserves_sign_of_zero():
    x = torch.tensor([[-0.5, -6.0, -0.5, -1.5, -1.0, 6.0, 0.0, -0.0], [-1.0, -6.0, -0.5, -0.0, 0.5, 0.5, -0.0, 0.0],
                      [-3.0, -6.0, -0.5, -2.0, -0.5, -1.5, -0.0, -0.0], [1.5, 6.0, -0.0, -0.5, 1.0, 1.0, -0.0, 0.0]], dtype=torch.bfloat16, device=DEV)
    packed = ops.pack_fp4_to_uint8(x)
    assert packed.dtype == torch.uint8
    back = ops.unpack_fp4_from_uint8(packed, *x.shape, dtype=torch.bfloat16)
    assert back.dtype == torch.bfloat16 and torch.equal(back, x) and torch.equal(torch.signbit(back), torch.signbit(x))


@pytest.mark.parametrize("x", [torch.tensor([[0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]]), torch.tensor([[-0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0, -0.0]]),
                               torch.tensor([[0.0, -0.5, 1.0, -1.5, 2.0, -3.0, 4.0, -6.0]])])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_ref_pack_matches_nearest_index_search(x, dt):
    x = x.to(dtype=dt, device=DEV)
    table = FP4.to(device=DEV, dtype=dt)
    idx = torch.argmin(torch.abs(x.abs().unsqueeze(-1) - table), dim=-1).to(torch.int8) + (torch.signbit(x).to(torch.int8) << 3)
    idx = idx.reshape(-1, 2)
    want = (idx[:, 0].to(torch.uint8) | (idx[:, 1].to(torch.uint8) << 4)).reshape(x.shape[0], x.shape[1] // 2)
    assert torch.equal(ops.pack_fp4_to_uint8(x), want)


def test_ref_pack_non_contiguous():
    base = torch.tensor([[0.0, -0.5, 1.0, -1.5], [2.0, -3.0, 4.0, -6.0], [0.5, -1.0, 1.5, -2.0], [3.0, -4.0, 6.0, -0.0]], dtype=torch.bfloat16, device=DEV)
    x = base[:, ::2]
    assert not x.is_contiguous()
    assert torch.equal(ops.pack_fp4_to_uint8(x), ops.pack_fp4_to_uint8(x.contiguous()))


def test_ref_compress_scale_defaults():
    from compressed_tensors_b200.compressors import MXFP4PackedCompressor, MXFP8QuantizationCompressor, NVFP4PackedCompressor

    s = torch.randn(10, dtype=torch.bfloat16, device=DEV).abs() + 1e-6
    assert NVFP4PackedCompressor._compress_scale(s, QuantizationArgs(num_bits=4, type="float", symmetric=True, group_size=16)).dtype == torch.float8_e4m3fn
    assert MXFP4PackedCompressor._compress_scale(s, QuantizationArgs(num_bits=4, type="float", symmetric=True, group_size=32)).dtype == torch.uint8
    assert MXFP4PackedCompressor._compress_scale(s, QuantizationArgs(num_bits=4, type="float", symmetric=True, group_size=32, scale_dtype=torch.uint8)).dtype == torch.uint8
    assert MXFP8QuantizationCompressor._compress_scale(s, QuantizationArgs(num_bits=8, type="float", symmetric=True, group_size=32)).dtype == torch.uint8


def test_ref_mxfp4_decompress_decodes_scales_and_restores_weight():
    from compressed_tensors_b200.compressors import MXFP4PackedCompressor
    from compressed_tensors_b200.quantization import QuantizationScheme

    a = QuantizationArgs(num_bits=4, type="float", symmetric=True, group_size=32, scale_dtype=torch.uint8)
    scale = torch.tensor([[0.25, 0.5]], dtype=torch.bfloat16, device=DEV)
    packed = ops.pack_fp4_to_uint8(torch.tensor([[0.5, 1.0, 1.5, 2.0]], dtype=torch.bfloat16, device=DEV))
    out = MXFP4PackedCompressor.decompress({"weight_packed": packed, "weight_scale": MXFP4PackedCompressor._compress_scale(scale, a)},
                                           QuantizationScheme(targets=["Linear"], weights=a))
    assert torch.equal(out["weight_scale"], scale)
    assert torch.equal(out["weight"], torch.tensor([[0.125, 0.25, 0.75, 1.0]], dtype=torch.bfloat16, device=DEV))


def test_ref_mxfp8_compress_decompress_and_scale_round_trip():
    from compressed_tensors_b200.compressors import MXFP8QuantizationCompressor
    from compressed_tensors_b200.quantization import QuantizationScheme
    from compressed_tensors_b200.quantization.utils import calculate_qparams

    a = QuantizationArgs(num_bits=8, type="float", strategy="group", group_size=32, scale_dtype=torch.uint8, zp_dtype=torch.uint8, symmetric=True)
    w = torch.randn((512, 1024), device=DEV)
    g = w.reshape(512, 32, 32)
    scale, zp = calculate_qparams(g.amin(-1), g.amax(-1), a)
    scheme = QuantizationScheme(targets=["Linear"], weights=a)
    comp = MXFP8QuantizationCompressor.compress({"weight": w, "weight_scale": scale, "weight_zero_point": zp}, scheme)
    assert comp["weight"].dtype == torch.float8_e4m3fn and comp["weight_scale"].dtype == torch.uint8
    decoded = 2.0 ** (comp["weight_scale"].to(torch.int32) - 127).to(torch.float32)
    assert torch.allclose(decoded, 2.0 ** torch.floor(torch.log2(scale)).to(torch.float32))
    back = MXFP8QuantizationCompressor.decompress(comp, scheme)
    assert back["weight"].shape == w.shape and torch.allclose(back["weight"].float(), w, atol=0.1, rtol=0.1)


def test_nvfp4_model_compressor_round_trip():
    """ModelCompressor over a small NVFP4A16 model: compress == per-module reference flow, decompress gives bf16 weights back"""
    from compressed_tensors_b200.compressors import ModelCompressor
    from compressed_tensors_b200.quantization import QuantizationConfig, QuantizationStatus, apply_quantization_config
    from compressed_tensors_b200.quantization.utils import calculate_qparams, generate_gparam

    model = torch.nn.Sequential(torch.nn.Linear(256, 128, bias=False), torch.nn.Linear(128, 64, bias=False)).to(DEV).to(torch.bfloat16)
    apply_quantization_config(model, QuantizationConfig(config_groups={"NVFP4A16": ["Linear"]}))
    dense = []
    for lin in model:
        a = lin.quantization_scheme.weights
        w = lin.weight.data
        gs = generate_gparam(w.min(), w.max())
        g = w.unflatten(-1, (-1, 16))
        s, z = calculate_qparams(g.amin(-1), g.amax(-1), a, global_scale=gs)
        lin.weight_scale = torch.nn.Parameter(s, requires_grad=False)
        lin.weight_global_scale = torch.nn.Parameter(gs, requires_grad=False)
        lin.quantization_status = QuantizationStatus.FROZEN
        dense.append(ops.fake_quantize(w, s, None, a, global_scale=gs))
    mc = ModelCompressor.from_pretrained_model(model)
    assert mc.quantization_config.format == "nvfp4-pack-quantized"
    mc.compress_model(model)
    for lin in model:
        assert lin.weight_packed.dtype == torch.uint8 and lin.weight_scale.dtype == torch.float8_e4m3fn and not hasattr(lin, "weight")
    mc.decompress_model(model)
    for lin, want in zip(model, dense):
        assert lin.weight.dtype == torch.bfloat16 and lin.weight.shape == want.shape
        # fake_quantize used the float32 scale, decompress the fp8-stored one: equal because calculate_qparams already rounded it to fp8
        assert torch.equal(lin.weight.data, want)


# ---- streaming fast path specifics ---------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mode,exponent", [(0, e) for e in (-100, -60, -17, -1, 0, 1, 9)] + [(1, e) for e in (-60, -13, 0, 8, 14, 59)])
def test_fp4_division_shortcut_is_exact(dt, mode, exponent):
    """every 16-bit x, every float32 significand of the divisor: reciprocal + residual step == IEEE division
    (mode 0: the E2M1 code of weight / scale; mode 1: the float32 value of scale / global_scale)"""
    import ctypes

    from compressed_tensors_b200 import _native as N

    bad = ctypes.c_uint64(123)
    N.check(N.lib().ct_selftest_fp4_division(N.DT[dt], exponent, mode, ctypes.byref(bad), 0), "selftest")
    assert bad.value == 0, f"{bad.value} mismatching (x, divisor) pairs"


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("scale_dt", ["same", torch.float32])
@pytest.mark.parametrize("zp_mode", [None, "zeros", "values"])
def test_nvfp4_fast_path_zero_points_and_extremes(dt, scale_dt, zp_mode):
    g = torch.Generator().manual_seed(11)
    rows, cols = 96, 1024
    x = (torch.randn(rows, cols, generator=g) * torch.exp2(torch.randint(-12, 6, (rows, cols // 16, 1), generator=g).float()).expand(-1, -1, 16).reshape(rows, cols)).to(dt)
    x[0, :8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 65504.0, -65504.0, 1e-7, -1e-7]).to(dt)
    gs = torch.tensor([2688.0 / 7.0])
    s = (x.float().unflatten(-1, (-1, 16)).abs().amax(-1).clamp(1e-6, 3e4) / 6.0 * gs).clamp(max=448.0).to(torch.float8_e4m3fn).float().clamp_min(2.0 ** -9)
    assert not s.isnan().any()
    s[1, :4] = torch.tensor([2.0 ** -9, 448.0, 1.0, 3.5])
    s = s.to(dt if scale_dt == "same" else scale_dt)
    zp = None
    if zp_mode == "zeros":
        zp = torch.zeros(rows, cols // 16).to(torch.float8_e4m3fn)
    elif zp_mode == "values":
        zp = (torch.randint(-4, 5, (rows, cols // 16), generator=g).float() * 0.5).to(torch.float8_e4m3fn)
    a = QuantizationArgs(**NV)
    want = oracle.pack_fp4_to_uint8(oracle.quantize(x, s, zp, global_scale=gs, **okw(a)))
    got = ops.quantize_pack_fp4(x.to(DEV), s.to(DEV), zp.to(DEV) if zp is not None else None, a, global_scale=gs.to(DEV))
    same(got.cpu(), want, "quantize_pack_fp4 fast path")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("zp_dt", [None, torch.uint8, torch.int8])
@pytest.mark.parametrize("gsize", [32, 64])
def test_mxfp4_fast_path(dt, zp_dt, gsize):
    g = torch.Generator().manual_seed(5)
    rows, cols = 64, 2048
    x = (torch.randn(rows, cols, generator=g) * 0.7).to(dt)
    x[0, :4] = torch.tensor([0.0, -0.0, 1e-6, -1e-6]).to(dt)
    e = torch.floor(torch.log2(x.float().unflatten(-1, (-1, gsize)).abs().amax(-1).clamp_min(1e-20))) - 2
    s = torch.exp2(e).to(dt)
    s[2, :3] = torch.tensor([0.3, 1.7, 0.011]).to(dt)       # not powers of two: the exact-division machinery has to hold
    zp = torch.zeros(rows, cols // gsize, dtype=zp_dt) if zp_dt is not None else None
    if zp is not None:
        zp[3, :2] = 1
    a = QuantizationArgs(num_bits=4, type="float", symmetric=True, strategy="group", group_size=gsize)
    want = oracle.pack_fp4_to_uint8(oracle.quantize(x, s, zp, **okw(a)))
    got = ops.quantize_pack_fp4(x.to(DEV), s.to(DEV), zp.to(DEV) if zp is not None else None, a)
    same(got.cpu(), want, "quantize_pack_fp4 (T arithmetic) fast path")
    back = ops.unpack_dequantize_fp4(got, s.to(DEV), dtype=dt)
    same(back.cpu(), oracle.dequantize(oracle.unpack_fp4_from_uint8(want, rows, cols, dt), s, None, dtype=dt), "unpack_dequantize_fp4 (float scale)")


def test_fp4_full_size_properties():
    """Llama-3-8B down_proj shape: fused == unfused through the C ABI, decompress(compress(w)) == fake_quantize(w)"""
    rows, cols = 4096, 14336
    w = (torch.randn(rows, cols, device=DEV) * 0.02).to(torch.bfloat16)
    a = QuantizationArgs(**NV)
    gs = (448.0 * 6.0 / w.float().abs().max()).reshape(1)
    s = (w.float().unflatten(-1, (-1, 16)).abs().amax(-1) / 6.0 * gs).to(torch.float8_e4m3fn)
    s = torch.where(s.float() == 0, torch.tensor(0.125, device=DEV).to(torch.float8_e4m3fn), s)
    sb = s.to(torch.bfloat16)
    packed = ops.quantize_pack_fp4(w, sb, None, a, global_scale=gs)
    assert torch.equal(packed, ops.pack_fp4_to_uint8(ops.quantize(w, sb, None, a, global_scale=gs)))
    back = ops.unpack_dequantize_fp4(packed, s, gs, stored_scale="fp8")
    assert torch.equal(back, ops.fake_quantize(w, sb, None, a, global_scale=gs))
    assert torch.equal(back, ops.dequantize(ops.unpack_fp4_from_uint8(packed, rows, cols, torch.bfloat16), sb, global_scale=gs, dtype=torch.bfloat16))


@pytest.mark.parametrize("preset", ["NVFP4A16", "MXFP4A16", "MXFP8A16"])
@pytest.mark.parametrize("where", [DEV, "cpu"])
def test_model_compressor_batched_equals_per_module(preset, where):
    """one multi-tensor launch (ModelCompressor; the cross-tensor host pipeline for a CPU-resident model) == the per-module
    compressor calls, both directions"""
    import copy

    from compressed_tensors_b200.compressors import ModelCompressor, compress_module, decompress_module
    from compressed_tensors_b200.quantization import QuantizationConfig, QuantizationStatus, apply_quantization_config
    from compressed_tensors_b200.quantization.utils import calculate_qparams, generate_gparam
    from compressed_tensors_b200.utils import get_direct_state_dict

    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(512, 256, bias=False), torch.nn.Linear(256, 128, bias=False), torch.nn.Linear(128, 96, bias=False)).to(where).to(torch.bfloat16)
    apply_quantization_config(model, QuantizationConfig(config_groups={preset: ["Linear"]}))
    for lin in model:
        a = lin.quantization_scheme.weights
        w = lin.weight.data
        g = w.unflatten(-1, (-1, a.group_size))
        if preset.startswith("NV"):
            gs = generate_gparam(w.min(), w.max())
            lin.weight_global_scale.data.copy_(gs)
            s, _ = calculate_qparams(g.amin(-1), g.amax(-1), a, global_scale=gs)
        else:
            s, _ = calculate_qparams(g.amin(-1), g.amax(-1), a)
        lin.weight_scale.data.copy_(s)          # the parameter has the weight's dtype, as after calibration
        lin.quantization_status = QuantizationStatus.FROZEN
    single = copy.deepcopy(model)
    mc = ModelCompressor.from_pretrained_model(model)
    mc.compress_model(model)
    for lin in single:
        compress_module(lin)
    for a, b in zip(model, single):
        sa, sb = get_direct_state_dict(a), get_direct_state_dict(b)
        assert set(sa) == set(sb), (sorted(sa), sorted(sb))
        for k in sa:
            x, y = sa[k], sb[k]
            if x is None or y is None:
                assert x is None and y is None, k
                continue
            if x.dtype == torch.float8_e4m3fn:
                x, y = x.view(torch.uint8), y.view(torch.uint8)
            assert x.dtype == y.dtype and torch.equal(x, y), (preset, k)
            assert x.device.type == torch.device(where).type, (k, x.device)
    mc.decompress_model(model)
    for lin in single:
        decompress_module(lin)
    for a, b in zip(model, single):
        sa, sb = get_direct_state_dict(a), get_direct_state_dict(b)
        assert set(sa) == set(sb)
        for k in sa:
            if sa[k] is None or sb[k] is None:
                assert sa[k] is None and sb[k] is None, k
                continue
            assert sa[k].dtype == sb[k].dtype and torch.equal(sa[k].float(), sb[k].float()), (preset, k)


# ---- NVFP4 with the group observer fused in (ct_observe_quantize_pack_nvfp4) ---------------------------------------------------
@pytest.mark.parametrize("i", [i for i, c in enumerate(G["nvfp4"]) if c["scale"].dtype == torch.float32 and c["x"].dtype != torch.float32 and c["x"].shape[1] % 32 == 0])
def test_nvfp4_fused_observer_golden(i):
    """the reference's generate_gparam -> calculate_qparams -> quantize -> pack chain, from its own outputs"""
    c = G["nvfp4"][i]
    a = qa(c["args"])
    packed, scale, gs = ops.observe_quantize_pack_nvfp4(c["x"].to(DEV), a)
    same(gs.cpu(), c["global_scale"], "global scale")
    same(scale.cpu().view(torch.uint8), c["qparams_scale"].to(torch.float8_e4m3fn).view(torch.uint8), "fp8 group scales")
    same(packed.cpu(), oracle.pack_fp4_to_uint8(c["q"]), "nibbles")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(64, 256), (128, 4096), (33, 1056), (5, 48)])
def test_nvfp4_fused_observer_vs_unfused(dt, shape):
    from compressed_tensors_b200.quantization.utils import calculate_qparams, generate_gparam

    g = torch.Generator().manual_seed(shape[0])
    x = (torch.randn(shape, generator=g) * torch.exp2(torch.randint(-10, 3, (shape[0], shape[1] // 16, 1), generator=g).float()).expand(-1, -1, 16).reshape(shape)).to(dt)
    x[0, :16] = 0                       # dead group: scale falls back to 0.125
    x[-1, -16:] = x[-1, -16:].abs()     # one-sided group
    a = QuantizationArgs(**NV, scale_dtype=torch.float8_e4m3fn)
    xd = x.to(DEV)
    gs = generate_gparam(xd.min(), xd.max())
    grp = xd.unflatten(-1, (-1, 16))
    s, _ = calculate_qparams(grp.amin(-1), grp.amax(-1), a, global_scale=gs)
    want_packed = ops.quantize_pack_fp4(xd, s, None, a, global_scale=gs)
    packed, scale, gs2 = ops.observe_quantize_pack_nvfp4(xd, a)
    assert torch.equal(gs2, gs)
    assert torch.equal(scale.view(torch.uint8), s.to(torch.float8_e4m3fn).view(torch.uint8)), "group scales"
    assert torch.equal(packed, want_packed), "nibbles"
    # and against the CPU oracle end to end
    os_ = oracle.quantize(x, s.cpu(), None, global_scale=gs.cpu(), **okw(a))
    same(packed.cpu(), oracle.pack_fp4_to_uint8(os_), "nibbles vs oracle")

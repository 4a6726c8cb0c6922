"""
The reference's own hot-path tests, re-expressed against this package and run on the B200
(tests/test_compressors/test_pack_quant.py, test_fp8_quant.py, test_int_quant.py and the block cases of
tests/test_quantization/lifecycle/test_forward.py of the reference).  Same inputs shapes / dtypes /
assertions (`torch.equal`), tensors placed on the GPU; a few cases also go in as CPU tensors, the way
the reference's suite calls the API.
"""
import math

import pytest
import torch

from compressed_tensors_b200.compressors import (
    FloatQuantizationCompressor,
    IntQuantizationCompressor,
    PackedQuantizationCompressor,
)
from compressed_tensors_b200.compressors.pack_quantized.helpers import pack_to_int32, unpack_from_int32
from compressed_tensors_b200.quantization import QuantizationArgs, QuantizationScheme, QuantizationStrategy
from compressed_tensors_b200.quantization.lifecycle.forward import dequantize, fake_quantize, quantize

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(sd, where=DEV):
    return {k: (v.to(where) if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}


def scheme_of(**kw):
    return QuantizationScheme(targets=["Linear"], weights=QuantizationArgs(**kw))


def g_idx_for(columns, group_size):
    return (torch.arange(columns, dtype=torch.int) // group_size)[torch.randperm(columns)]


# ---- test_pack_quant.py ----------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(512, 1024), (830, 545), (342, 512), (256, 700)])
@pytest.mark.parametrize("where", [DEV, "cpu"])
def test_quant_format(shape, where):
    sd = dev({"weight": torch.rand(shape), "weight_scale": torch.tensor(0.01, dtype=torch.float32),
              "weight_zero_point": torch.tensor(0, dtype=torch.int8)}, where)
    out = PackedQuantizationCompressor.compress(sd, scheme=scheme_of(num_bits=4, symmetric=True))
    assert "weight" not in out and "weight_zero_point" not in out
    assert out["weight_packed"].dtype == torch.int32 and out["weight_packed"].device.type == torch.device(where).type
    assert out["weight_packed"].shape == (shape[0], math.ceil(shape[1] * 4 / 32))
    assert torch.equal(out["weight_shape"], torch.tensor(shape))
    assert out["weight_scale"].dtype == torch.float32


@pytest.mark.parametrize("value", [
    torch.tensor([[1, 2], [3, 4]]),
    torch.tensor([[1, 2, 3, 4, 5, 6, 7, 0], [-1, -2, -3, -4, -5, -6, -7, -8]]),
    (torch.rand((32, 100)) * 16 - 8),
])
def test_repack_4bit(value):
    value = value.to(torch.int8).to(DEV)
    assert torch.equal(value, unpack_from_int32(pack_to_int32(value, 4), 4, value.shape))


@pytest.mark.parametrize("value", [torch.tensor([[30, 40], [50, 60]]),
                                   torch.tensor([[10, 15, 20, 25, 30, 35, 40, 45], [-10, -20, -30, -40, -50, -60, -70, -80]]),
                                   (torch.rand((32, 100)) * 256 - 128)])
def test_repack_8bit(value):
    value = value.to(torch.int8).to(DEV)
    assert torch.equal(value, unpack_from_int32(pack_to_int32(value, 8), 8, value.shape))


@pytest.mark.parametrize("num_bits", range(1, 9))
@pytest.mark.parametrize("shape", [(1, 32), (4, 33), (16, 64), (3, 100), (8, 700), (5, 1024)])
def test_pack_unpack_round_trip(num_bits, shape):
    lo, hi = -(1 << (num_bits - 1)), (1 << (num_bits - 1))
    v = torch.randint(lo, hi, shape, dtype=torch.int8, device=DEV)
    p = pack_to_int32(v, num_bits)
    assert p.dtype == torch.int32 and p.shape == (shape[0], math.ceil(shape[1] * num_bits / 32))
    assert torch.equal(unpack_from_int32(p, num_bits, torch.Size(shape)), v)


@pytest.mark.parametrize("num_bits", range(1, 9))
def test_compress_decompress_match(num_bits):
    sd = dev({"weight": torch.rand((511, 350)), "weight_scale": torch.tensor(0.01, dtype=torch.float32),
              "weight_zero_point": torch.tensor(0, dtype=torch.int8)})
    scheme = scheme_of(num_bits=num_bits, symmetric=False)
    back = PackedQuantizationCompressor.decompress(PackedQuantizationCompressor.compress(sd.copy(), scheme=scheme), scheme=scheme)
    fq = fake_quantize(sd["weight"], scale=sd["weight_scale"], zero_point=sd["weight_zero_point"], args=scheme.weights)
    assert torch.equal(fq, back["weight"].to(torch.float32))


@pytest.mark.parametrize("strategy", [QuantizationStrategy.GROUP, QuantizationStrategy.CHANNEL])
def test_asymmetric_packed_support(strategy):
    shape, group = (1024, 1024), (128 if strategy == QuantizationStrategy.GROUP else None)
    qshape = (shape[0], shape[1] // group) if group else (shape[0], 1)
    sd = dev({"weight": torch.rand(shape), "weight_scale": torch.rand(qshape).to(torch.float32),
              "weight_zero_point": torch.rand(qshape).to(torch.int8)})
    out = PackedQuantizationCompressor.compress(sd, scheme=scheme_of(num_bits=4, strategy=strategy.value, symmetric=False, group_size=group))
    assert out["weight_packed"].dtype == torch.int32 and out["weight_zero_point"].dtype == torch.int32
    assert out["weight_packed"].shape == (shape[0], math.ceil(shape[1] / 8))
    assert out["weight_zero_point"].shape == (math.ceil(shape[0] / 8), qshape[1]) and out["weight_zero_point"].is_contiguous()
    assert torch.equal(out["weight_shape"], torch.tensor(shape))


@pytest.mark.parametrize("actorder", ["group", "weight", None])
def test_actorder_compress_decompress_match(actorder):
    shape, group = (512, 1024), 128
    sd = {"weight": torch.rand(shape), "weight_scale": torch.rand((shape[0], shape[1] // group)) * 0.01 + 1e-4,
          "weight_zero_point": torch.randint(-8, 8, (shape[0], shape[1] // group), dtype=torch.int8)}
    if actorder == "group":
        sd["weight_g_idx"] = g_idx_for(shape[1], group)
    sd = dev(sd)
    scheme = scheme_of(num_bits=4, strategy="group", group_size=group, actorder=actorder, symmetric=False)
    out = PackedQuantizationCompressor.compress(sd, scheme=scheme)
    back = PackedQuantizationCompressor.decompress(out, scheme=scheme)
    fq = fake_quantize(sd["weight"], scale=sd["weight_scale"], zero_point=sd["weight_zero_point"], g_idx=sd.get("weight_g_idx"), args=scheme.weights)
    assert torch.equal(fq, back["weight"])


@pytest.mark.parametrize("num_bits", [4, 8])
@pytest.mark.parametrize("rows,groups", [(1024, 8), (100, 3), (512, 1)])
def test_zero_point_pack_unpack_consistency(num_bits, rows, groups):
    lo, hi = -(1 << (num_bits - 1)), (1 << (num_bits - 1))
    zp = torch.randint(lo, hi, (rows, groups), dtype=torch.int8, device=DEV)
    packed = pack_to_int32(zp, num_bits, packed_dim=0)
    assert packed.shape == (math.ceil(rows * num_bits / 32), groups)
    assert torch.equal(unpack_from_int32(packed.contiguous(), num_bits, zp.shape, packed_dim=0), zp)


@pytest.mark.parametrize("num_bits", [3, 4, 8])
def test_pack_unpack_3d_matches_stacked_2d(num_bits):
    lo, hi = -(1 << (num_bits - 1)), (1 << (num_bits - 1))
    v = torch.randint(lo, hi, (4, 16, 96), dtype=torch.int8, device=DEV)
    p = pack_to_int32(v, num_bits)
    assert torch.equal(p, torch.stack([pack_to_int32(v[i], num_bits) for i in range(4)]))
    assert torch.equal(unpack_from_int32(p, num_bits, v.shape), v)


# ---- test_fp8_quant.py ---------------------------------------------------------------------------
@pytest.mark.parametrize("strategy,group_size,sc,zp", [
    (QuantizationStrategy.TENSOR, None, torch.tensor(0.01), torch.tensor(0)),
    (QuantizationStrategy.GROUP, 128, torch.rand((512, 8)) * 0.01, torch.zeros((512, 8), dtype=torch.int8)),
    (QuantizationStrategy.CHANNEL, None, torch.rand((512, 1)) * 0.01, torch.zeros((512, 1), dtype=torch.int8)),
])
def test_fp8_quant_format_and_match(strategy, group_size, sc, zp):
    sd = {"weight": torch.rand((512, 1024)), "weight_scale": sc.to(torch.float32) + 1e-5, "weight_zero_point": zp.to(torch.float32)}
    if group_size is not None:
        sd["weight_g_idx"] = g_idx_for(1024, group_size)
    sd = dev(sd)
    scheme = scheme_of(strategy=strategy, type="float", group_size=group_size)
    out = FloatQuantizationCompressor.compress(sd, scheme=scheme)
    assert "weight_zero_point" not in out and out["weight"].dtype == torch.float8_e4m3fn and out["weight"].shape == (512, 1024)
    assert torch.equal(out["weight_scale"], sd["weight_scale"])
    if group_size is not None:
        assert torch.equal(out["weight_g_idx"], sd["weight_g_idx"])
    back = FloatQuantizationCompressor.decompress(out, scheme=scheme)
    fq = fake_quantize(sd["weight"], scale=sd["weight_scale"], zero_point=None, g_idx=sd.get("weight_g_idx"), args=scheme.weights)
    assert torch.equal(fq, back["weight"].to(torch.float32))


@pytest.mark.parametrize("rows,cols", [(10944, 2048), (2048, 10944), (256, 256), (300, 400), (256, 300), (300, 256)])
def test_block_quant_compression_padding(rows, cols):
    nrb, ncb = math.ceil(rows / 128), math.ceil(cols / 128)
    sd = dev({"weight": torch.rand((rows, cols)), "weight_scale": torch.rand((nrb, ncb)) * 0.01 + 0.001,
              "weight_zero_point": torch.zeros((nrb, ncb))})
    out = FloatQuantizationCompressor.compress(sd, scheme=scheme_of(strategy=QuantizationStrategy.BLOCK, type="float", block_structure=[128, 128]))
    assert out["weight"].shape == (rows, cols) and out["weight"].dtype == torch.float8_e4m3fn
    assert out["weight_scale"].shape == (nrb, ncb)


# ---- test_int_quant.py ---------------------------------------------------------------------------
@pytest.mark.parametrize("strategy,symmetric,group_size,sc,zp", [
    (QuantizationStrategy.TENSOR, True, None, torch.tensor(0.01), torch.tensor(0)),
    (QuantizationStrategy.GROUP, True, 128, torch.rand((512, 8)) * 0.01, torch.zeros((512, 8), dtype=torch.int8)),
    (QuantizationStrategy.CHANNEL, False, None, torch.rand((512, 1)) * 0.01, ((torch.rand((512, 1)) - 0.5) * 127).to(torch.int8)),
])
def test_int_quant_format(strategy, symmetric, group_size, sc, zp):
    sd = dev({"weight": torch.rand((512, 1024)), "weight_scale": sc.to(torch.float32) + 1e-5, "weight_zero_point": zp.to(torch.int32)})
    out = IntQuantizationCompressor.compress(sd, scheme=scheme_of(strategy=strategy, group_size=group_size, symmetric=symmetric))
    if symmetric:
        assert "weight_zero_point" not in out
    else:
        assert out["weight_zero_point"].dtype == torch.int32
    assert out["weight"].dtype == torch.int8 and out["weight_scale"].dtype == torch.float32


@pytest.mark.parametrize("strategy,group_size,sc,zp", [
    (QuantizationStrategy.TENSOR, None, torch.tensor(0.01), torch.tensor(0)),
    (QuantizationStrategy.GROUP, 128, torch.rand((300, 8)) * 0.01, torch.zeros((300, 8), dtype=torch.int8)),
    (QuantizationStrategy.CHANNEL, None, torch.rand((300, 1)) * 0.01, torch.zeros((300, 1), dtype=torch.int8)),
])
def test_int_compress_decompress_match(strategy, group_size, sc, zp):
    sd = dev({"weight": torch.rand((300, 1024)), "weight_scale": sc.to(torch.float32) + 1e-5, "weight_zero_point": zp.to(torch.int32)})
    scheme = scheme_of(strategy=strategy, group_size=group_size)
    back = IntQuantizationCompressor.decompress(IntQuantizationCompressor.compress(sd, scheme=scheme), scheme=scheme)
    fq = fake_quantize(sd["weight"], scale=sd["weight_scale"], zero_point=sd["weight_zero_point"], args=scheme.weights)
    assert torch.equal(fq, back["weight"].to(torch.float32))


# ---- test_forward.py (block cases, fused vs sequential) ------------------------------------------
@pytest.mark.parametrize("rows,cols,bh,bw", [(128, 128, 128, 128), (256, 384, 128, 128), (200, 300, 128, 128), (100, 512, 64, 128), (7, 9, 4, 4)])
def test_block_lossless_and_shapes(rows, cols, bh, bw):
    args = QuantizationArgs(num_bits=8, type="float", strategy="block", block_structure=[bh, bw])
    nrb, ncb = math.ceil(rows / bh), math.ceil(cols / bw)
    for value in (1.0, 0.5):
        x = torch.full((rows, cols), value, device=DEV)
        scale = torch.ones((nrb, ncb), device=DEV)
        q = quantize(x, scale, None, args, dtype=torch.float8_e4m3fn)
        assert q.shape == x.shape
        assert torch.equal(dequantize(q, scale, None, args=args), x)
        assert torch.equal(fake_quantize(x, scale, None, args), x)


@pytest.mark.parametrize("kw", [dict(num_bits=8, type="int", strategy="channel"), dict(num_bits=4, type="int", strategy="group", group_size=32),
                                dict(num_bits=8, type="float", strategy="tensor"), dict(num_bits=8, type="int", strategy="token", dynamic=True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_fused_qdq_equals_sequential(kw, dtype):
    """fake_quantize == dequantize(quantize()) exactly (the reference only asserts atol=1 / rtol=0.15 here)"""
    args = QuantizationArgs(**kw)
    x = (torch.randn(64, 256, device=DEV) * 0.1).to(dtype)
    if args.strategy in ("channel", "token"):
        scale = (x.float().abs().amax(-1, keepdim=True) / 127).to(dtype)
    elif args.strategy == "group":
        scale = (x.float().unflatten(-1, (-1, 32)).abs().amax(-1) / 7.5).to(dtype)
    else:
        scale = (x.float().abs().max() / 448).to(dtype).reshape(1)
    q = quantize(x, scale, None, args, dtype=args.pytorch_dtype())
    seq = dequantize(q, scale, None, args=args)
    assert torch.equal(fake_quantize(x, scale, None, args).to(seq.dtype), seq)


# --------------------------------------------------------------------------- #
# test_static_lifecycle.py: the reference's known answers for fake_quantize, strategy by strategy (tests/kat_static.py)
# --------------------------------------------------------------------------- #
from tests import kat_static  # noqa: E402


@pytest.mark.parametrize("case", kat_static.CASES, ids=[c[0] for c in kat_static.CASES])
def test_static_lifecycle_known_answers(case):
    out, want = kat_static.run(case, lambda x, s, z, a, gs: fake_quantize(x, s, z, a, global_scale=gs), device=DEV)
    assert out.dtype == torch.bfloat16 and out.device.type == "cuda"
    assert torch.equal(out.cpu(), want), (out.cpu(), want)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("symmetric", [True, False])
def test_attn_head_strategy_matches_oracle(dt, symmetric):
    """attn_head: one scale per head, shape [heads, 1, 1], broadcast against [batch, heads, seq, head_dim] (forward.py:229-241)"""
    import oracle

    torch.manual_seed(3)
    x = (torch.randn(3, 8, 17, 64) * 2).to(dt)
    args = QuantizationArgs(num_bits=8, type="int", symmetric=symmetric, strategy="attn_head")
    s = (torch.rand(8, 1, 1) * 0.05 + 0.01).to(dt)
    z = None if symmetric else torch.randint(-20, 20, (8, 1, 1)).to(torch.int8)
    kw = dict(strategy="attn_head", num_bits=8, qtype="int")
    X, S, Z = x.to(DEV), s.to(DEV), None if z is None else z.to(DEV)
    want_q = oracle.quantize(x, s, z, dtype=torch.int8, **kw)
    assert torch.equal(quantize(X, S, Z, args, dtype=torch.int8).cpu(), want_q)
    assert torch.equal(fake_quantize(X, S, Z, args).cpu(), oracle.fake_quantize(x, s, z, **kw))
    assert torch.equal(dequantize(want_q.to(DEV), S, Z, args=args).cpu(), oracle.dequantize(want_q, s, z, strategy="attn_head"))


# --------------------------------------------------------------------------- #
# test_compress_decompress_module.py: every preset through compress_module / decompress_module on the device, Linear and Embedding
# (the reference runs this file on "cuda" only: its body is @requires_gpu)
# --------------------------------------------------------------------------- #
from compressed_tensors_b200.compressors.base import compress_module, decompress_module  # noqa: E402
from compressed_tensors_b200.config import CompressionFormat  # noqa: E402
from compressed_tensors_b200.quantization import ActivationOrdering, initialize_module_for_quantization, preset_name_to_scheme  # noqa: E402
from compressed_tensors_b200.utils import get_direct_state_dict  # noqa: E402

_F, _G = CompressionFormat, ActivationOrdering.GROUP
_LINEAR_PRESETS = [
    ("UNQUANTIZED", _F.dense, None), ("W8A16", _F.pack_quantized, None), ("W4A16", _F.pack_quantized, None), ("W4A16", _F.pack_quantized, _G),
    ("W4A16_ASYM", _F.pack_quantized, None), ("W4A16_ASYM", _F.pack_quantized, _G), ("W8A8", _F.int_quantized, None), ("W4A8", _F.int_quantized, None),
    ("W4AFP8", _F.int_quantized, None), ("FP8", _F.float_quantized, None), ("FP8_DYNAMIC", _F.float_quantized, None), ("FP8_BLOCK", _F.float_quantized, None),
    ("NVFP4A16", _F.nvfp4_pack_quantized, None), ("NVFP4", _F.nvfp4_pack_quantized, None), ("MXFP4A16", _F.mxfp4_pack_quantized, None), ("MXFP4", _F.mxfp4_pack_quantized, None),
]
_EMBEDDING_PRESETS = [
    ("UNQUANTIZED", _F.dense, None), ("W8A16", _F.pack_quantized, None), ("W4A16", _F.pack_quantized, None), ("W4A16", _F.pack_quantized, _G),
    ("W4A16_ASYM", _F.pack_quantized, None), ("NVFP4A16", _F.nvfp4_pack_quantized, None), ("MXFP4A16", _F.mxfp4_pack_quantized, None),
]


def _compress_decompress_module(scheme_name, expected_format, actorder, module, targets):
    module = module.to(dtype=torch.bfloat16, device=DEV)
    scheme = preset_name_to_scheme(scheme_name, list(targets))
    if actorder is not None:
        scheme.weights.actorder = actorder
    initialize_module_for_quantization(module, scheme)
    with torch.no_grad():
        for _, param in list(module.named_parameters()):
            param.fill_(1)
    pre = {n: (t.shape, t.dtype) for n, t in get_direct_state_dict(module).items() if t is not None}
    w = module.weight.detach().clone()
    compress_module(module)
    assert module.quantization_scheme.format == expected_format
    decompress_module(module)
    for n, t in get_direct_state_dict(module).items():
        if n in pre:
            assert t.shape == pre[n][0] and t.dtype == pre[n][1], (n, t.shape, t.dtype, pre[n])
            assert t.device.type == "cuda"
    # values: the reference's fill sets every parameter to 1 -- scales AND (for the symmetric presets too) zero points, which the
    # symmetric formats do not store -- so only the trivially exact cases are compared here; value parity of every format is the
    # subject of the tests above and of tests/test_gpu_compressors.py
    if scheme.weights is None:
        assert torch.equal(module.weight.detach(), w)
    assert bool(torch.isfinite(module.weight.detach().float()).all())


@pytest.mark.parametrize("scheme_name,expected_format,actorder", _LINEAR_PRESETS, ids=[f"{n}-{a.value if a else 'none'}" for n, _, a in _LINEAR_PRESETS])
def test_compress_decompress_module(scheme_name, expected_format, actorder):
    _compress_decompress_module(scheme_name, expected_format, actorder, torch.nn.Linear(256, 256, bias=False), ("Linear",))


@pytest.mark.parametrize("scheme_name,expected_format,actorder", _EMBEDDING_PRESETS, ids=[f"{n}-{a.value if a else 'none'}" for n, _, a in _EMBEDDING_PRESETS])
def test_compress_decompress_embedding(scheme_name, expected_format, actorder):
    _compress_decompress_module(scheme_name, expected_format, actorder, torch.nn.Embedding(256, 256), ("Embedding",))


_ATTN = __import__("tests.golden", fromlist=["load"]).load("attn")


@pytest.mark.parametrize("i", range(len(_ATTN)))
def test_attn_head_golden(i):
    """the reference's own outputs for the ATTN_HEAD strategy (tests/golden/make_golden_attn.py) through the CUDA kernels"""
    from tests.util import bits_equal

    c = _ATTN[i]
    args = QuantizationArgs(**{k: v for k, v in c["args"].items() if v is not None})
    x, s, z = c["x"].to(DEV), c["scale"].to(DEV), c["zp"].to(DEV)
    q = quantize(x, s, z, args, dtype=c["q"].dtype).cpu()
    qa, qb = (q.view(torch.uint8), c["q"].view(torch.uint8)) if q.dtype == torch.float8_e4m3fn else (q, c["q"])
    assert torch.equal(qa, qb)
    assert bits_equal(dequantize(c["q"].to(DEV), s, z, args=args).cpu(), c["dq"])
    assert bits_equal(fake_quantize(x, s, z, args).cpu(), c["fq"])

"""
The `device = -1` CPU twins of the seven per-tensor hot-path entry points (csrc/cpu_twin.cu, host code inside libct_b200.so)
through the PRODUCT's public ops, against the golden vectors the reference itself produced (tests/golden/*.pt.gz) and the reference
tests' known answers.  No GPU: on a GPU-less host ImplBackend selects the eager body = these twins; with a GPU the same body is
forced with CT_ENFORCE_EAGER=1.  BASELINE config 1 ("int4 round-trip on [1024, 4096], CPU only") passes through the real product here.
"""
import ctypes

import pytest
import torch

from compressed_tensors_b200 import _native as N
from compressed_tensors_b200 import ops
from compressed_tensors_b200.utils import ImplBackend
from tests.golden import load
from tests.test_oracle_golden import KATS, _case_id, _old_pack, _u32
from tests.util import bits_equal, diff_report


@pytest.fixture(autouse=True)
def eager(monkeypatch):
    """select the eager body also on a box that has a GPU"""
    monkeypatch.setenv("CT_ENFORCE_EAGER", "1")
    before = N.launch_count() if torch.cuda.is_available() else 0
    yield
    if torch.cuda.is_available():
        assert N.launch_count() == before, "the eager body must not launch kernels"


class A:   # QuantizationArgs stand-in built from a golden case's dumped args
    def __init__(self, d):
        self.strategy, self.group_size, self.block_structure = d["strategy"], d.get("group_size"), d.get("block_structure")
        self.num_bits, self.type, self.symmetric = d["num_bits"], d["type"], d["symmetric"]


@pytest.mark.parametrize("bits,vals,words", KATS)
def test_pack_kat(bits, vals, words):
    v = torch.tensor(vals, dtype=torch.int8)
    want = torch.stack([_u32(w) for w in words])
    got = ops.pack_to_int32(v, bits)
    assert got.device.type == "cpu" and torch.equal(got, want)
    assert torch.equal(ops.unpack_from_int32(got, bits, v.shape), v)


def test_config1_int4_round_trip_1024x4096():
    """BASELINE.json configs[0]: int4 PackedQuantizationCompressor round trip on a single [1024, 4096] int tensor, CPU only"""
    g = torch.Generator().manual_seed(1000)
    q = torch.randint(-8, 8, (1024, 4096), dtype=torch.int8, generator=g)
    packed = ops.pack_to_int32(q, 4)
    assert packed.shape == (1024, 512) and packed.dtype == torch.int32
    assert torch.equal(packed, _old_pack(q, 4)), "independent element-aligned packer (reference test_pack_quant.py:27-39)"
    assert torch.equal(ops.unpack_from_int32(packed, 4, q.shape), q)
    # and through the compressor plugin on a float weight
    from compressed_tensors_b200.compressors import PackedQuantizationCompressor
    from compressed_tensors_b200.quantization import preset_name_to_scheme

    w = (torch.randn(1024, 4096, generator=g) * 0.02).bfloat16()
    sc = (w.float().unflatten(-1, (-1, 128)).abs().amax(-1) / 7.5).bfloat16()
    scheme = preset_name_to_scheme("W4A16", ["Linear"])
    sd = PackedQuantizationCompressor.compress({"weight": w, "weight_scale": sc, "weight_zero_point": torch.zeros(sc.shape, dtype=torch.int8)}, scheme)
    back = PackedQuantizationCompressor.decompress(sd, scheme)["weight"]
    assert torch.equal(back, ops.fake_quantize(w, sc, None, scheme.weights)), "decompress(compress(w)) == fake_quantize(w) (reference test_pack_quant.py:160-183)"


def test_pack_golden():
    for c in load("pack"):
        got = ops.pack_to_int32(c["value"], c["bits"], c["packed_dim"])
        assert torch.equal(got.contiguous(), c["packed"]), (c["bits"], c["packed_dim"], tuple(c["value"].shape))
        if not c.get("out_of_range"):
            assert torch.equal(ops.unpack_from_int32(c["packed"], c["bits"], c["value"].shape, c["packed_dim"]), c["value"])


_Q = load("quant")


@pytest.mark.parametrize("c", _Q["cases"], ids=_case_id)
def test_quant_golden(c):
    x = _Q["x"][c["x"]] if isinstance(c["x"], str) else c["x"]
    a = A(c["args"])
    q = ops.quantize(x, c["scale"], c["zp"], a, dtype=c["q"].dtype, g_idx=c["g_idx"])
    assert bits_equal(q, c["q"]), "quantize: " + diff_report(q, c["q"])
    qf = ops.quantize(x, c["scale"], c["zp"], a, dtype=None, g_idx=c["g_idx"])
    assert bits_equal(qf, c["qf"]), "quantize(dtype=None): " + diff_report(qf, c["qf"])
    dq = ops.dequantize(c["q"], c["scale"], c["zp"], a, g_idx=c["g_idx"])
    assert bits_equal(dq, c["dq"]), "dequantize: " + diff_report(dq, c["dq"])
    if c["dq_inferred"] is not None:
        dqi = ops.dequantize(c["q"], c["scale"], c["zp"], g_idx=c["g_idx"])
        assert bits_equal(dqi, c["dq_inferred"]), "dequantize(inferred): " + diff_report(dqi, c["dq_inferred"])
    fq = ops.fake_quantize(x, c["scale"], c["zp"], a, g_idx=c["g_idx"])
    assert bits_equal(fq, c["fq"]), "fake_quantize: " + diff_report(fq, c["fq"])


def test_sweep_golden():
    """every bf16 / fp16 bit pattern x 4 scales: int4, int8 + zero point, fp8 codes, fake_quantize -- the host rounding chain (incl.
    cuda_fp8.h's software e4m3 conversion standing in for the PTX instruction) equals the reference on all of them"""
    sw = load("sweep")
    pat = torch.arange(65536, dtype=torch.int32).to(torch.uint16)
    t = A(dict(strategy="tensor", num_bits=4, type="int", symmetric=True))
    t8 = A(dict(strategy="tensor", num_bits=8, type="int", symmetric=False))
    f8a = A(dict(strategy="tensor", num_bits=8, type="float", symmetric=True))
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        x = pat.view(dt).reshape(256, 256).clone()
        x[x.isnan()] = 0
        for sval in (2.0 ** -7, 0.01, 1.0, 37.5):
            s = torch.tensor([sval]).to(dt)
            key = f"{name}/s{sval}"
            assert torch.equal(ops.quantize(x, s, None, t, dtype=torch.int8), sw[key + "/int4"]), key
            assert torch.equal(ops.quantize(x, s, torch.tensor([3], dtype=torch.int8), t8, dtype=torch.int8), sw[key + "/int8zp3"]), key
            assert torch.equal(ops.quantize(x, s, None, f8a, dtype=torch.float8_e4m3fn).view(torch.uint8), sw[key + "/fp8"]), key
            assert torch.equal(ops.fake_quantize(x, s, None, t).view(torch.int16), sw[key + "/fq_int4"]), key
            assert torch.equal(ops.fake_quantize(x, s, None, f8a).view(torch.int16), sw[key + "/fq_fp8"]), key
    codes = torch.arange(-128, 128, dtype=torch.int8).reshape(1, 256)
    f8 = torch.arange(256, dtype=torch.int32).to(torch.uint8).view(torch.float8_e4m3fn).reshape(1, 256)
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16), ("fp32", torch.float32)):
        for sval in (0.00731, 0.02, 1.0, 1.7):
            s = torch.tensor([sval]).to(dt)
            assert bits_equal(ops.dequantize(codes, s, None), sw[f"dq/{name}/s{sval}/int8"])
            assert bits_equal(ops.dequantize(codes, s, torch.tensor([-5], dtype=torch.int8)), sw[f"dq/{name}/s{sval}/int8zp"])
            d = ops.dequantize(f8, s, None)
            d[d.isnan()] = 0
            assert bits_equal(d, sw[f"dq/{name}/s{sval}/fp8"])


@pytest.mark.parametrize("c", load("compressors"), ids=lambda c: f"{c['format']}-{c['tag']}")
def test_compressor_golden(c):
    """the plugin classes on CPU state dicts == the state dicts the reference produced (formats whose kernels have a CPU twin)"""
    from compressed_tensors_b200.compressors import BaseCompressor
    from compressed_tensors_b200.quantization import QuantizationScheme

    if c["format"] not in ("pack-quantized", "naive-quantized", "int-quantized", "float-quantized"):
        pytest.skip("no CPU twin for this format's kernels (fp4 / mx): the call raises, see test_ops_without_a_twin_still_refuse")
    scheme = QuantizationScheme.model_validate(c["scheme"])
    comp = BaseCompressor.get_value_from_registry(c["format"])
    got = comp.compress(c["state_dict"], scheme)
    assert set(got) == set(c["compressed"])
    for k, w in c["compressed"].items():
        g = got[k]
        assert g.dtype == w.dtype and g.shape == w.shape and g.device.type == "cpu", k
        a, b = (g.view(torch.uint8), w.view(torch.uint8)) if g.dtype == torch.float8_e4m3fn else (g.contiguous(), w.contiguous())
        assert torch.equal(a, b), f"compress[{k}]: " + diff_report(a, b)
    back = comp.decompress(c["compressed"], scheme)
    for k, w in c["decompressed"].items():
        g = back[k]
        a, b = (g.view(torch.uint8), w.view(torch.uint8)) if g.dtype == torch.float8_e4m3fn else (g.contiguous(), w.contiguous())
        assert g.dtype == w.dtype and bits_equal(a, b), f"decompress[{k}]: " + diff_report(a, b)


def test_explicit_twin_through_the_c_abi_and_no_silent_fallback():
    """device = -1 is the ONLY way into host code: device >= 0 on a GPU-less host still answers CT_E_NODEV"""
    L = N.lib()
    q = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 0], [-1, -2, -3, -4, -5, -6, -7, -8]], dtype=torch.int8)
    out = torch.zeros(2, 1, dtype=torch.int32)
    assert L.ct_pack_int32(N.ptr(q), N.ptr(out), 2, 8, 4, 1, -1, None) == N.CT_OK
    assert [v & 0xFFFFFFFF for v in out.flatten().tolist()] == [0x8FEDCBA9, 0x01234567]
    if not torch.cuda.is_available():
        assert L.ct_pack_int32(N.ptr(q), N.ptr(out), 2, 8, 4, 1, 0, None) == N.CT_E_NODEV
        assert "no CPU path" in N.last_error() or "CUDA" in N.last_error()
    # entry points without a twin refuse device = -1
    x = torch.zeros(4, 32, dtype=torch.bfloat16)
    o = torch.zeros(4, 16, dtype=torch.uint8)
    assert L.ct_pack_fp4(N.ptr(x), N.DT[x.dtype], N.ptr(o), 4, 32, -1, None) == N.CT_E_NODEV
    assert L.ct_sparse24_compress(N.ptr(x), N.DT[x.dtype], N.ptr(o), N.ptr(o), 4, 32, -1, None) == N.CT_E_NODEV


def test_impl_backend_registry_holds_both_bodies():
    q = torch.randint(-8, 8, (8, 64), dtype=torch.int8)
    eager = ImplBackend.call("pack_to_int32_eager", q, 4)
    assert torch.equal(eager, _old_pack(q, 4))
    assert {"quantize_sm100", "quantize_eager", "unpack_dequantize_sm100", "unpack_dequantize_eager"} <= set(ImplBackend._fn_registry)
    if not torch.cuda.is_available():
        with pytest.raises(N.NativeLibraryError):
            ImplBackend.call("pack_to_int32_sm100", q, 4)      # the CUDA backend never falls back

"""
The oracle (oracle/ct_oracle.c) against the golden vectors produced by running
the reference itself (tests/golden/make_golden.py) and against the reference
tests' known-answer vectors.  CPU only.
"""
import pytest
import torch

import oracle
from tests.golden import load
from tests.util import bits_equal, diff_report, okw


# --------------------------------------------------------------------------- #
# known-answer vectors (reference tests/test_compressors/test_pack_quant.py:103-131)
# --------------------------------------------------------------------------- #
def _u32(words):
    return torch.tensor([w - (1 << 32) if w >= (1 << 31) else w for w in words], dtype=torch.int32)


KATS = [
    (4, [[1, 2, 3, 4, 5, 6, 7, 0], [-1, -2, -3, -4, -5, -6, -7, -8]], [[0x8FEDCBA9], [0x01234567]]),
    (1, [[0, -1] * 16], [[0x55555555]]),
    (2, [[1, -2, -1, 0] * 8], [[0x93939393, 0x93939393]]),
    (3, [[-1, -4, 1, -2, 3, 0, -3, 2] * 4], [[0x43C67543, 0x7543C675, 0xC67543C6]]),
    (4, [[-5, 0, 5, -6, -1, 4, -7, -2, 3, -8, -3, 2, 7, -4, 1, 6] * 2], [[0x61C72D83, 0xE94FA50B, 0x61C72D83, 0xE94FA50B]]),
    (5, [[-16 + (5 * i + 3) % 32 for i in range(32)]], [[0x79793503, 0xFD560B30, 0x77137249, 0x1BB45871, 0xF668F514]]),
    (6, [[-32 + (5 * i + 3) % 64 for i in range(32)]], [[0x1748D203, 0x5C2B9A17, 0x38913FEB, 0x2789D613, 0x503BDB1B, 0x79950F28]]),
    (7, [[-64 + (5 * i + 3) % 128 for i in range(32)]], [[0x72434403, 0x2B4C84E1, 0x23F74D58, 0x6C539D26, 0xC7667C57, 0x41407BED, 0x3C64A0F1]]),
    (8, [[-128 + (5 * i + 3) for i in range(32)]], [[0x120D0803, 0x26211C17, 0x3A35302B, 0x4E49443F, 0x625D5853, 0x76716C67, 0x8A85807B, 0x9E99948F]]),
    (4, [[(i % 16) - 8 for i in range(33)]], [[0x76543210, 0xFEDCBA98, 0x76543210, 0xFEDCBA98, 0x00000000]]),
    (3, [[(i % 8) - 4 for i in range(33)]], [[0x88FAC688, 0xC688FAC6, 0xFAC688FA, 0x00000000]]),
]


@pytest.mark.parametrize("bits,vals,words", KATS)
def test_pack_kat(bits, vals, words):
    v = torch.tensor(vals, dtype=torch.int8)
    want = torch.stack([_u32(w) for w in words])
    got = oracle.pack_to_int32(v, bits)
    assert torch.equal(got, want)
    assert torch.equal(oracle.unpack_from_int32(got, bits, v.shape), v)


def test_pack_kat_dim0():
    z = torch.tensor([[-8, 7], [0, 1], [2, -3], [4, 5], [-6, 6], [3, -1], [-2, -4], [-7, -5], [1, 1]], dtype=torch.int8)
    want = torch.stack([_u32([0x16B2CA80, 0x347ED59F]), _u32([0x00000009, 0x00000009])])
    got = oracle.pack_to_int32(z, 4, packed_dim=0)
    assert torch.equal(got, want)
    assert torch.equal(oracle.unpack_from_int32(got, 4, z.shape, packed_dim=0), z)


def _old_pack(value: torch.Tensor, bits: int) -> torch.Tensor:
    """independent element-aligned packer in the spirit of the reference's
    `_old_pack_to_int32` known-answer check (test_pack_quant.py:27-39, :406-416):
    for power-of-two widths the dense layout equals one-element-per-slot packing."""
    per = 32 // bits
    rows, cols = value.shape
    pad = (-cols) % per
    u = (value.to(torch.int64) + (1 << (bits - 1)))
    u = torch.nn.functional.pad(u, (0, pad))
    u = u.reshape(rows, -1, per)
    sh = torch.arange(per) * bits
    w = (u << sh).sum(-1) & 0xFFFFFFFF
    w = torch.where(w >= (1 << 31), w - (1 << 32), w)
    return w.to(torch.int32)


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
@pytest.mark.parametrize("k", [33, 64, 100, 1024])
def test_pack_power_of_two_equals_element_aligned(bits, k):
    g = torch.Generator().manual_seed(bits * 1000 + k)
    v = torch.randint(-(1 << (bits - 1)), 1 << (bits - 1), (4, k), dtype=torch.int8, generator=g)
    assert torch.equal(oracle.pack_to_int32(v, bits), _old_pack(v, bits))


def test_pack_errors():
    with pytest.raises(ValueError, match="torch.int8"):
        oracle.pack_to_int32(torch.zeros(2, 2, dtype=torch.int32), 4)
    with pytest.raises(ValueError, match="num_bits"):
        oracle.pack_to_int32(torch.zeros(2, 2, dtype=torch.int8), 9)
    with pytest.raises(ValueError):
        oracle.unpack_from_int32(torch.zeros(2, 2, dtype=torch.int8), 4, (2, 2))


# --------------------------------------------------------------------------- #
# golden: pack
# --------------------------------------------------------------------------- #
def test_pack_golden():
    cases = load("pack")
    assert len(cases) > 100
    for c in cases:
        got = oracle.pack_to_int32(c["value"], c["bits"], c["packed_dim"])
        assert torch.equal(got.contiguous(), c["packed"]), (c["bits"], c["packed_dim"], tuple(c["value"].shape))
        if not c.get("out_of_range"):
            back = oracle.unpack_from_int32(c["packed"], c["bits"], c["value"].shape, c["packed_dim"])
            assert torch.equal(back, c["value"])


# --------------------------------------------------------------------------- #
# golden: quantize / dequantize / fake_quantize
# --------------------------------------------------------------------------- #
def _case_id(c):
    a = c["args"]
    x = c["x"] if isinstance(c["x"], str) else "act"
    return f"{x}-{a['strategy']}-g{a.get('group_size')}-b{a['num_bits']}{a['type']}-{'sym' if a['symmetric'] else 'asym'}-{c['tag']}"


_Q = load("quant")


@pytest.mark.parametrize("c", _Q["cases"], ids=_case_id)
def test_quant_golden(c):
    x = _Q["x"][c["x"]] if isinstance(c["x"], str) else c["x"]
    a = c["args"]
    kw = okw(a)
    q = oracle.quantize(x, c["scale"], c["zp"], dtype=c["q"].dtype, g_idx=c["g_idx"], **kw)
    assert bits_equal(q, c["q"]), "quantize: " + diff_report(q, c["q"])
    qf = oracle.quantize(x, c["scale"], c["zp"], dtype=None, g_idx=c["g_idx"], **kw)
    assert bits_equal(qf, c["qf"]), "quantize(dtype=None): " + diff_report(qf, c["qf"])
    dkw = dict(strategy=kw["strategy"], group_size=kw["group_size"], block_structure=kw["block_structure"])
    dq = oracle.dequantize(c["q"], c["scale"], c["zp"], g_idx=c["g_idx"], **dkw)
    assert bits_equal(dq, c["dq"]), "dequantize: " + diff_report(dq, c["dq"])
    if c["dq_inferred"] is not None:
        dqi = oracle.dequantize(c["q"], c["scale"], c["zp"], g_idx=c["g_idx"])
        assert bits_equal(dqi, c["dq_inferred"]), "dequantize(inferred): " + diff_report(dqi, c["dq_inferred"])
    fq = oracle.fake_quantize(x, c["scale"], c["zp"], g_idx=c["g_idx"], **kw)
    assert bits_equal(fq, c["fq"]), "fake_quantize: " + diff_report(fq, c["fq"])


# --------------------------------------------------------------------------- #
# golden: exhaustive bit-pattern sweeps
# --------------------------------------------------------------------------- #
def test_sweep_golden():
    sw = load("sweep")
    pat = torch.arange(65536, dtype=torch.int32).to(torch.uint16)
    n = 0
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        x = pat.view(dt).reshape(256, 256).clone()
        x[x.isnan()] = 0
        for sval in (2.0 ** -7, 0.01, 1.0, 37.5):
            s = torch.tensor([sval]).to(dt)
            key = f"{name}/s{sval}"
            got = oracle.quantize(x, s, None, strategy="tensor", num_bits=4, dtype=torch.int8)
            assert torch.equal(got, sw[key + "/int4"]), key
            zp = torch.tensor([3], dtype=torch.int8)
            got = oracle.quantize(x, s, zp, strategy="tensor", num_bits=8, dtype=torch.int8)
            assert torch.equal(got, sw[key + "/int8zp3"]), key
            got = oracle.quantize(x, s, None, strategy="tensor", num_bits=8, qtype="float", dtype=torch.float8_e4m3fn)
            assert torch.equal(got.view(torch.uint8), sw[key + "/fp8"]), key + " " + diff_report(got.view(torch.uint8), sw[key + "/fp8"])
            got = oracle.fake_quantize(x, s, None, strategy="tensor", num_bits=4)
            assert torch.equal(got.view(torch.int16), sw[key + "/fq_int4"]), key
            got = oracle.fake_quantize(x, s, None, strategy="tensor", num_bits=8, qtype="float")
            assert torch.equal(got.view(torch.int16), sw[key + "/fq_fp8"]), key
            n += 5
    codes = torch.arange(-128, 128, dtype=torch.int8).reshape(1, 256)
    f8 = torch.arange(256, dtype=torch.int32).to(torch.uint8).view(torch.float8_e4m3fn).reshape(1, 256)
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16), ("fp32", torch.float32)):
        for sval in (0.00731, 0.02, 1.0, 1.7):
            s = torch.tensor([sval]).to(dt)
            zp = torch.tensor([-5], dtype=torch.int8)
            assert bits_equal(oracle.dequantize(codes, s, None), sw[f"dq/{name}/s{sval}/int8"])
            assert bits_equal(oracle.dequantize(codes, s, zp), sw[f"dq/{name}/s{sval}/int8zp"])
            d = oracle.dequantize(f8, s, None)
            d[d.isnan()] = 0
            assert bits_equal(d, sw[f"dq/{name}/s{sval}/fp8"])
            n += 3
    assert n == len(sw)


def test_fp8_cast_matches_torch_exhaustively():
    """the oracle's float->e4m3fn rounding vs the third-party arithmetic the
    reference uses (torch `.to(float8_e4m3fn)`), on every bf16 and fp16 pattern"""
    import ctypes

    pat = torch.arange(65536, dtype=torch.int32).to(torch.uint16)
    for dt in (torch.bfloat16, torch.float16):
        x = pat.view(dt).float()
        x = x[~x.isnan()].contiguous()
        want = x.to(torch.float8_e4m3fn).view(torch.uint8)
        got = torch.empty(x.numel(), dtype=torch.uint8)
        oracle.lib().orc_cast_f32_to_f8e4m3(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(got.data_ptr()), ctypes.c_int64(x.numel()))
        assert torch.equal(got, want)
    g = torch.Generator().manual_seed(3)
    x = torch.cat([torch.randn(1 << 18, generator=g) * s for s in (1e-3, 0.1, 1.0, 30.0, 300.0)]).contiguous()
    want = x.to(torch.float8_e4m3fn).view(torch.uint8)
    got = torch.empty(x.numel(), dtype=torch.uint8)
    oracle.lib().orc_cast_f32_to_f8e4m3(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(got.data_ptr()), ctypes.c_int64(x.numel()))
    assert torch.equal(got, want)


def test_narrowing_casts_match_torch():
    import ctypes

    g = torch.Generator().manual_seed(5)
    x = torch.cat([torch.randn(1 << 18, generator=g) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 1e3, 7e4)]).contiguous()
    for fn, dt in (("orc_cast_f32_to_bf16", torch.bfloat16), ("orc_cast_f32_to_f16", torch.float16)):
        got = torch.empty(x.numel(), dtype=torch.int16)
        getattr(oracle.lib(), fn)(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(got.data_ptr()), ctypes.c_int64(x.numel()))
        assert torch.equal(got, x.to(dt).view(torch.int16)), fn
    h = torch.arange(65536, dtype=torch.int32).to(torch.uint16).view(torch.float16)
    got = torch.empty(65536, dtype=torch.float32)
    oracle.lib().orc_cast_f16_to_f32(ctypes.c_void_p(h.data_ptr()), ctypes.c_void_p(got.data_ptr()), ctypes.c_int64(65536))
    assert bits_equal(got[~h.isnan()], h.float()[~h.isnan()])


# --------------------------------------------------------------------------- #
# golden: bitmasks
# --------------------------------------------------------------------------- #
def test_bitmask_golden():
    sp = load("sparse")
    for c in sp["bitmask"]:
        got = oracle.pack_bitmasks(c["mask"])
        assert torch.equal(got, c["packed"])
        assert torch.equal(oracle.unpack_bitmasks(c["packed"], list(c["mask"].shape)), c["mask"])


def test_sparse24_restated_roundtrip():
    """restated format (parity unpinned): structural properties only"""
    g = torch.Generator().manual_seed(11)
    for dt in (torch.bfloat16, torch.float16, torch.float32, torch.int8):
        x = (torch.randn(16, 64, generator=g) * 10)
        x = x.round().to(dt) if dt == torch.int8 else x.to(dt)
        vals, bm = oracle.sparse24_compress(x)
        assert vals.shape == (16, 32) and bm.shape == (16, 8)
        mask = oracle.unpack_bitmasks(bm, x.shape)
        assert torch.equal(mask.reshape(-1, 4).sum(-1), torch.full((256,), 2))
        # kept = the two largest magnitudes of each quad
        xa = x.float().abs().reshape(-1, 4)
        kept_min = torch.where(mask.reshape(-1, 4), xa, torch.full_like(xa, float("inf"))).amin(-1)
        drop_max = torch.where(~mask.reshape(-1, 4), xa, torch.full_like(xa, -1.0)).amax(-1)
        assert bool((kept_min >= drop_max).all())
        dense = oracle.sparse24_decompress(vals, bm, x.shape)
        assert torch.equal(dense, x * mask.to(dt))
        # a 2:4 tensor round-trips exactly
        v2, b2 = oracle.sparse24_compress(dense)
        assert torch.equal(oracle.sparse24_decompress(v2, b2, x.shape), dense)


def test_bitmask_restated_roundtrip():
    g = torch.Generator().manual_seed(12)
    x = torch.randn(9, 37, generator=g).bfloat16()
    x[torch.rand(9, 37, generator=g) < 0.6] = 0
    vals, bm, offs = oracle.bitmask_compress(x)
    assert torch.equal(vals, x[x != 0])
    assert torch.equal(bm, oracle.pack_bitmasks(x != 0))
    cnt = (x != 0).sum(-1)
    assert torch.equal(offs, torch.cumsum(cnt, 0) - cnt)
    assert torch.equal(oracle.bitmask_decompress(vals, bm, x.shape), x)


def _quiet_snan_halves(t: torch.Tensor) -> torch.Tensor:
    """what the reference's CPU path does to fp32 data it scatters through a float16 view
    (semi_structured_conversions.py:289-293): 16-bit halves that look like fp16 signalling NaNs
    come back with the quiet bit set.  Device-dependent accident, not part of the format."""
    h = t.contiguous().view(torch.int16).clone()
    snan = ((h & 0x7C00) == 0x7C00) & ((h & 0x03FF) != 0)
    h[snan] |= 0x0200
    return h.view(t.dtype)


def test_semi_structured_golden():
    """2:4 CUTLASS metadata encode / reorder / decode against the reference's own functions"""
    sp = load("sparse")
    assert len(sp["semi"]) == 4
    for c in sp["semi"]:
        sparse, meta = oracle.semi_structured_from_dense(c["dense"])
        assert bits_equal(sparse, c["sparse"]), diff_report(sparse, c["sparse"])
        assert torch.equal(meta, c["meta"])
        back = oracle.semi_structured_to_dense(c["sparse"], c["meta"])
        if back.dtype == torch.float32:
            # the oracle moves bits; the reference additionally quiets fp16-sNaN-looking halves
            assert bits_equal(_quiet_snan_halves(back), c["back"])
            assert (back.view(torch.int32) != c["back"].view(torch.int32)).sum() < 0.05 * back.numel()
        else:
            assert bits_equal(back, c["back"]), diff_report(back, c["back"])


# --------------------------------------------------------------------------- #
# known answers held by the reference's own lifecycle tests (tests/kat_static.py)
# --------------------------------------------------------------------------- #
from tests import kat_static  # noqa: E402


@pytest.mark.parametrize("case", kat_static.CASES, ids=[c[0] for c in kat_static.CASES])
def test_static_lifecycle_known_answers(case):
    def fq(x, s, z, a, gs):
        return oracle.fake_quantize(x, s, z, strategy=a.strategy, group_size=a.group_size, block_structure=a.block_structure,
                                    num_bits=a.num_bits, qtype=a.type, global_scale=gs)

    out, want = kat_static.run(case, fq)
    assert out.dtype == torch.bfloat16 and torch.equal(out, want), diff_report(out, want)


# --------------------------------------------------------------------------- #
# ATTN_HEAD strategy: scale [heads, 1, 1] against [batch, heads, seq, head_dim] (tests/golden/make_golden_attn.py)
# --------------------------------------------------------------------------- #
_ATTN = load("attn")


@pytest.mark.parametrize("i", range(len(_ATTN)))
def test_attn_head_golden(i):
    c = _ATTN[i]
    a = c["args"]
    kw = dict(strategy="attn_head", num_bits=a["num_bits"], qtype=a["type"])
    q = oracle.quantize(c["x"], c["scale"], c["zp"], dtype=c["q"].dtype, **kw)
    qa, qb = (q.view(torch.uint8), c["q"].view(torch.uint8)) if q.dtype == torch.float8_e4m3fn else (q, c["q"])
    assert torch.equal(qa, qb), "quantize: " + diff_report(qa, qb)
    dq = oracle.dequantize(c["q"], c["scale"], c["zp"], strategy="attn_head")
    assert bits_equal(dq, c["dq"]), "dequantize: " + diff_report(dq, c["dq"])
    fq = oracle.fake_quantize(c["x"], c["scale"], c["zp"], **kw)
    assert bits_equal(fq, c["fq"]), "fake_quantize: " + diff_report(fq, c["fq"])

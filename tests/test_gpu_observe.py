"""
Fused observer + quantize + pack against the composition of pinned pieces:
group min/max -> oracle calculate_qparams (pinned by qparams golden) -> oracle quantize -> oracle pack.
"""
from types import SimpleNamespace

import pytest
import torch

import oracle
from compressed_tensors_b200 import _native as N
from compressed_tensors_b200 import ops
from oracle.qparams import calculate_qparams as orc_qparams
from tests.util import same, same_values

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _args(bits, group, sym):
    return SimpleNamespace(strategy="group", group_size=group, block_structure=None, num_bits=bits, type="int", symmetric=sym,
                           zp_dtype=torch.int8, scale_dtype=None)


def _want(w, bits, group, sym):
    wr = w.unflatten(-1, (-1, group))
    mn, mx = wr.amin(-1), wr.amax(-1)          # exact, in the weight dtype
    scale, zp = orc_qparams(mn, mx, num_bits=bits, qtype="int", symmetric=sym)
    zp = None if sym else zp
    q = oracle.quantize(w, scale, zp, strategy="group", group_size=group, num_bits=bits, dtype=torch.int8)
    return oracle.pack_to_int32(q, bits), scale, zp


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("group", [32, 64, 128, 256])
@pytest.mark.parametrize("sym", [True, False])
def test_fused_observer_vs_oracle(dtype, bits, group, sym):
    g = torch.Generator().manual_seed(bits * 100 + group)
    w = (torch.randn(192, 2048, generator=g) * 0.02).to(dtype)
    w[0, :group] = 0                      # all-zero group -> eps scale
    w[1, :group] = w[1, :group].abs()     # one-sided groups
    w[2, :group] = -w[2, :group].abs()
    w[3, 5] = 3.0                         # outlier
    want_p, want_s, want_z = _want(w, bits, group, sym)
    launches = N.launch_count()
    packed, scale, zp = ops.observe_quantize_pack(w.to(DEV), _args(bits, group, sym))
    assert N.launch_count() - launches == 1, "observer + quantize + pack must be one kernel"
    same(scale.cpu(), want_s, "scale")
    if sym:
        assert zp is None
    else:
        same_values(zp.cpu(), want_z, "zero point")
    same_values(packed.cpu(), want_p, "packed")


def test_fused_observer_full_size_properties():
    """Llama-3-8B gate_proj shape: fused result == separate observer + quantize_pack"""
    g = torch.Generator(device=DEV).manual_seed(7)
    w = (torch.randn(14336, 4096, device=DEV, generator=g) * 0.02).bfloat16()
    a = _args(4, 128, True)
    packed, scale, zp = ops.observe_quantize_pack(w, a)
    ref_scale = (w.unflatten(-1, (-1, 128)).abs().amax(-1).float() / 7.5).bfloat16()
    same(scale, ref_scale, "scale vs torch observer")
    same_values(packed, ops.quantize_pack(w, scale, None, a), "packed vs quantize_pack with the same scale")


def test_unfused_cases_compose():
    w = (torch.randn(64, 96) * 0.02).to(torch.bfloat16)      # group 48: not a fused size
    a = SimpleNamespace(strategy="channel", group_size=None, block_structure=None, num_bits=4, type="int", symmetric=True,
                        zp_dtype=torch.int8, scale_dtype=None)
    packed, scale, zp = ops.observe_quantize_pack(w.to(DEV), a)
    mn, mx = w.amin(-1, keepdim=True), w.amax(-1, keepdim=True)
    ws, _ = orc_qparams(mn, mx, num_bits=4, symmetric=True)
    same(scale.cpu(), ws, "channel scale")
    same_values(packed.cpu(), oracle.pack_to_int32(oracle.quantize(w, ws, None, strategy="channel", num_bits=4, dtype=torch.int8), 4), "packed")


# ---- channel-wise fused observer (ct_observe_quantize_channel) ------------------------------------------------------------
def _chan_oracle(x, qtype, bits, symmetric, pack):
    import oracle
    from oracle.qparams import calculate_qparams as oq

    mn, mx = x.amin(-1, keepdim=True), x.amax(-1, keepdim=True)
    s, z = oq(mn, mx, num_bits=bits, qtype=qtype, symmetric=symmetric)
    z = None if symmetric else z
    kw = dict(strategy="channel", num_bits=bits, qtype=qtype)
    q = oracle.quantize(x, s, z, dtype=(torch.float8_e4m3fn if qtype == "float" else torch.int8), **kw)
    return (oracle.pack_to_int32(q, bits) if pack else q), s, z


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("kind", [("int", 8, True, False), ("int", 8, False, False), ("float", 8, True, False), ("int", 4, True, True),
                                  ("int", 4, False, True), ("int", 8, True, True)])
@pytest.mark.parametrize("shape", [(64, 4096), (33, 14336), (7, 8), (5, 1000), (16, 16384)])
def test_channel_observer_vs_oracle(dt, kind, shape):
    from compressed_tensors_b200.quantization import QuantizationArgs
    from tests.util import same

    qtype, bits, sym, pack = kind
    g = torch.Generator().manual_seed(shape[1] + bits)
    x = (torch.randn(shape, generator=g) * torch.exp2(torch.randint(-8, 2, (shape[0], 1), generator=g).float())).to(dt)
    x[0] = 0                      # dead channel: scale falls back to eps
    if shape[0] > 2:
        x[1] = x[1].abs()         # one-sided channel
        x[2] = -x[2].abs()
    a = QuantizationArgs(num_bits=bits, type=qtype, symmetric=sym, strategy="channel")
    q, s, z = ops.observe_quantize(x.to(DEV), a, pack=pack)
    wq, ws, wz = _chan_oracle(x, qtype, bits, sym, pack)
    same(s.cpu(), ws, "channel observer scale")
    if sym:
        assert z is None
    else:
        same(z.cpu(), wz, "channel observer zero point")
    if wq.dtype == torch.float8_e4m3fn:
        same(q.cpu().view(torch.uint8), wq.view(torch.uint8), "channel observer fp8 codes")
    else:
        same(q.cpu(), wq, "channel observer codes")


def test_channel_observer_equals_unfused_flow_at_full_size():
    from compressed_tensors_b200.quantization import QuantizationArgs
    from compressed_tensors_b200.quantization.utils import calculate_qparams

    x = (torch.randn(14336, 4096, device=DEV) * 0.02).to(torch.bfloat16)
    for kw, pack in ((dict(num_bits=8, type="int", symmetric=True), False), (dict(num_bits=8, type="float", symmetric=True), False),
                     (dict(num_bits=4, type="int", symmetric=False), True)):
        a = QuantizationArgs(strategy="channel", **kw)
        q, s, z = ops.observe_quantize(x, a, pack=pack)
        s2, z2 = calculate_qparams(x.amin(-1, keepdim=True), x.amax(-1, keepdim=True), a)
        z2 = None if a.symmetric else z2
        assert torch.equal(s, s2) and (z is None or torch.equal(z, z2))
        q2 = ops.quantize_pack(x, s2, z2, a) if pack else ops.quantize(x, s2, z2, a, dtype=a.pytorch_dtype())
        assert torch.equal(q.view(torch.uint8) if q.dtype == torch.float8_e4m3fn else q, q2.view(torch.uint8) if q2.dtype == torch.float8_e4m3fn else q2)

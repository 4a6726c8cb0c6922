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


# ---- per-TENSOR fused observer (ct_observe_tensor / ct_observe_quantize_tensor) -----------------------------------------------
def _tensor_oracle(x, qtype, bits, symmetric, pack):
    mn, mx = x.amin().reshape(1), x.amax().reshape(1)
    s, z = orc_qparams(mn, mx, num_bits=bits, qtype=qtype, symmetric=symmetric)        # pinned by tests/golden/qparams.pt.gz
    z = None if symmetric else z
    q = oracle.quantize(x, s, z, strategy="tensor", num_bits=bits, qtype=qtype, dtype=(torch.float8_e4m3fn if qtype == "float" else torch.int8))
    return (oracle.pack_to_int32(q, bits) if pack else q), s, z


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("kind", [("float", 8, True, False), ("int", 8, True, False), ("int", 8, False, False), ("int", 4, True, True), ("int", 4, False, True)])
@pytest.mark.parametrize("shape", [(64, 4096), (1, 8), (300, 1000), (4096, 4096), (14336, 4096)])
def test_tensor_observer_vs_oracle(dt, kind, shape):
    """one ABI call: grid-wide min / max -> calculate_qparams on the device (last CTA) -> quantize [+ pack] with the device-resident
    scale == the oracle's calculate_qparams (golden-pinned) -> quantize -> pack_to_int32; no host round trip in between"""
    from compressed_tensors_b200.quantization import QuantizationArgs

    qtype, bits, sym, pack = kind
    if dt == torch.float32 and shape[0] > 4096:
        pytest.skip("fp32 at this size adds nothing")
    g = torch.Generator().manual_seed(shape[1] + bits + shape[0])
    x = (torch.randn(shape, generator=g) * 0.02).to(dt)
    if shape[0] > 1:
        x[shape[0] // 2, 3] = -0.31 if sym else 0.5      # the extreme sits in the middle of the tensor, one-sided for the asymmetric case
    a = QuantizationArgs(num_bits=bits, type=qtype, symmetric=sym, strategy="tensor")
    launches = N.launch_count()
    q, s, z = ops.observe_quantize(x.to(DEV), a, pack=pack)
    assert N.launch_count() - launches == 2, "min/max + qparams kernel, then the streaming quantize kernel"
    wq, ws, wz = _tensor_oracle(x, qtype, bits, sym, pack)
    assert s.shape == (1,) and s.dtype == dt
    same(s.cpu(), ws, "tensor observer scale")
    if sym:
        assert z is None
    else:
        same(z.cpu(), wz, "tensor observer zero point")
    if wq.dtype == torch.float8_e4m3fn:
        same(q.cpu().view(torch.uint8), wq.view(torch.uint8), "tensor observer fp8 codes")
    else:
        same(q.cpu(), wq, "tensor observer codes")


def test_tensor_observer_corner_cases():
    from compressed_tensors_b200.quantization import QuantizationArgs
    from compressed_tensors_b200.quantization.utils import calculate_qparams, generate_gparam

    a8 = QuantizationArgs(num_bits=8, type="float", symmetric=True, strategy="tensor")
    for x in (torch.zeros(32, 64, dtype=torch.bfloat16), torch.full((16, 16), -0.0, dtype=torch.float16),         # dead tensor: eps scale
              torch.full((8, 8), 3.0e38, dtype=torch.bfloat16), -torch.rand(64, 64).bfloat16() - 1.0,                # huge; all negative
              torch.rand(128, 8).half() * 1e-7):                                                                      # subnormal halves
        xd = x.to(DEV)
        s, z = ops.observe_tensor_qparams(xd, a8)
        ws, _ = calculate_qparams(*(t.reshape(1) for t in torch.aminmax(xd)), a8)                                    # host mirror (golden-pinned)
        same(s, ws, "qparams corner case")
        assert z is None
        g = ops.observe_tensor_gparam(xd)
        same(g.cpu(), generate_gparam(x.min(), x.max()), "gparam corner case")                                    # the CPU value is the pinned one
        os_, _ = orc_qparams(x.amin().reshape(1), x.amax().reshape(1), num_bits=8, qtype="float", symmetric=True)
        same(s.cpu(), os_, "qparams corner case vs oracle")
    # odd sizes are declined loudly (the Python layer falls back to torch reductions)
    with pytest.raises(NotImplementedError):
        ops.observe_tensor_qparams(torch.randn(3, 5, device=DEV).bfloat16(), a8)
    q, s, z = ops.observe_quantize(torch.randn(3, 5, device=DEV).bfloat16(), a8)
    assert q.dtype == torch.float8_e4m3fn and s.shape == (1,)


def test_nvfp4_observer_uses_the_device_gparam():
    """observe_quantize_pack_nvfp4 without a given global scale: generate_gparam comes from ct_observe_tensor (kind 1), bit-equal to
    the host mirror's generate_gparam(min, max) (pinned by tests/golden/fp4.pt.gz), and nothing synchronises"""
    from compressed_tensors_b200.quantization import QuantizationArgs
    from compressed_tensors_b200.quantization.utils import generate_gparam

    a = QuantizationArgs(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16)
    x = (torch.randn(512, 2048, device=DEV) * 0.02).bfloat16()
    launches = N.launch_count()
    packed, scale, gs = ops.observe_quantize_pack_nvfp4(x, a)
    assert N.launch_count() - launches == 2
    want = generate_gparam(x.min(), x.max())
    same(gs, want, "global scale")
    p2, s2, _ = ops.observe_quantize_pack_nvfp4(x, a, global_scale=want)
    same_values(packed, p2, "nibbles"); same_values(scale.view(torch.uint8), s2.view(torch.uint8), "fp8 scales")

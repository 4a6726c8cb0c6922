"""
BASELINE config 4 -- "Sparse24BitMask + int4": the fused 2:4 select + quantize + pack kernels (csrc/fast_sparse24q.cu) against the
oracle's restated composition of pinned pieces (quantize, pack_to_int32, pack_bitmasks) and the restated 2:4 selection rule.
PARITY UNPINNED as a composite (the compressor pair is absent from the reference snapshot); what can be pinned is pinned:
the kept codes equal the reference-pinned `quantize` on the kept columns, the words follow `pack_to_int32`'s bitstream, the mask
bytes follow `pack_bitmasks`, and on `w * mask_creator(w)` inputs (mask_creator is golden-pinned) the mask equals mask_creator's.
"""
from types import SimpleNamespace

import pytest
import torch

import oracle
from compressed_tensors_b200 import _native as N
from compressed_tensors_b200 import ops
from compressed_tensors_b200.utils.semi_structured_conversions import mask_creator
from tests.util import same, same_values

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ns(**kw):
    d = dict(strategy="group", group_size=128, block_structure=None, num_bits=4, type="int", symmetric=True)
    d.update(kw)
    return SimpleNamespace(**d)


def _qparams(w, strategy, group, sym):
    wf = w.float()
    if strategy == "group":
        g = wf.unflatten(-1, (-1, group))
        mn, mx = g.amin(-1).clamp(max=0), g.amax(-1).clamp(min=0)
    elif strategy == "channel":
        mn, mx = wf.amin(-1, keepdim=True).clamp(max=0), wf.amax(-1, keepdim=True).clamp(min=0)
    else:
        mn, mx = wf.min().clamp(max=0).reshape(1), wf.max().clamp(min=0).reshape(1)
    if sym:
        scale, zp = torch.maximum(mn.abs(), mx.abs()) / 7.5, None
    else:
        scale = (mx - mn) / 15.0
        zp = (-8 - mn / scale).clamp(-8, 7).round().to(torch.int8)
    scale = scale.to(w.dtype)
    scale[scale == 0] = torch.finfo(w.dtype).eps
    return scale, zp


CASES = [
    # (shape, dtype, strategy, group, symmetric, expect the fused kernel)
    ((64, 256), torch.bfloat16, "group", 128, True, True),
    ((64, 256), torch.float16, "group", 128, False, True),
    ((128, 4096), torch.bfloat16, "group", 32, True, True),
    ((33, 1024), torch.bfloat16, "group", 128, True, True),
    ((3, 32), torch.bfloat16, "channel", None, True, False),       # rows * cols % 64 != 0: unfused composition
    ((96, 2048), torch.bfloat16, "channel", None, False, True),
    ((64, 512), torch.float16, "tensor", None, True, True),
    ((2048, 14336), torch.bfloat16, "group", 128, True, True),     # many tiles, dynamic schedule
    ((16, 72), torch.bfloat16, "channel", None, True, False),      # cols % 32 != 0: unfused composition
    ((32, 128), torch.float32, "group", 64, True, False),          # fp32: unfused composition
]


@pytest.mark.parametrize("shape,dtype,strategy,group,sym,fused", CASES, ids=lambda v: str(v).replace("torch.", ""))
def test_fused_sparse24_int4_vs_oracle(shape, dtype, strategy, group, sym, fused):
    g = torch.Generator().manual_seed(shape[0] * 7 + shape[1])
    w = (torch.randn(shape, generator=g) * 0.02).to(dtype)
    w[0, :8] = torch.tensor([0.0, -0.0, 0.0, 0.0, 0.01, 0.01, -0.01, 0.01], dtype=dtype)   # all-zero quad (ties -> lower columns), magnitude ties
    if shape[0] >= 64:
        w = w * mask_creator(w).to(dtype)                                                    # config 4's input: 2:4-pruned weights
    sc, zp = _qparams(w, strategy, group, sym)
    a = ns(strategy=strategy, group_size=group, symmetric=sym)
    want_p, want_b = oracle.sparse24_quantize_pack(w, sc, zp, strategy=strategy, group_size=group)
    launches = N.launch_count()
    got_p, got_b = ops.sparse24_quantize_pack(w.to(DEV), sc.to(DEV), zp.to(DEV) if zp is not None else None, a)
    n_launch = N.launch_count() - launches
    eligible = dtype != torch.float32 and shape[1] % 32 == 0 and (shape[0] * shape[1]) % 64 == 0
    assert (n_launch == 1) == eligible, f"{n_launch} launches; fused kernel expected: {eligible}"
    same_values(got_b.cpu(), want_b, "bitmask bytes")
    same_values(got_p.cpu(), want_p, "packed words")
    if shape[0] >= 64:
        # (row 0 holds an all-zero quad, where topk's choice inside mask_creator is implementation-defined)
        same_values(ops.unpack_bitmasks(got_b, shape).cpu()[1:], mask_creator(w)[1:].bool(), "on 2:4-pruned input the mask is mask_creator's (golden-pinned)")
    # the way back
    want_d = oracle.sparse24_unpack_dequantize(want_p, want_b, sc, zp, 4, shape)
    launches = N.launch_count()
    got_d = ops.sparse24_unpack_dequantize(got_p, got_b, sc.to(DEV), zp.to(DEV) if zp is not None else None, 4, shape)
    assert ((N.launch_count() - launches) == 1) == eligible
    same(got_d.cpu(), want_d, "dequantized dense")
    # property: decompress(compress(w)) == mask * fake_quantize(w) (the composite of two reference-pinned ops)
    fq = oracle.fake_quantize(w, sc, zp, strategy=strategy, group_size=group, num_bits=4)
    mask = oracle.unpack_bitmasks(want_b, shape)
    same_values(got_d.cpu(), torch.where(mask, fq, torch.zeros_like(fq)), "== mask * fake_quantize")


def test_fused_sparse24_int4_batched_llama_layer():
    """one multi-tensor launch per direction over Llama-3-8B-shaped 2:4-pruned weights (the bench row of config 4) == per-tensor calls"""
    shapes = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336)]
    a = ns()
    cprobs, dprobs, singles, keep = [], [], [], []
    for i, (R, C) in enumerate(shapes):
        g = torch.Generator(device=DEV).manual_seed(2000 + i)
        w = (torch.randn(R, C, device=DEV, generator=g) * 0.02).bfloat16()
        w = w * mask_creator(w).to(w.dtype)
        sc = (w.unflatten(-1, (-1, 128)).abs().amax(-1).float() / 7.5).bfloat16()
        singles.append(ops.sparse24_quantize_pack(w, sc, None, a))
        p = ops._resolve(w, sc, None, a, None)
        packed = torch.zeros(R, C // 16, dtype=torch.int32, device=DEV)
        bm = torch.zeros(R, C // 8, dtype=torch.uint8, device=DEV)
        back = torch.empty(R, C, dtype=torch.bfloat16, device=DEV)
        d = ops._desc(p, w.dtype, sc.dtype, None, torch.bfloat16, torch.int8, None, N.Q_INT, 4)
        d.aux = bm.data_ptr()
        d2 = ops._desc(p, None, sc.dtype, None, None, torch.int8, torch.bfloat16, N.Q_INT, 4)
        d2.aux = bm.data_ptr()
        cprobs.append((d, w, sc, None, packed))
        dprobs.append((d2, packed, sc, None, back))
        keep.append((w, sc, packed, bm, back))
    launches = N.launch_count()
    ops.batched(N.OP_SPARSE24_QUANTIZE_PACK, cprobs)
    ops.batched(N.OP_SPARSE24_UNPACK_DEQUANTIZE, dprobs)
    assert N.launch_count() - launches == 2
    for (w, sc, packed, bm, back), (sp, sb) in zip(keep, singles):
        same_values(packed, sp, "batched packed")
        same_values(bm, sb, "batched mask")
        same(back, ops.sparse24_unpack_dequantize(sp, sb, sc, None, 4, w.shape), "batched decompress")
        # full-tensor oracle check of the largest row block is covered per tensor above; here a row slab against the oracle
        rows = slice(0, 256)
        wp, wb = oracle.sparse24_quantize_pack(w[rows].cpu(), sc[rows].cpu(), None, group_size=128)
        same_values(packed[rows].cpu(), wp, "slab vs oracle")
        same_values(bm[rows].cpu(), wb, "slab mask vs oracle")


def test_composite_plugin_round_trip():
    """the stack as ONE registry plugin with the reference's compressor interface (classmethods on local state dicts, input not mutated)"""
    from compressed_tensors_b200.compressors import BaseCompressor
    from compressed_tensors_b200.compressors.sparse.bitmask import SPARSE24_PACK_QUANTIZED
    from compressed_tensors_b200.quantization import QuantizationArgs, QuantizationScheme

    comp = BaseCompressor.get_value_from_registry(SPARSE24_PACK_QUANTIZED)
    for sym in (True, False):
        scheme = QuantizationScheme(targets=["Linear"], weights=QuantizationArgs(num_bits=4, strategy="group", group_size=128, symmetric=sym))
        g = torch.Generator().manual_seed(5)
        w = (torch.randn(256, 1024, generator=g) * 0.02).bfloat16()
        w = (w.to(DEV) * mask_creator(w.to(DEV)).to(w.dtype))
        sc, zp = _qparams(w.cpu(), "group", 128, sym)
        sd = {"weight": w, "weight_scale": sc.to(DEV), "weight_zero_point": (zp if zp is not None else torch.zeros(sc.shape, dtype=torch.int8)).to(DEV)}
        before = {k: v.clone() for k, v in sd.items()}
        out = comp.compress(sd, scheme)
        assert all(torch.equal(sd[k], before[k]) for k in sd) and set(sd) == set(before), "input state dict was mutated"
        assert set(out) == set(comp.compression_param_names(scheme)), sorted(out)
        wp, wb = oracle.sparse24_quantize_pack(w.cpu(), sc, zp, group_size=128)
        same_values(out["weight_packed"].cpu(), wp, "plugin packed")
        same_values(out["bitmask"].cpu(), wb, "plugin mask")
        back = comp.decompress(out, scheme)
        assert "weight_packed" not in back and back["weight"].dtype == torch.bfloat16
        same(back["weight"].cpu(), oracle.sparse24_unpack_dequantize(wp, wb, sc, zp, 4, w.shape), "plugin decompress")
        meta = comp.compress({k: torch.empty_like(v, device="meta") for k, v in sd.items()}, scheme)
        assert {k: (tuple(v.shape), v.dtype) for k, v in meta.items() if k != "weight_shape"} == {k: (tuple(v.shape), v.dtype) for k, v in out.items() if k != "weight_shape"}

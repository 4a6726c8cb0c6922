"""GPU tests: multi-tensor pack/unpack launches, 2:4 semi-structured (marlin-24 remnant) kernels."""
import pytest
import torch

import oracle
from compressed_tensors_b200 import _native as N
from compressed_tensors_b200 import ops
from compressed_tensors_b200.utils.permutations_24 import get_permutations_24
from compressed_tensors_b200.utils.semi_structured_conversions import (
    mask_creator,
    sparse_semi_structured_from_dense_cutlass,
    sparse_semi_structured_to_dense_cutlass,
)
from tests.golden import load
from tests.util import same, same_values

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_batched_pack_unpack_equals_per_tensor():
    g = torch.Generator().manual_seed(3)
    shapes = [(1024, 4096), (64, 128), (512, 14336), (7, 64), (33, 100)]   # the last one takes the generic kernel
    codes = [torch.randint(-8, 8, s, dtype=torch.int8, generator=g).to(DEV) for s in shapes]
    want = [ops.pack_to_int32(c, 4) for c in codes]
    outs = [torch.zeros_like(w) for w in want]
    descs = []
    for c in codes:
        d = N.QuantDesc()
        d.rows, d.cols, d.num_bits = c.shape[0], c.shape[1], 4
        descs.append(d)
    launches = N.launch_count()
    ops.batched(N.OP_PACK_INT32, [(d, c, None, None, o) for d, c, o in zip(descs, codes, outs)])
    assert N.launch_count() - launches == 2, "4 flat tensors in one launch + 1 generic launch"
    for o, w, c in zip(outs, want, codes):
        same_values(o, w, "batched pack")
        same_values(o.cpu(), oracle.pack_to_int32(c.cpu(), 4), "batched pack vs oracle")
    back = [torch.zeros_like(c) for c in codes]
    ops.batched(N.OP_UNPACK_INT32, [(d, o, None, None, b) for d, o, b in zip(descs, outs, back)])
    for b, c in zip(back, codes):
        same_values(b, c, "batched unpack")


def test_semi_structured_golden_gpu():
    sp = load("sparse")
    for c in sp["semi"]:
        sparse, meta = sparse_semi_structured_from_dense_cutlass(c["dense"].to(DEV))
        assert sparse.is_cuda and meta.dtype == c["meta"].dtype and meta.shape == c["meta"].shape
        same(sparse.cpu(), c["sparse"], f"semi sparse {c['dense'].dtype}")
        same_values(meta.cpu(), c["meta"], f"semi meta {c['dense'].dtype}")
        back = sparse_semi_structured_to_dense_cutlass(c["sparse"].to(DEV), c["meta"].to(DEV)).cpu()
        same(back, oracle.semi_structured_to_dense(c["sparse"], c["meta"]), "semi to_dense vs oracle")
        if back.dtype != torch.float32:   # fp32: the reference quiets fp16-sNaN-looking halves (see oracle test)
            same(back, c["back"], "semi to_dense vs reference")


@pytest.mark.parametrize("dtype,shape", [(torch.bfloat16, (256, 4096)), (torch.float16, (128, 512)), (torch.int8, (64, 1024)), (torch.float32, (64, 64))])
def test_semi_structured_vs_oracle(dtype, shape):
    g = torch.Generator().manual_seed(shape[1])
    x = torch.randn(shape, generator=g) * 10
    x = x.round().clamp(-127, 127).to(dtype) if dtype == torch.int8 else x.to(dtype)
    pruned = (x.float() * mask_creator(x.to(DEV)).cpu()).to(dtype)
    sparse, meta = sparse_semi_structured_from_dense_cutlass(pruned.to(DEV))
    ws, wm = oracle.semi_structured_from_dense(pruned)
    same(sparse.cpu(), ws, "sparse")
    same_values(meta.cpu(), wm, "meta")
    dense = sparse_semi_structured_to_dense_cutlass(sparse, meta).cpu()
    same(dense, oracle.semi_structured_to_dense(ws, wm), "dense")
    if dtype != torch.float32:
        same_values(dense, pruned, "round trip of a 2:4 matrix")


def test_mask_creator_golden_gpu():
    for c in load("sparse")["mask_creator"]:
        m = mask_creator(c["x"].float().to(DEV))
        assert m.dtype == torch.float32 and m.shape == c["x"].shape
        same_values(m.cpu().bool(), c["mask"], f"mask_creator {c['x'].dtype}")
    with pytest.raises(ValueError, match="can't be evenly divided"):
        mask_creator(torch.zeros(3, 3, device=DEV))


def test_semi_structured_errors():
    with pytest.raises(RuntimeError, match="2-dimensional"):
        sparse_semi_structured_from_dense_cutlass(torch.zeros(4, device=DEV))
    with pytest.raises(RuntimeError, match="divisible by 32"):
        sparse_semi_structured_from_dense_cutlass(torch.zeros(48, 64, dtype=torch.half, device=DEV))
    with pytest.raises(RuntimeError, match="divisible by 16"):
        sparse_semi_structured_from_dense_cutlass(torch.zeros(64, 24, dtype=torch.half, device=DEV))
    with pytest.raises(RuntimeError, match="Invalid datatype"):
        sparse_semi_structured_from_dense_cutlass(torch.zeros(64, 64, dtype=torch.float64, device=DEV))
    assert get_permutations_24(4)[0].numel() == 1024


# --------------------------------------------------------------------------------------------
# BLOCK strategy / one-row group scales on the streaming fast path (2-D scale addressing)
# --------------------------------------------------------------------------------------------
from types import SimpleNamespace  # noqa: E402


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape,block", [((1000, 4096), (128, 128)), ((256, 1024), (64, 256)), ((130, 520), (32, 8))])
def test_block_quant_fast_path_vs_oracle(dtype, shape, block):
    g = torch.Generator().manual_seed(shape[0])
    w = (torch.randn(shape, generator=g) * 0.02).to(dtype)
    bh, bw = block
    nr, nc = -(-shape[0] // bh), -(-shape[1] // bw)
    wp = torch.zeros(nr * bh, nc * bw)
    wp[: shape[0], : shape[1]] = w.float()
    scale = (wp.reshape(nr, bh, nc, bw).abs().amax((1, 3)) / 448).to(dtype)
    scale[scale == 0] = 1.0
    a = SimpleNamespace(strategy="block", group_size=None, block_structure=[bh, bw], num_bits=8, type="float", symmetric=True)
    kw = dict(strategy="block", block_structure=[bh, bw], qtype="float", num_bits=8)
    want = oracle.quantize(w, scale, None, dtype=torch.float8_e4m3fn, **kw)
    before = N.launch_count()
    got = ops.quantize(w.to(DEV), scale.to(DEV), None, a, dtype=torch.float8_e4m3fn)
    assert N.launch_count() == before + 1
    same_values(got.cpu().view(torch.uint8), want.view(torch.uint8), "block fp8 quantize")
    same(ops.dequantize(got, scale.to(DEV), None, args=a).cpu(), oracle.dequantize(want, scale, None, strategy="block", block_structure=[bh, bw]), "block dequantize")
    same(ops.fake_quantize(w.to(DEV), scale.to(DEV), None, a).cpu(), oracle.fake_quantize(w, scale, None, **kw), "block fake_quantize")
    # int4 + pack with block scales
    a4 = SimpleNamespace(strategy="block", group_size=None, block_structure=[bh, bw], num_bits=4, type="int", symmetric=True)
    s4 = (wp.reshape(nr, bh, nc, bw).abs().amax((1, 3)) / 7.5).to(dtype)
    s4[s4 == 0] = 1.0
    want4 = oracle.pack_to_int32(oracle.quantize(w, s4, None, strategy="block", block_structure=[bh, bw], num_bits=4, dtype=torch.int8), 4)
    same_values(ops.quantize_pack(w.to(DEV), s4.to(DEV), None, a4).cpu(), want4, "block int4 quantize_pack")


def test_one_row_group_scale_fast_path():
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(512, 1024, generator=g) * 0.02).bfloat16()
    scale = (w.float().unflatten(-1, (-1, 64)).abs().amax((0, 2)) / 7.5).bfloat16().reshape(1, -1)
    a = SimpleNamespace(strategy="group", group_size=64, block_structure=None, num_bits=4, type="int", symmetric=True)
    want = oracle.quantize(w, scale, None, strategy="group", group_size=64, num_bits=4, dtype=torch.int8)
    same_values(ops.quantize(w.to(DEV), scale.to(DEV), None, a, dtype=torch.int8).cpu(), want, "one-row group scale")
    same_values(ops.quantize_pack(w.to(DEV), scale.to(DEV), None, a).cpu(), oracle.pack_to_int32(want, 4), "one-row group scale, packed")


# ---- sparse24 vectorised path (bf16 / fp16, cols % 8 == 0): ties, signed zeros, arbitrary masks ---------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_sparse24_vectorised_ties_and_zeros(dtype):
    """masks produced by compress always hold exactly two bits per quad; other masks are outside the format"""
    import oracle
    from tests.util import same

    g = torch.Generator().manual_seed(9)
    x = (torch.randint(-3, 4, (256, 1024), generator=g).float() * 0.5).to(dtype)    # few distinct magnitudes: lots of ties and zeros
    x[0, :8] = torch.tensor([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1.0, -1.0]).to(dtype)
    vals, bm = oracle.sparse24_compress(x)
    gv, gb = ops.sparse24_compress(x.to(DEV))
    same(gb.cpu(), bm, "sparse24 bitmask")
    same(gv.cpu(), vals, "sparse24 values")
    same(ops.sparse24_decompress(gv, gb, x.shape).cpu(), oracle.sparse24_decompress(vals, bm, x.shape), "sparse24 decompress")


def test_sparse24_full_size_round_trip():
    x = (torch.randn(4096, 14336, device=DEV) * 0.02).to(torch.bfloat16)
    v, m = ops.sparse24_compress(x)
    assert v.shape == (4096, 7168) and m.shape == (4096, 1792)
    dense = ops.sparse24_decompress(v, m, x.shape)
    # every quad keeps exactly two elements, and they are the two of largest magnitude
    q, dq = x.view(-1, 4).float().abs(), dense.view(-1, 4)
    assert torch.equal((dq != 0).sum(-1) + ((dq == 0) & (x.view(-1, 4) == 0) & (torch.arange(4, device=DEV) < 0)).sum(-1), (dq != 0).sum(-1))
    kept = dq != 0
    assert int(kept.sum(-1).max()) <= 2
    third = q.sort(-1, descending=True).values[:, 2:3]
    assert bool(((q >= third) | ~kept).all()), "a kept element is smaller than a dropped one"
    assert torch.equal(torch.where(kept, x.view(-1, 4), torch.zeros_like(dq)), dq)
    v2, m2 = ops.sparse24_compress(dense)              # idempotent on its own output
    assert torch.equal(ops.sparse24_decompress(v2, m2, x.shape), dense)


# ---- unstructured bitmask, vectorised path (bf16 / fp16, cols % 32 == 0) ---------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,density", [((64, 4096), 0.3), ((33, 2048), 0.5), ((5, 32), 1.0), ((7, 8192), 0.0), ((16, 14336), 0.9), ((3, 2080), 0.5)])
def test_bitmask_vectorised_vs_oracle(dtype, shape, density):
    import oracle
    from tests.util import same

    g = torch.Generator().manual_seed(shape[1])
    x = (torch.randn(shape, generator=g) * 3).to(dtype)
    x[torch.rand(shape, generator=g) >= density] = 0
    if shape[0] > 2:
        x[1] = 0                    # an empty row in the middle: its offset equals the next row's
        x[2, ::7] = -0.0            # negative zeros are zeros
    vals, bm, offs = oracle.bitmask_compress(x)
    gv, gb, go = ops.bitmask_compress(x.to(DEV))
    same(gb.cpu(), bm, "bitmask")
    same(go.cpu(), offs, "row offsets")
    same(gv.cpu(), vals, "values")
    dense = ops.bitmask_decompress(gv, gb, go, x.shape)
    same(dense.cpu(), oracle.bitmask_decompress(vals, bm, x.shape), "decompress")


def test_bitmask_full_size_round_trip():
    x = (torch.randn(4096, 14336, device=DEV) * 0.02).to(torch.bfloat16)
    x = torch.where(torch.rand(x.shape, device=DEV) < 0.5, x, torch.zeros_like(x))
    v, m, o = ops.bitmask_compress(x)
    assert v.numel() == int((x != 0).sum()) and torch.equal(v, x[x != 0])
    assert torch.equal(o, torch.cat([torch.zeros(1, dtype=torch.int64, device=DEV), (x != 0).sum(-1).cumsum(0)[:-1]]))
    assert torch.equal(ops.bitmask_decompress(v, m, o, x.shape), x)

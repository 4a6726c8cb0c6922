"""
Robustness of the C ABI on the B200:
  * thread safety -- the reference's convert_checkpoint calls decompress from a thread pool
    (entrypoints/convert/convert_checkpoint.py:110-134): concurrent callers on separate CUDA streams must get the
    results of a serial run, including launches that use the dynamic tile schedule (own scratch per launch);
  * seeded randomized differential test against the CPU oracle across strategies / dtypes / shapes that straddle the
    fast-path / generic-path boundary;
  * stream semantics -- work is enqueued on the caller's current stream and nothing synchronises inside.
"""
import ctypes
import random
from concurrent.futures import ThreadPoolExecutor

import pytest
import torch

import oracle
from compressed_tensors_b200 import _native as N
from compressed_tensors_b200 import ops
from compressed_tensors_b200.quantization import QuantizationArgs
from tests.util import same

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _w4():
    return QuantizationArgs(num_bits=4, type="int", symmetric=True, strategy="group", group_size=128)


def test_concurrent_callers_on_separate_streams():
    torch.manual_seed(0)
    a = _w4()
    shapes = [(4096, 4096), (1024, 4096), (14336, 4096), (512, 1024), (4096, 14336), (96, 256)]   # large ones take the dynamic schedule
    work = []
    for i, sh in enumerate(shapes * 2):
        w = (torch.randn(sh, device=DEV) * 0.02).to(torch.bfloat16)
        s = (w.float().unflatten(-1, (-1, 128)).abs().amax(-1) / 7.5).to(torch.bfloat16)
        work.append((w, s))
    serial = []
    for w, s in work:
        p = ops.quantize_pack(w, s, None, a)
        serial.append((p, ops.unpack_dequantize(p, s, None, 4, w.shape)))
    torch.cuda.synchronize()

    def job(i):
        w, s = work[i]
        st = torch.cuda.Stream(device=DEV)
        st.wait_stream(torch.cuda.default_stream(torch.device(DEV)))
        with torch.cuda.stream(st):
            out = []
            for _ in range(3):   # several launches per thread so that calls overlap in time
                p = ops.quantize_pack(w, s, None, a)
                out = (p, ops.unpack_dequantize(p, s, None, 4, w.shape))
            st.synchronize()
        return i, out

    with ThreadPoolExecutor(max_workers=8) as ex:
        for i, (p, d) in ex.map(job, range(len(work))):
            assert torch.equal(p, serial[i][0]), f"packed words of job {i} differ under concurrency"
            assert torch.equal(d.view(torch.int16), serial[i][1].view(torch.int16)), f"dequantized output of job {i} differs under concurrency"


def test_enqueue_only_no_sync_and_stream_order():
    """a long kernel on a side stream followed by ours on the same stream must see the long kernel's output"""
    a = _w4()
    st = torch.cuda.Stream(device=DEV)
    w = torch.empty(8192, 8192, device=DEV, dtype=torch.bfloat16)
    s = torch.full((8192, 64), 0.01, device=DEV, dtype=torch.bfloat16)
    with torch.cuda.stream(st):
        for _ in range(4):
            w.normal_(0, 0.02)        # producer on the same stream: ordering is the only synchronisation
        p = ops.quantize_pack(w, s, None, a)
    st.synchronize()
    assert torch.equal(p, ops.quantize_pack(w, s, None, a))


_STRATS = ["tensor", "channel", "group", "group_row", "block", "token"]


@pytest.mark.parametrize("seed", range(40))
def test_randomized_differential_vs_oracle(seed):
    rnd = random.Random(seed)
    g = torch.Generator().manual_seed(seed)
    dt = rnd.choice([torch.bfloat16, torch.float16, torch.float32])
    sdt = dt if rnd.random() < 0.7 else rnd.choice([torch.bfloat16, torch.float16, torch.float32])
    strat = rnd.choice(_STRATS)
    qtype, bits = rnd.choice([("int", 4), ("int", 8), ("int", rnd.randint(1, 8)), ("float", 8), ("float", 4)])
    gsz = rnd.choice([16, 32, 64, 128])
    rows = rnd.choice([1, 3, 8, 33, 64, 257])
    cols = gsz * rnd.choice([1, 2, 3, 8, 17]) if strat.startswith("group") or rnd.random() < 0.5 else rnd.choice([8, 24, 40, 100, 136, 1000])
    x = (torch.randn(rows, cols, generator=g) * 10 ** rnd.uniform(-3, 1)).to(dt)
    qmax = {"int": 2 ** (bits - 1) - 0.5, "float": 448.0 if bits == 8 else 6.0}[qtype]
    kw = dict(num_bits=bits, type=qtype, symmetric=True)
    if strat == "tensor":
        a, shape = QuantizationArgs(strategy="tensor", **kw), (1,)
    elif strat in ("channel", "token"):
        a, shape = QuantizationArgs(strategy=strat, **({**kw, "dynamic": True} if strat == "token" else kw)), (rows, 1)
    elif strat == "group":
        a, shape = QuantizationArgs(strategy="group", group_size=gsz, **kw), (rows, cols // gsz)
    elif strat == "group_row":
        a, shape = QuantizationArgs(strategy="group", group_size=gsz, **kw), (1, cols // gsz)
    else:
        bh, bw = rnd.choice([(4, 8), (16, 16), (128, 128), (8, 24)])
        a, shape = QuantizationArgs(strategy="block", block_structure=[bh, bw], **kw), (-(-rows // bh), -(-cols // bw))
    s = ((torch.rand(shape, generator=g) + 0.25) * float(x.float().abs().max().clamp_min(1e-6)) / qmax).to(sdt)
    zp = None
    if qtype == "int" and rnd.random() < 0.4:
        zp = torch.randint(-(2 ** (bits - 1)), 2 ** (bits - 1), shape, generator=g).to(torch.int8)
    okw = dict(strategy=a.strategy, group_size=a.group_size, block_structure=a.block_structure, num_bits=bits, qtype=qtype)
    qdt = torch.int8 if qtype == "int" else (torch.float8_e4m3fn if bits == 8 else None)
    X, S, Z = x.to(DEV), s.to(DEV), zp.to(DEV) if zp is not None else None
    what = f"seed {seed}: {strat} {dt} scale {sdt} {qtype}{bits} {rows}x{cols} zp={zp is not None}"
    want_q = oracle.quantize(x, s, zp, dtype=qdt, **okw)
    same(ops.quantize(X, S, Z, a, dtype=qdt).cpu(), want_q, "quantize " + what)
    same(ops.fake_quantize(X, S, Z, a).cpu(), oracle.fake_quantize(x, s, zp, **okw), "fake_quantize " + what)
    same(ops.dequantize(want_q.to(DEV), S, Z, args=a).cpu(), oracle.dequantize(want_q, s, zp, strategy=a.strategy, group_size=a.group_size,
                                                                                block_structure=a.block_structure), "dequantize " + what)
    if qtype == "int":
        want_p = oracle.pack_to_int32(oracle.quantize(x, s, zp, dtype=torch.int8, **okw), bits)
        same(ops.quantize_pack(X, S, Z, a).cpu(), want_p, "quantize_pack " + what)
    elif bits == 4 and cols % 2 == 0:
        same(ops.quantize_pack_fp4(X, S, Z, a).cpu(), oracle.pack_fp4_to_uint8(want_q), "quantize_pack_fp4 " + what)


def test_per_tensor_launches_are_cuda_graph_capturable():
    """single-tensor entry points only enqueue (kernel + stream-ordered scratch from the library's pool): a sequence of them can be
    captured once and replayed -- the way a caller would hide launch latency for the 224 small launches of a per-module loop"""
    a = _w4()
    ws = [(torch.randn(sh, device=DEV) * 0.02).to(torch.bfloat16) for sh in ((1024, 4096), (4096, 4096), (14336, 4096))]
    ss = [(w.float().unflatten(-1, (-1, 128)).abs().amax(-1) / 7.5).to(torch.bfloat16) for w in ws]
    want = [ops.quantize_pack(w, s, None, a) for w, s in zip(ws, ss)]
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for w, s in zip(ws, ss):          # warm-up on the capture stream
            ops.quantize_pack(w, s, None, a)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            outs = [ops.quantize_pack(w, s, None, a) for w, s in zip(ws, ss)]
    for w in ws:
        w.mul_(-1.0)                      # new inputs in the captured buffers
    graph.replay()
    torch.cuda.synchronize()
    for o, w, s in zip(outs, ws, ss):
        assert torch.equal(o, ops.quantize_pack(w, s, None, a))
    assert not torch.equal(outs[0], want[0])


def test_multi_tensor_launch_is_cuda_graph_capturable():
    """a whole-model ct_batched call (job table + dynamic tile counter) captured once and replayed on new data in the same buffers"""
    a = _w4()
    shapes = [(1024, 4096), (4096, 4096), (14336, 4096)] * 14          # 42 tensors: two upload kernels (32 + 10 jobs), dynamic schedule
    ws = [(torch.randn(sh, device=DEV) * 0.02).to(torch.bfloat16) for sh in shapes]
    ss = [(w.float().unflatten(-1, (-1, 128)).abs().amax(-1) / 7.5).to(torch.bfloat16) for w in ws]
    outs = [torch.zeros(w.shape[0], w.shape[1] // 8, dtype=torch.int32, device=DEV) for w in ws]
    probs = []
    for w, s, o in zip(ws, ss, outs):
        p = ops._resolve(w, s, None, a, None)
        probs.append((ops._desc(p, w.dtype, s.dtype, None, torch.bfloat16, torch.int8, None, N.Q_INT, 4), w, s, None, o))
    plan = ops.BatchedPlan(N.OP_QUANTIZE_PACK, probs)
    plan.run()
    torch.cuda.synchronize()
    first = [o.clone() for o in outs]
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        plan.run()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            plan.run()
    for w in ws:
        w.mul_(-1.0)
    for o in outs:
        o.zero_()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    for o, w, s, f in zip(outs[::5], ws[::5], ss[::5], first[::5]):
        assert torch.equal(o, ops.quantize_pack(w, s, None, a)) and not torch.equal(o, f)


# ---- error behaviour of the newer entry points: the reference's exception types, never a silent fallback ---------------------------
def test_error_paths_raise_like_the_reference():
    from compressed_tensors_b200.compressors import NVFP4PackedCompressor
    from compressed_tensors_b200.quantization import QuantizationScheme

    x = torch.randn(8, 30, device=DEV).to(torch.bfloat16)
    with pytest.raises(ValueError, match="even number of columns"):
        ops.pack_fp4_to_uint8(torch.zeros(3, 7, device=DEV))
    with pytest.raises(ValueError):
        ops.unpack_fp4_from_uint8(torch.zeros(3, 4, dtype=torch.uint8, device=DEV), 3, 6)          # 12 bytes do not hold 3 x 6
    with pytest.raises(ValueError, match="divisble"):
        ops.quantize(x, torch.ones(8, 2, device=DEV, dtype=torch.bfloat16), None, QuantizationArgs(num_bits=4, type="float", strategy="group", group_size=16))
    with pytest.raises(ValueError):
        ops.awq_repack(torch.zeros(8, 4, device=DEV))                                             # not int32
    with pytest.raises(ValueError, match="NVFP4 args"):
        ops.observe_quantize_pack_nvfp4(x, _w4())
    with pytest.raises(NotImplementedError):
        ops.cast_to_fp4(torch.zeros(4, dtype=torch.int32, device=DEV))
    # the raw ABI refuses what the fused kernels do not cover instead of running something else
    d = N.QuantDesc()
    d.rows, d.cols, d.rdiv, d.cdiv, d.s_row_stride = 4, 24, 1, N.INF, 1
    d.x_dtype = d.scale_dtype = d.compute_dtype = N.DT[torch.float32]
    d.zp_dtype, d.q_dtype, d.out_dtype, d.qtype, d.num_bits = N.DT_NONE, N.DT[torch.int8], N.DT_NONE, N.Q_INT, 8
    xf = torch.zeros(4, 24, device=DEV)
    s = torch.zeros(4, 1, device=DEV)
    o = torch.zeros(4, 24, dtype=torch.int8, device=DEV)
    rc = N.lib().ct_observe_quantize_channel(ctypes.byref(d), N.ptr(xf), N.ptr(s), None, N.ptr(o), 0, N.stream_ptr(0))
    assert rc == N.CT_E_UNSUPPORTED and "fused channel observer" in N.last_error()
    # ... while the Python op falls back to observer + quantize as separate GPU kernels and still matches the oracle
    a = QuantizationArgs(num_bits=8, type="int", symmetric=True, strategy="channel")
    xr = torch.randn(4, 24, device=DEV)
    q, sc, _ = ops.observe_quantize(xr, a)
    from oracle.qparams import calculate_qparams as oq
    ws, _ = oq(xr.cpu().amin(-1, keepdim=True), xr.cpu().amax(-1, keepdim=True), num_bits=8, qtype="int", symmetric=True)
    assert torch.equal(sc.cpu(), ws) and torch.equal(q.cpu(), oracle.quantize(xr.cpu(), ws, None, strategy="channel", num_bits=8, dtype=torch.int8))
    # a compressor fed the wrong scheme fails loudly as well
    with pytest.raises(ValueError):
        NVFP4PackedCompressor.compress({"weight": x, "weight_scale": torch.ones(8, 2, device=DEV, dtype=torch.bfloat16)},
                                       QuantizationScheme(targets=["Linear"], weights=_w4()))

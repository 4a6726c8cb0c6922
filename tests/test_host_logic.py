"""
CPU-only checks of the product's host logic (no kernel runs here):
  * libct_b200.so loads and exports every symbol include/ct_b200.h declares
  * the strategy -> scale-addressing resolution of compressed_tensors_b200.ops feeds the
    ORACLE arithmetic and must then reproduce the reference's golden outputs
  * error behaviour mirrors the reference
"""
import ctypes
import os
import re
from types import SimpleNamespace

import pytest
import torch

import oracle
from compressed_tensors_b200 import _native as N
from compressed_tensors_b200 import ops
from tests.golden import load
from tests.util import bits_equal, diff_report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ct_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in ct_b200.h but not exported"
    assert declared == set(N.EXPORTED_SYMBOLS), declared ^ set(N.EXPORTED_SYMBOLS)
    assert b"sm_100a" in N.lib().ct_version()


def test_every_declared_symbol_cites_the_reference_and_is_mapped_in_integration_md():
    """include/ct_b200.h says which reference interface each entry point replaces (file:line); INTEGRATION.md maps every symbol"""
    hdr = open(os.path.join(ROOT, "include", "ct_b200.h")).read()
    declared = set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(n for n in declared if n not in integ)
    assert not missing, f"INTEGRATION.md does not mention {missing}"
    assert len(re.findall(r"\.py:\d+", hdr)) >= 20, "the header should cite reference file:line for its entry points"


def test_no_silent_cpu_fallback():
    """without a GPU a CUDA entry point (device >= 0) refuses; host code runs only when asked for explicitly (device = -1 /
    ImplBackend's eager body), and only for the seven per-tensor hot-path ops -- everything else still raises"""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert N.lib().ct_device_count() == 0
    q = torch.zeros(4, 32, dtype=torch.int8)
    out = torch.zeros(4, 4, dtype=torch.int32)
    rc = N.lib().ct_pack_int32(N.ptr(q), N.ptr(out), 4, 32, 4, 1, 0, None)
    assert rc == N.CT_E_NODEV
    assert "no CPU path" in N.last_error() or "CUDA" in N.last_error()
    from compressed_tensors_b200.utils import ImplBackend

    with pytest.raises(N.NativeLibraryError):
        ImplBackend.call("pack_to_int32_sm100", q, 4)
    for fn, args in ((ops.sparse24_compress, (torch.zeros(4, 32, dtype=torch.bfloat16),)), (ops.cast_to_fp4, (torch.zeros(4, 32),)),
                     (ops.pack_bitmasks, (torch.zeros(4, 32, dtype=torch.bool),)), (ops.bitmask_compress, (torch.zeros(4, 32, dtype=torch.bfloat16),))):
        with pytest.raises(N.NativeLibraryError):
            fn(*args)


def test_only_test_infrastructure_touches_the_oracle():
    """oracle/ is the checker: nothing in the product package, the drop-in alias or tools/ may import it; bench.py only inside its CPU legs
    (the cpu-worker subprocess of cpu_baseline / the reference arm) and as the checker of timed outputs, __graft_entry__ only inside build() / smoke()"""
    pat = re.compile(r"^\s*(import oracle\b|from oracle\b)", re.M)
    for top in ("compressed_tensors_b200", "compat", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    assert not pat.search(src) and "libct_oracle" not in src and "orc_" not in src, os.path.join(dirpath, f)
    bench = open(os.path.join(ROOT, "bench.py")).read()
    for m in pat.finditer(bench):
        fn = re.findall(r"^def (\w+)", bench[: m.start()], flags=re.M)[-1]
        # the CPU legs (timed as the baseline) and the two places where it is the CHECKER of what the GPU arm wrote
        assert fn in ("run_cpu_worker", "oracle_compress_layer", "run_cfg5_70b_sharded", "verify_timed_outputs"), f"bench.py imports the oracle inside {fn}()"
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    for m in pat.finditer(entry):
        assert re.findall(r"^def (\w+)", entry[: m.start()], flags=re.M)[-1] in ("build", "smoke")


def test_scripts_compile():
    """bench.py, __graft_entry__.py, tools/ and the reference-compat scripts at least parse (they only run on the GPU box / in the build container)"""
    import glob

    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "reference_compat", "*.py"))
    files += glob.glob(os.path.join(ROOT, "tests", "golden", "*.py"))
    for f in files:
        compile(open(f).read(), f, "exec")       # syntax only; nothing is written or executed


def test_error_messages_match_reference():
    with pytest.raises(ValueError, match="Tensor must be quantized to torch.int8 before packing"):
        ops.pack_to_int32(torch.zeros(2, 2, dtype=torch.int32), 4)
    with pytest.raises(ValueError, match=r"Packing is only supported for num_bits in \[1, 8\], got 9"):
        ops.pack_to_int32(torch.zeros(2, 2, dtype=torch.int8), 9)
    with pytest.raises(ValueError, match="Aborting unpack"):
        ops.unpack_from_int32(torch.zeros(2, 2, dtype=torch.int8), 4, (2, 2))
    with pytest.raises(ValueError, match=r"Unpacking is only supported for num_bits in \[1, 8\], got 0"):
        ops.unpack_from_int32(torch.zeros(2, 2, dtype=torch.int32), 0, (2, 2))
    a = SimpleNamespace(strategy="group", group_size=128, num_bits=4, type="int", block_structure=None)
    with pytest.raises(ValueError, match="tensor column shape must be divisble"):
        ops.quantize(torch.zeros(4, 200), torch.ones(4, 2), None, a)


def test_meta_tensors_only_compute_shapes():
    a = SimpleNamespace(strategy="group", group_size=128, num_bits=4, type="int", block_structure=None)
    x = torch.empty(64, 512, dtype=torch.bfloat16, device="meta")
    s = torch.empty(64, 4, dtype=torch.bfloat16, device="meta")
    q = ops.quantize(x, s, None, a, dtype=torch.int8)
    assert q.device.type == "meta" and q.dtype == torch.int8 and q.shape == x.shape
    p = ops.quantize_pack(x, s, None, a)
    assert p.device.type == "meta" and p.dtype == torch.int32 and p.shape == (64, 64)
    assert ops.pack_to_int32(torch.empty(8, 100, dtype=torch.int8, device="meta"), 3).shape == (8, 10)
    assert ops.fake_quantize(x, s, None, a).dtype == torch.bfloat16


_Q = load("quant")


def _oracle_with_product_addressing(c, x):
    """run the oracle's C arithmetic with the addressing computed by the PRODUCT's host logic"""
    a = SimpleNamespace(**c["args"])
    p = ops._resolve(x, c["scale"], c["zp"], a, c["g_idx"])
    cd = torch.result_type(x, c["scale"])
    x2 = x.reshape(p.rows, p.cols).contiguous()
    out = torch.empty(p.rows, p.cols, dtype=c["q"].dtype)
    L = oracle.lib()
    sc = p.scale.contiguous()
    zp = p.zp.contiguous() if p.zp is not None else None
    gi = p.g_idx.to(torch.int32).contiguous() if p.g_idx is not None else None
    rc = L.orc_quantize(oracle._p(x2), oracle.DT[x2.dtype], oracle._p(sc), oracle.DT[sc.dtype], oracle._p(zp),
                        oracle.DT[zp.dtype] if zp is not None else -1, oracle._p(gi), oracle._p(out), oracle.DT[out.dtype],
                        ctypes.c_int64(p.rows), ctypes.c_int64(p.cols), ctypes.c_int64(p.rdiv), ctypes.c_int64(p.cdiv),
                        ctypes.c_int64(p.srs), oracle.DT[cd], 0 if a.type == "int" else 1, a.num_bits)
    assert rc == 0
    return out.reshape(x.shape)


@pytest.mark.parametrize("i", range(len(_Q["cases"])))
def test_product_addressing_reproduces_golden(i):
    c = _Q["cases"][i]
    x = _Q["x"][c["x"]] if isinstance(c["x"], str) else c["x"]
    got = _oracle_with_product_addressing(c, x)
    assert bits_equal(got, c["q"]), diff_report(got, c["q"])


def test_dequant_strategy_inference():
    x = torch.empty(8, 64, dtype=torch.int8)
    assert ops._infer_dequant_args(x, torch.ones(1)).strategy == "tensor"
    assert ops._infer_dequant_args(x, torch.ones(())).strategy == "tensor"
    assert ops._infer_dequant_args(x, torch.ones(8, 1)).strategy == "channel"
    g = ops._infer_dequant_args(x, torch.ones(8, 4))
    assert g.strategy == "group" and g.group_size == 16
    g = ops._infer_dequant_args(x, torch.ones(1, 2))
    assert g.strategy == "group" and g.group_size == 32
    b = ops._infer_dequant_args(x, torch.ones(2, 2))
    assert b.strategy == "block" and b.block_structure == [4, 32]
    with pytest.raises(ValueError, match="Could not infer"):
        ops._infer_dequant_args(x, torch.ones(2, 2, 2))


def test_generate_gparam_rounds_like_the_reference():
    """`float / tensor` is reciprocal-then-multiply, each rounded to the tensor's dtype (reference helpers.py:328).  Known answers
    produced by the reference in the build container (tests/reference_compat/fuzz_host_mirror.py found the 1541 / 1540 split)."""
    import torch

    from compressed_tensors_b200.quantization.utils import generate_gparam

    cases = [(torch.float16, -1.7451171875, 1.68359375, 1541.0), (torch.float16, -484.0, 131.625, 5.55078125),
             (torch.bfloat16, -0.59765625, 13.5, 200.0), (torch.float16, -224.125, 207.25, 12.0)]
    for dt, lo, hi, want in cases:
        g = generate_gparam(torch.tensor(lo, dtype=dt), torch.tensor(hi, dtype=dt))
        assert g.dtype == torch.float32 and g.shape == (1,) and g.item() == want, (dt, lo, hi, g)


def test_recouple_bucket_layout():
    """the all-gather recouple's byte layout (distributed/module_parallel.py::_layout): every owner's tensors in module order, cut
    into buckets of <= _BUCKET bytes per owner (a larger tensor gets a bucket of its own), offsets 256-byte aligned, slot = the
    largest owner's fill; identical for every rank by construction (pure function of shapes / dtypes / owners)"""
    import torch

    import compressed_tensors_b200.distributed.module_parallel as mp

    class M(torch.nn.Module):
        def __init__(self, n):
            super().__init__()
            self.weight_packed = torch.nn.Parameter(torch.empty(n, dtype=torch.int32, device="meta"), requires_grad=False)
            self.weight_scale = torch.nn.Parameter(torch.empty(max(n // 16, 1), dtype=torch.bfloat16, device="meta"), requires_grad=False)

    sizes = [1000, 70000, 10, 300000, 5000, 64, 90000, 1]
    mods = [M(n) for n in sizes]
    owner = {m: i % 3 for i, m in enumerate(mods)}
    old = mp._BUCKET
    try:
        mp._BUCKET = 256 << 10
        buckets, total = mp._layout(mods, owner, 3)
    finally:
        mp._BUCKET = old
    assert total == sum(slot for slot, _ in buckets) * 3
    seen = []
    for slot, entries in buckets:
        assert slot > 0 and slot % mp._ALIGN == 0
        for r in range(3):
            end = 0
            for m, name, off, nbytes, shape, dtype in entries[r]:
                assert owner[m] == r and off % mp._ALIGN == 0 and off >= end and off + nbytes <= slot
                assert nbytes == getattr(m, name).numel() * getattr(m, name).element_size()
                end = off + nbytes
                seen.append((id(m), name))
            # a bucket holds more than _BUCKET bytes of one owner only when it is a single oversized tensor
            assert end <= (256 << 10) or len(entries[r]) == 1
    assert sorted(seen) == sorted((id(m), n) for m in mods for n in ("weight_packed", "weight_scale"))   # every tensor exactly once
    assert len(buckets) > 1                                                                              # 1.2 MB tensor: several buckets

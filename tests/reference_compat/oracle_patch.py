"""
pytest plugin, TEST INFRASTRUCTURE ONLY (see tests/test_reference_suite_host.py).

In the build container there is no GPU, and on the GPU box there is no /root/reference.  To still run the REFERENCE'S OWN test files
against this package's host mirror (compressor classes, registry, schema, lifecycle glue, converters), this plugin rebinds the
tensor-level front end `compressed_tensors.ops` (the drop-in alias of compressed_tensors_b200.ops) to the CPU oracle for the duration
of a pytest run.  What that proves: the Python layers above the C ABI behave like the reference's (same keys, shapes, dtypes,
exceptions, values through the oracle).  What it does NOT prove: anything about the CUDA kernels -- `pytest -m gpu` does that,
through the real ABI.  Nothing under compressed_tensors_b200/ knows about this file.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch


def pytest_configure(config):
    apply("compressed_tensors")
    _stub_reference_test_harness()


def apply(package: str = "compressed_tensors"):
    """rebind `<package>.ops` (the drop-in alias by default; tests/reference_compat/fuzz_compressors.py passes "compressed_tensors_b200",
    because there the name `compressed_tensors` is the reference itself) to the CPU oracle"""
    import importlib

    import oracle

    ops = importlib.import_module(package + ".ops")

    _PATCHED = ("quantize", "dequantize", "fake_quantize", "quantize_pack", "unpack_dequantize", "pack_to_int32", "unpack_from_int32", "cast_to_fp4",
                "pack_fp4_to_uint8", "unpack_fp4_from_uint8", "compress_mx_scale", "decompress_mx_scale", "quantize_pack_fp4", "unpack_dequantize_fp4",
                "awq_repack", "awq_repack_zeros", "dequantize_block_fp8", "pack_bitmasks", "unpack_bitmasks")
    _ORIGINAL = {name: getattr(ops, name) for name in _PATCHED}

    def kw(args):
        s = getattr(args, "strategy", None); s = getattr(s, "value", s)
        t = getattr(args, "type", "int"); t = getattr(t, "value", t)
        return dict(strategy=s, group_size=getattr(args, "group_size", None), block_structure=getattr(args, "block_structure", None),
                    num_bits=getattr(args, "num_bits", 8), qtype=t)

    def quantize(x, scale, zero_point, args, dtype=None, g_idx=None, global_scale=None):
        return oracle.quantize(x, scale, zero_point, dtype=dtype, g_idx=g_idx, global_scale=global_scale, **kw(args))

    def dequantize(x_q, scale, zero_point=None, args=None, dtype=None, g_idx=None, global_scale=None):
        k = kw(args) if args is not None else {}
        k.pop("num_bits", None); k.pop("qtype", None)
        return oracle.dequantize(x_q, scale, zero_point, dtype=dtype, g_idx=g_idx, global_scale=global_scale, **({} if args is None else k))

    def fake_quantize(x, scale, zero_point, args, g_idx=None, global_scale=None):
        return oracle.fake_quantize(x, scale, zero_point, g_idx=g_idx, global_scale=global_scale, **kw(args))

    def quantize_pack(x, scale, zero_point, args, g_idx=None, global_scale=None):
        if x.ndim > 2:
            return torch.stack([quantize_pack(x[i], scale[i] if scale.ndim == x.ndim else scale,
                                              zero_point[i] if (zero_point is not None and zero_point.ndim == x.ndim) else zero_point, args, g_idx, global_scale) for i in range(x.shape[0])])
        return oracle.pack_to_int32(quantize(x, scale, zero_point, args, dtype=torch.int8, g_idx=g_idx, global_scale=global_scale), args.num_bits)

    def unpack_dequantize(packed, scale, zero_point, num_bits, shape, g_idx=None, dtype=None):
        if packed.ndim > 2:
            return torch.stack([unpack_dequantize(packed[i], scale[i] if scale.ndim == packed.ndim else scale,
                                                  zero_point[i] if (zero_point is not None and zero_point.ndim == packed.ndim) else zero_point, num_bits, shape[1:], g_idx, dtype) for i in range(packed.shape[0])])
        q = oracle.unpack_from_int32(packed, num_bits, shape)
        return oracle.dequantize(q, scale, zero_point, g_idx=g_idx, dtype=dtype)

    ops.quantize, ops.dequantize, ops.fake_quantize = quantize, dequantize, fake_quantize
    ops.quantize_pack, ops.unpack_dequantize = quantize_pack, unpack_dequantize
    ops.pack_to_int32 = lambda v, b, packed_dim=1: oracle.pack_to_int32(v, b, packed_dim)
    ops.unpack_from_int32 = lambda v, b, shape, packed_dim=1: oracle.unpack_from_int32(v, b, shape, packed_dim)
    ops.cast_to_fp4 = oracle.cast_to_fp4
    ops.pack_fp4_to_uint8 = oracle.pack_fp4_to_uint8
    ops.unpack_fp4_from_uint8 = oracle.unpack_fp4_from_uint8
    ops.compress_mx_scale = oracle.compress_mx_scale
    ops.decompress_mx_scale = oracle.decompress_mx_scale
    ops.quantize_pack_fp4 = lambda x, s, z, a, g_idx=None, global_scale=None: oracle.pack_fp4_to_uint8(quantize(x, s, z, a, g_idx=g_idx, global_scale=global_scale))

    def unpack_dequantize_fp4(packed, scale, global_scale=None, dtype=torch.bfloat16, stored_scale=None):
        m, n = packed.shape[0], packed.shape[1] * 2
        if stored_scale == "fp8":
            scale = scale.to(dtype)
        elif stored_scale == "e8m0":
            scale = oracle.decompress_mx_scale(scale).to(dtype)
        return oracle.dequantize(oracle.unpack_fp4_from_uint8(packed, m, n, dtype), scale, None, global_scale=global_scale, dtype=dtype)

    ops.unpack_dequantize_fp4 = unpack_dequantize_fp4
    ops.awq_repack = oracle.awq_repack
    ops.awq_repack_zeros = oracle.awq_repack_zeros
    ops.dequantize_block_fp8 = lambda w, s, block, dtype=torch.bfloat16: oracle.dequantize_block_fp8(w, s, block, dtype)
    ops.pack_bitmasks = lambda b: oracle.pack_bitmasks(b)
    ops.unpack_bitmasks = lambda p, shape: oracle.unpack_bitmasks(p, shape)
    # meta tensors (the non-owner ranks of the distributed path, shape-only validation in the converters) never reach a kernel in the
    # product either: the front end answers with an empty meta tensor of the right shape.  Keep that behaviour under the patch.
    import functools

    def meta_aware(name, patched):
        original = _ORIGINAL[name]

        @functools.wraps(patched)
        def call(*args, **kwargs):
            if any(isinstance(a, torch.Tensor) and a.device.type == "meta" for a in list(args) + list(kwargs.values())):
                return original(*args, **kwargs)
            return patched(*args, **kwargs)

        return call

    for _name in _PATCHED:
        setattr(ops, _name, meta_aware(_name, getattr(ops, _name)))

    # modules that bound the names at import time
    for name, mod in list(sys.modules.items()):
        if name.startswith(package + ".") and mod is not None and mod is not ops:
            for fn in ("pack_to_int32", "unpack_from_int32", "pack_fp4_to_uint8", "unpack_fp4_from_uint8", "pack_bitmasks", "unpack_bitmasks"):
                if hasattr(mod, fn) and getattr(getattr(mod, fn), "__module__", "") == package + ".ops":
                    setattr(mod, fn, getattr(ops, fn))


def _stub_reference_test_harness():

    # The reference's ModelCompressor / lifecycle test files import a `torchrun` decorator from tests/test_offload/conftest.py,
    # whose module body needs torch.accelerator (a GPU) and the offload subsystem (out of scope).  The tests
    # that USE the decorator are `requires_gpu(2)` and skip here anyway; a stand-in module lets the rest of those files import.
    import types

    import pytest

    stub = types.ModuleType("tests.test_offload.conftest")

    def torchrun(world_size=1, init_dist=False):
        return lambda fn: pytest.mark.skip(reason="torchrun harness of the reference's offload tests is not mirrored")(fn)

    stub.torchrun = torchrun
    sys.modules.setdefault("tests.test_offload.conftest", stub)

"""
oracle -- TEST INFRASTRUCTURE ONLY (see oracle/ct_oracle.c header).

ctypes front-end of the plain-C CPU restatement.  Takes and returns CPU torch
tensors so that the parity tests read like the reference's own tests.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (compressed_tensors_b200/) never does.

The strategy -> scale-addressing mapping and the compute-dtype rule are
restated here independently of the product so that a bug in the product's host
logic cannot hide in a shared helper:
  * strategies: quantization/lifecycle/forward.py:184-241,
    forward_helpers.py:62-177 (reference paths under src/compressed_tensors)
  * compute dtype = torch type promotion of `x / scale`
    (forward_helpers.py:538), SURVEY.md Appendix B1
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libct_oracle.so")
_SRC = os.path.join(_HERE, "ct_oracle_all.c")  # includes ct_oracle.c and ct_oracle_qparams.c
_PARTS = [os.path.join(_HERE, f) for f in ("ct_oracle_all.c", "ct_oracle.c", "ct_oracle_qparams.c", "ct_oracle_fp4.c", "ct_oracle_convert.c")]

INT64_MAX = (1 << 63) - 1

DT = {
    torch.float32: 0,
    torch.float16: 1,
    torch.bfloat16: 2,
    torch.int8: 3,
    torch.float8_e4m3fn: 4,
    torch.int32: 5,
    torch.uint8: 6,
    torch.int64: 7,
    torch.bool: 6,
}


def build(force: bool = False) -> str:
    """Compile ct_oracle.c -> libct_oracle.so with gcc (OpenMP if available)."""
    if (
        not force
        and os.path.exists(_SO)
        and os.path.getmtime(_SO) >= max(os.path.getmtime(p) for p in _PARTS)
    ):
        return _SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-fno-fast-math", "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for name in (
            "orc_pack_int32 orc_unpack_int32 orc_quantize orc_dequantize orc_fake_quantize "
            "orc_pack_bitmasks orc_unpack_bitmasks orc_sparse24_compress orc_sparse24_decompress "
            "orc_bitmask_compress orc_bitmask_decompress orc_quantize_pack orc_unpack_dequantize "
            "orc_semi_structured_from_dense orc_semi_structured_to_dense "
            "orc_quantize_gs orc_dequantize_gs orc_fake_quantize_gs "
            "orc_cast_to_fp4 orc_pack_fp4 orc_unpack_fp4 orc_mx_scale_compress orc_mx_scale_decompress "
            "orc_awq_repack orc_awq_repack_zeros "
            "orc_num_threads"
        ).split():
            getattr(_lib, name).restype = ctypes.c_int
        _lib.orc_set_num_threads.restype = None
        _lib.orc_set_num_threads.argtypes = [ctypes.c_int]
    return _lib


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> int:
    """use `n` OpenMP threads from now on (bench.py's CPU legs: torchrun exports OMP_NUM_THREADS=1 to its workers); returns the new maximum"""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.device.type == "cpu" and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _i64(v) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


# --------------------------------------------------------------------------- #
# pack / unpack
# --------------------------------------------------------------------------- #
def pack_to_int32(value: torch.Tensor, num_bits: int, packed_dim: int = 1) -> torch.Tensor:
    """helpers.py:20-101.  For packed_dim=0 returns the same transposed *view*
    shape as the reference (non-contiguous), built from a contiguous buffer."""
    if value.dtype is not torch.int8:
        raise ValueError("Tensor must be quantized to torch.int8 before packing")
    if not 1 <= num_bits <= 8:
        raise ValueError(f"Packing is only supported for num_bits in [1, 8], got {num_bits}")
    if value.ndim > 2:
        return torch.stack([pack_to_int32(v, num_bits, packed_dim) for v in value])
    value = value.contiguous()
    rows, cols = value.shape
    if packed_dim == 1:
        out = torch.empty(rows, math.ceil(cols * num_bits / 32), dtype=torch.int32)
    else:
        out = torch.empty(math.ceil(rows * num_bits / 32), cols, dtype=torch.int32)
    _check(lib().orc_pack_int32(_p(value), _p(out), _i64(rows), _i64(cols), num_bits, packed_dim), "pack")
    return out


def unpack_from_int32(value: torch.Tensor, num_bits: int, shape: Sequence[int], packed_dim: int = 1) -> torch.Tensor:
    """helpers.py:104-180"""
    if value.dtype is not torch.int32:
        raise ValueError(f"Expected {torch.int32} but got {value.dtype}, Aborting unpack.")
    if not 1 <= num_bits <= 8:
        raise ValueError(f"Unpacking is only supported for num_bits in [1, 8], got {num_bits}")
    if value.ndim > 2:
        return torch.stack([unpack_from_int32(v, num_bits, tuple(shape)[1:], packed_dim) for v in value])
    value = value.contiguous()
    rows, cols = int(shape[0]), int(shape[1])
    out = torch.empty(rows, cols, dtype=torch.int8)
    _check(lib().orc_unpack_int32(_p(value), _p(out), _i64(rows), _i64(cols), num_bits, packed_dim), "unpack")
    return out


# --------------------------------------------------------------------------- #
# quantize / dequantize / fake_quantize
# --------------------------------------------------------------------------- #
def _addressing(x2d_shape, scale: torch.Tensor, strategy: str, group_size, block_structure):
    """-> (rdiv, cdiv, s_row_stride, n_scale_expected)"""
    rows, cols = x2d_shape
    if strategy in ("tensor",):
        return INT64_MAX, INT64_MAX, 0
    if strategy in ("channel", "token", "attn_head"):
        if scale.numel() == 1:
            return INT64_MAX, INT64_MAX, 0
        return 1, INT64_MAX, 1
    if strategy in ("group", "tensor_group"):
        g = int(group_size)
        if cols >= g and cols % g != 0:
            raise ValueError(
                "tensor column shape must be divisble " f"by the given group_size {g} but got {cols}"
            )
        ngroups = scale.shape[-1] if scale.ndim >= 1 else 1
        srows = scale.numel() // max(ngroups, 1)
        return 1, g, (ngroups if srows > 1 else 0)
    if strategy == "block":
        bh, bw = block_structure
        nrb, ncb = math.ceil(rows / bh), math.ceil(cols / bw)
        if scale.numel() != nrb * ncb:      # the reference fails on the broadcast (forward_helpers.py:62-115); never index past the scale
            raise ValueError(f"block scale has {scale.numel()} elements, expected {nrb}x{ncb}")
        return bh, bw, ncb
    raise ValueError(strategy)


def _prep(x, scale, zero_point, g_idx, strategy):
    x2 = x.reshape(-1, x.shape[-1]).contiguous() if x.ndim != 2 else x.contiguous()
    s, z = scale, zero_point
    if strategy in ("channel", "token", "attn_head") and scale.numel() not in (1, x2.shape[0]) and scale.ndim >= 1 and scale.shape[-1] == 1:
        # plain broadcasting against the leading dims of x (forward.py:229-241), e.g. attn_head: scale [H, 1, 1] vs x [B, H, S, D]
        lead = tuple(x.shape[:-1]) + (1,)
        s = scale.expand(lead)
        z = zero_point.expand(lead) if zero_point is not None else None
    s = s.contiguous()
    z = z.contiguous() if z is not None else None
    gi = None
    if g_idx is not None and strategy in ("group", "tensor_group") and g_idx.device.type != "meta" and not bool((g_idx == -1).any()):
        gi = g_idx.to(torch.int32).contiguous()
    return x2, s, z, gi


def _qtype(qtype: str, num_bits: int) -> int:
    """0 int, 1 fp8 e4m3, 2 fp4 e2m1 (quant_args.py:482-487: FLOAT with num_bits 8 / 4)"""
    if qtype == "int":
        return 0
    if int(num_bits) == 8:
        return 1
    if int(num_bits) == 4:
        return 2
    raise NotImplementedError("Only num_bits in (4, 8) are supported")


def _global_scale(scale, global_scale):
    """`scale = scale / global_scale` (forward_helpers.py:535-536) is evaluated inside the C code in
    se = result_type(scale, global_scale); supported: one float32 value held in a 1-D (or higher) tensor"""
    if global_scale is None:
        return None, scale.dtype
    if global_scale.dtype != torch.float32 or global_scale.numel() != 1 or global_scale.ndim == 0:
        raise NotImplementedError("oracle: global_scale must be a float32 tensor of shape [1]")
    return global_scale.contiguous(), torch.result_type(scale, global_scale)


def quantize(x, scale, zero_point=None, *, strategy="tensor", group_size=None, block_structure=None,
             num_bits=8, qtype="int", dtype=None, g_idx=None, global_scale=None) -> torch.Tensor:
    """forward.py:36-73 -> _process_quantization(do_quantize=True, do_dequantize=False)."""
    gs, se = _global_scale(scale, global_scale)
    cd = torch.result_type(x, torch.empty_like(scale, dtype=se))
    x2, s, z, gi = _prep(x, scale, zero_point, g_idx, strategy)
    rdiv, cdiv, srs = _addressing(x2.shape, s, strategy, group_size, block_structure)
    if strategy in ("group", "tensor_group"):
        # _process_group: flatten().to(output_dtype); output_dtype = dtype or x.dtype (forward_helpers.py:134,171)
        out_dtype = dtype if dtype is not None else x.dtype
    else:
        out_dtype = dtype if dtype is not None else cd
    # when the group path casts compute-dtype values to x.dtype no extra rounding occurs
    # for integers / fp8 grid values, so storing straight to out_dtype is exact.
    out = torch.empty(x2.shape, dtype=out_dtype)
    rc = lib().orc_quantize_gs(_p(x2), DT[x2.dtype], _p(s), DT[s.dtype], _p(z), DT[z.dtype] if z is not None else -1,
                               _p(gi), _p(out), DT[out_dtype], _i64(x2.shape[0]), _i64(x2.shape[1]),
                               _i64(rdiv), _i64(cdiv), _i64(srs), DT[cd], _qtype(qtype, num_bits), int(num_bits), _p(gs), DT[se])
    _check(rc, "quantize")
    return out.reshape(x.shape)


def dequantize(x_q, scale, zero_point=None, *, strategy=None, group_size=None, block_structure=None,
               dtype=None, g_idx=None, global_scale=None) -> torch.Tensor:
    """forward.py:76-145 (strategy inferred from the scale shape when not given)."""
    if strategy is None:
        if scale.ndim in (0, 1):
            strategy = "tensor"
        elif scale.ndim == 2:
            if scale.shape[1] == 1:
                strategy = "channel"
            elif scale.shape[0] == 1 or scale.shape[0] == x_q.shape[0]:
                strategy = "group"
                group_size = int(x_q.shape[1] / scale.shape[1])
            else:
                strategy = "block"
                block_structure = [x_q.shape[-2] // scale.shape[0], x_q.shape[-1] // scale.shape[1]]
        else:
            raise ValueError(
                f"Could not infer a quantization strategy from scale with {scale.ndim} "
                "dimmensions. Expected 0 or 2 dimmensions."
            )
    if dtype is None:
        dtype = scale.dtype
    gs, se = _global_scale(scale, global_scale)
    x2, s, z, gi = _prep(x_q, scale, zero_point, g_idx, strategy)
    rdiv, cdiv, srs = _addressing(x2.shape, s, strategy, group_size, block_structure)
    # dtype= is honoured only on the group path
    out_dtype = dtype if strategy in ("group", "tensor_group") else se
    out = torch.empty(x2.shape, dtype=out_dtype)
    rc = lib().orc_dequantize_gs(_p(x2), DT[x2.dtype], _p(s), DT[s.dtype], _p(z), DT[z.dtype] if z is not None else -1,
                                 _p(gi), _p(out), DT[out_dtype], _i64(x2.shape[0]), _i64(x2.shape[1]),
                                 _i64(rdiv), _i64(cdiv), _i64(srs), _p(gs), DT[se])
    _check(rc, "dequantize")
    return out.reshape(x_q.shape)


def fake_quantize(x, scale, zero_point=None, *, strategy="tensor", group_size=None, block_structure=None,
                  num_bits=8, qtype="int", g_idx=None, global_scale=None) -> torch.Tensor:
    """forward.py:148-181 -> _quantize_dequantize (forward_helpers.py:180-215)."""
    gs, se = _global_scale(scale, global_scale)
    cd = torch.result_type(x, torch.empty_like(scale, dtype=se))
    x2, s, z, gi = _prep(x, scale, zero_point, g_idx, strategy)
    rdiv, cdiv, srs = _addressing(x2.shape, s, strategy, group_size, block_structure)
    out_dtype = x.dtype if strategy in ("group", "tensor_group") else se
    out = torch.empty(x2.shape, dtype=out_dtype)
    rc = lib().orc_fake_quantize_gs(_p(x2), DT[x2.dtype], _p(s), DT[s.dtype], _p(z), DT[z.dtype] if z is not None else -1,
                                    _p(gi), _p(out), DT[out_dtype], _i64(x2.shape[0]), _i64(x2.shape[1]),
                                    _i64(rdiv), _i64(cdiv), _i64(srs), DT[cd], _qtype(qtype, num_bits), int(num_bits), _p(gs), DT[se])
    _check(rc, "fake_quantize")
    return out.reshape(x.shape)


# --------------------------------------------------------------------------- #
# bitmasks and sparse formats
# --------------------------------------------------------------------------- #
def pack_bitmasks(bytemasks: torch.Tensor) -> torch.Tensor:
    """utils/helpers.py:306-317"""
    bm = bytemasks.to(torch.uint8).reshape(-1, bytemasks.shape[-1]).contiguous()
    rows, cols = bm.shape
    out = torch.empty(rows, (cols + 7) // 8, dtype=torch.uint8)
    _check(lib().orc_pack_bitmasks(_p(bm), _p(out), _i64(rows), _i64(cols)), "pack_bitmasks")
    return out.reshape(*bytemasks.shape[:-1], (cols + 7) // 8)


def unpack_bitmasks(packed: torch.Tensor, original_shape: Sequence[int]) -> torch.Tensor:
    """utils/helpers.py:320-343"""
    cols = int(original_shape[-1])
    pk = packed.reshape(-1, packed.shape[-1]).contiguous()
    rows = pk.shape[0]
    out = torch.empty(rows, cols, dtype=torch.uint8)
    _check(lib().orc_unpack_bitmasks(_p(pk), _p(out), _i64(rows), _i64(cols)), "unpack_bitmasks")
    return out.reshape(tuple(original_shape)).to(torch.bool)


def sparse24_compress(x: torch.Tensor):
    """restated Sparse24BitMaskCompressor.compress (parity unpinned) -> (values [R,C/2], bitmask u8 [R,C/8])"""
    x = x.contiguous()
    rows, cols = x.shape
    values = torch.empty(rows, cols // 2, dtype=x.dtype)
    bitmask = torch.empty(rows, (cols + 7) // 8, dtype=torch.uint8)
    _check(lib().orc_sparse24_compress(_p(x), DT[x.dtype], _p(values), _p(bitmask), _i64(rows), _i64(cols)), "s24c")
    return values, bitmask


def sparse24_decompress(values: torch.Tensor, bitmask: torch.Tensor, shape) -> torch.Tensor:
    rows, cols = int(shape[0]), int(shape[1])
    out = torch.empty(rows, cols, dtype=values.dtype)
    _check(lib().orc_sparse24_decompress(_p(values.contiguous()), DT[values.dtype], _p(bitmask.contiguous()), _p(out), _i64(rows), _i64(cols)), "s24d")
    return out


def bitmask_compress(x: torch.Tensor):
    """restated BitmaskCompressor.compress (parity unpinned) -> (values [nnz], bitmask, row_offsets i64 [R])"""
    x = x.contiguous()
    rows, cols = x.shape
    values = torch.empty(rows * cols, dtype=x.dtype)
    bitmask = torch.empty(rows, (cols + 7) // 8, dtype=torch.uint8)
    row_offsets = torch.empty(rows, dtype=torch.int64)
    nnz = ctypes.c_int64(0)
    _check(lib().orc_bitmask_compress(_p(x), DT[x.dtype], _p(values), _p(bitmask), _p(row_offsets), ctypes.byref(nnz), _i64(rows), _i64(cols)), "bmc")
    return values[: nnz.value].clone(), bitmask, row_offsets


def bitmask_decompress(values: torch.Tensor, bitmask: torch.Tensor, shape) -> torch.Tensor:
    rows, cols = int(shape[0]), int(shape[1])
    out = torch.empty(rows, cols, dtype=values.dtype)
    _check(lib().orc_bitmask_decompress(_p(values.contiguous()), DT[values.dtype], _p(bitmask.contiguous()), _p(out), _i64(rows), _i64(cols)), "bmd")
    return out


# --------------------------------------------------------------------------- #
# 2:4 semi-structured (CUTLASS / marlin-24 metadata) -- utils/semi_structured_conversions.py:66-298
# --------------------------------------------------------------------------- #
def semi_structured_from_dense(dense: torch.Tensor):
    dense = dense.contiguous()
    m, k = dense.shape
    meta_dtype = torch.int32 if dense.dtype == torch.int8 else torch.int16
    qpe = 8 if meta_dtype == torch.int32 else 4
    ks = 2 if dense.dtype == torch.float32 else 4
    sparse = torch.empty(m, k // 2, dtype=dense.dtype)
    meta = torch.empty(m, k // (ks * qpe), dtype=meta_dtype)
    _check(lib().orc_semi_structured_from_dense(_p(dense), DT[dense.dtype], _p(sparse), _p(meta), _i64(m), _i64(k)), "semi_from_dense")
    return sparse, meta


def semi_structured_to_dense(sparse: torch.Tensor, meta: torch.Tensor) -> torch.Tensor:
    sparse, meta = sparse.contiguous(), meta.contiguous()
    m, k = sparse.shape
    dense = torch.empty(m, 2 * k, dtype=sparse.dtype)
    _check(lib().orc_semi_structured_to_dense(_p(sparse), DT[sparse.dtype], _p(meta), _p(dense), _i64(m), _i64(k)), "semi_to_dense")
    return dense


# --------------------------------------------------------------------------- #
# FP4 (E2M1) and MX (E8M0 scale) pieces -- ct_oracle_fp4.c
# --------------------------------------------------------------------------- #
def cast_to_fp4(x: torch.Tensor) -> torch.Tensor:
    """quantization/utils/fp4_utils.py:77-98"""
    x = x.contiguous()
    out = torch.empty_like(x)
    _check(lib().orc_cast_to_fp4(_p(x), DT[x.dtype], _p(out), _i64(x.numel())), "cast_to_fp4")
    return out


def pack_fp4_to_uint8(x: torch.Tensor) -> torch.Tensor:
    """compressors/nvfp4/helpers.py:108-158"""
    m, n = x.shape
    if n % 2 != 0:
        raise ValueError("tensor must have an even number of columns for nvfp4 compression")
    x = x.contiguous()
    out = torch.empty((m, n // 2), dtype=torch.uint8)
    _check(lib().orc_pack_fp4(_p(x), DT[x.dtype], _p(out), _i64(m), _i64(n)), "pack_fp4_to_uint8")
    return out


def unpack_fp4_from_uint8(a: torch.Tensor, m: int, n: int, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """compressors/nvfp4/helpers.py:162-193"""
    assert a.dtype == torch.uint8
    a = a.contiguous()
    out = torch.empty((m, n), dtype=dtype)
    _check(lib().orc_unpack_fp4(_p(a), _p(out), DT[dtype], _i64(m), _i64(n)), "unpack_fp4_from_uint8")
    return out


def compress_mx_scale(scale: torch.Tensor, scale_dtype: torch.dtype = torch.uint8) -> torch.Tensor:
    """compressors/mx_utils.py:18-31"""
    scale = scale.contiguous()
    out = torch.empty(scale.shape, dtype=torch.uint8)
    _check(lib().orc_mx_scale_compress(_p(scale), DT[scale.dtype], _p(out), _i64(scale.numel())), "compress_mx_scale")
    return out.to(scale_dtype)


def decompress_mx_scale(scale: torch.Tensor) -> torch.Tensor:
    """compressors/mx_utils.py:34-44"""
    scale = scale.contiguous()
    out = torch.empty(scale.shape, dtype=torch.bfloat16)
    _check(lib().orc_mx_scale_decompress(_p(scale), _p(out), _i64(scale.numel())), "decompress_mx_scale")
    return out


# --------------------------------------------------------------------------- #
# checkpoint converters -- ct_oracle_convert.c
# --------------------------------------------------------------------------- #
def awq_repack(qweight: torch.Tensor) -> torch.Tensor:
    """entrypoints/convert/converters/autoawq.py: qweight int32 [K, N/8] -> weight_packed int32 [N, ceil(K/8)]"""
    k, nw = qweight.shape
    qweight = qweight.contiguous()
    out = torch.empty((nw * 8, (k + 7) // 8), dtype=torch.int32)
    _check(lib().orc_awq_repack(_p(qweight), _p(out), _i64(k), _i64(nw * 8)), "awq_repack")
    return out


def awq_repack_zeros(qzeros: torch.Tensor) -> torch.Tensor:
    """qzeros int32 [G, N/8] -> weight_zero_point int32 [N/8, G]"""
    g, nw = qzeros.shape
    qzeros = qzeros.contiguous()
    out = torch.empty((nw, g), dtype=torch.int32)
    _check(lib().orc_awq_repack_zeros(_p(qzeros), _p(out), _i64(g), _i64(nw * 8)), "awq_repack_zeros")
    return out


def dequantize_block_fp8(weight: torch.Tensor, scale_inv: torch.Tensor, block, dtype=torch.bfloat16) -> torch.Tensor:
    """entrypoints/convert/converters/fp8block_dequantizer.py:111-158: (w.to(f32) * s.to(f32) per block).to(dtype)"""
    return dequantize(weight, scale_inv.to(torch.float32), None, strategy="block", block_structure=list(block)).to(dtype)


# --------------------------------------------------------------------------- #
# BASELINE config 4: Sparse24BitMask + int4, restated as the composition of the restated / pinned pieces above
# (parity unpinned as a composite: the compressor pair is absent from the reference snapshot)
# --------------------------------------------------------------------------- #
def sparse24_quantize_pack(x: torch.Tensor, scale: torch.Tensor, zero_point, *, strategy="group", group_size=None, num_bits=4):
    """mask = 2:4 selection on |x| (sparse24_compress); codes = quantize(x)[mask] (forward.py:36-73); -> (pack_to_int32(codes), pack_bitmasks(mask))"""
    rows, cols = x.shape
    _, bitmask = sparse24_compress(x)
    mask = unpack_bitmasks(bitmask, (rows, cols))
    q = quantize(x, scale, zero_point, strategy=strategy, group_size=group_size, num_bits=num_bits, dtype=torch.int8)
    kept = q[mask].view(rows, cols // 2).contiguous()
    return pack_to_int32(kept, num_bits), bitmask


def sparse24_unpack_dequantize(packed: torch.Tensor, bitmask: torch.Tensor, scale: torch.Tensor, zero_point, num_bits: int, shape):
    """kept codes -> their columns -> dequantize (forward.py:76-145, strategy inferred); dropped columns are +0"""
    rows, cols = int(shape[0]), int(shape[1])
    kept = unpack_from_int32(packed, num_bits, (rows, cols // 2))
    mask = unpack_bitmasks(bitmask, (rows, cols))
    q = torch.zeros(rows, cols, dtype=torch.int8)
    q[mask] = kept.reshape(-1)
    out = dequantize(q, scale, zero_point)
    return torch.where(mask, out, torch.zeros_like(out))

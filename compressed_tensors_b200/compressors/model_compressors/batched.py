"""
Whole-model launches: every pack-quantized / naive-quantized module whose tensors are resident on
one CUDA device is compressed (or decompressed) by a single multi-tensor kernel launch per
(format, dtype, bit-width) signature.  Results are bit-identical to the per-module plugin path
(tests/test_gpu_model.py); modules that do not qualify go through that path unchanged.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Optional

import torch

from ... import _native as N
from ... import ops
from ...config import CompressionFormat
from ...quantization import QuantizationScheme, QuantizationStatus
from ...utils.module import get_direct_state_dict, replace_direct_state_dict
from ..base import BaseCompressor, _resolve_format, compress_module, decompress_module

__all__ = ["compress_modules_batched", "decompress_modules_batched"]

_BATCHABLE = (CompressionFormat.pack_quantized, CompressionFormat.naive_quantized, CompressionFormat.int_quantized,
              CompressionFormat.float_quantized)


def _eligible(module, fmt) -> bool:
    if fmt not in _BATCHABLE:
        return False
    w = getattr(module, "weight", None) if fmt is not None else None
    sd = module._parameters
    t = sd.get("weight", None) if "weight" in sd else sd.get("weight_packed", None)
    sc = sd.get("weight_scale", None)
    if t is None or sc is None or not t.is_cuda or not sc.is_cuda or t.ndim != 2:
        return False
    if sd.get("weight_g_idx", None) is not None and bool((sd["weight_g_idx"] != -1).all()):
        return False  # activation ordering: per-module path
    return True


def compress_modules_batched(modules, force_format: Optional[CompressionFormat] = None) -> None:
    groups = defaultdict(list)
    for m in modules:
        scheme = getattr(m, "quantization_scheme", None)
        if not isinstance(scheme, QuantizationScheme):
            continue
        fmt = _resolve_format(m, scheme, force_format)
        if _eligible(m, fmt) and scheme.weights is not None and "weight" in m._parameters:
            groups[(fmt, m.weight.device.index)].append(m)
        else:
            compress_module(m, force_format)

    for (fmt, dev), mods in groups.items():
        comp = BaseCompressor.get_value_from_registry(fmt.value)
        pack = fmt == CompressionFormat.pack_quantized
        probs, staged = [], []
        for m in mods:
            scheme = m.quantization_scheme
            scheme.format = fmt
            args = scheme.weights
            sd = get_direct_state_dict(m)
            w, sc = sd["weight"], sd["weight_scale"]
            zp = sd.get("weight_zero_point", None) if not args.symmetric else None
            try:
                p = ops._resolve(w, sc, zp, args, None)
                qtype, bits = ops._qparams(args)
                cd = torch.result_type(w, sc)
                if pack:
                    out = torch.empty((p.rows, -(-p.cols * bits // 32)), dtype=torch.int32, device=w.device)
                    d = ops._desc(p, w.dtype, p.scale.dtype, p.zp.dtype if p.zp is not None else None, cd, torch.int8, None, qtype, bits)
                else:
                    qd = args.pytorch_dtype()
                    out = torch.empty(w.shape, dtype=qd, device=w.device)
                    d = ops._desc(p, w.dtype, p.scale.dtype, p.zp.dtype if p.zp is not None else None, cd, qd, qd, qtype, bits)
            except (ValueError, NotImplementedError):
                compress_module(m, force_format)
                continue
            probs.append((d, w.contiguous(), p.scale.contiguous(), p.zp.contiguous() if p.zp is not None else None, out))
            staged.append((m, sd, out))
        if probs:
            ops.batched(N.OP_QUANTIZE_PACK if pack else N.OP_QUANTIZE, probs, dev)
        for m, sd, out in staged:
            scheme = m.quantization_scheme
            args = scheme.weights
            new = sd.copy()
            w = new.pop("weight")
            if pack:
                new["weight_packed"] = out
                new["weight_shape"] = torch.tensor(w.shape)
                if not args.symmetric and args.strategy in ("group", "channel"):
                    new["weight_zero_point"] = ops.pack_to_int32(sd["weight_zero_point"], args.num_bits, packed_dim=0).contiguous()
            else:
                new["weight"] = out
            new = comp._remove_symmetric_zp(new, scheme)
            replace_direct_state_dict(m, new)
            m.quantization_status = QuantizationStatus.COMPRESSED


def decompress_modules_batched(modules, force_format: Optional[CompressionFormat] = None) -> None:
    groups = defaultdict(list)
    for m in modules:
        scheme = getattr(m, "quantization_scheme", None)
        if not isinstance(scheme, QuantizationScheme):
            continue
        fmt = _resolve_format(m, scheme, force_format)
        key_t = "weight_packed" if fmt == CompressionFormat.pack_quantized else "weight"
        if _eligible(m, fmt) and key_t in m._parameters and scheme.weights is not None:
            groups[(fmt, m._parameters[key_t].device.index)].append(m)
        else:
            decompress_module(m, force_format)

    for (fmt, dev), mods in groups.items():
        pack = fmt == CompressionFormat.pack_quantized
        probs, staged = [], []
        for m in mods:
            scheme = m.quantization_scheme
            scheme.format = fmt
            args = scheme.weights
            sd = get_direct_state_dict(m)
            sc = sd["weight_scale"]
            zp = sd.get("weight_zero_point", None)
            new = sd.copy()
            try:
                if pack:
                    packed = sd["weight_packed"]
                    shape = tuple(int(v) for v in sd["weight_shape"].tolist())
                    if zp is not None and not args.symmetric and args.strategy in ("group", "channel"):
                        zp = ops.unpack_from_int32(zp, args.num_bits, (*shape[:-1], sc.shape[-1]), packed_dim=0)
                        new["weight_zero_point"] = zp
                    like = torch.empty(shape, dtype=torch.int8, device="meta")
                    iargs = ops._infer_dequant_args(like, sc)
                    p = ops._resolve(like, sc, zp, iargs, None)
                    out_dtype = sc.dtype
                    out = torch.empty(shape, dtype=out_dtype, device=packed.device)
                    d = ops._desc(p, None, p.scale.dtype, p.zp.dtype if p.zp is not None else None, None, torch.int8, out_dtype, N.Q_INT, args.num_bits)
                    src = packed.contiguous()
                    new.pop("weight_packed")
                else:
                    q = sd["weight"]
                    iargs = ops._infer_dequant_args(q, sc)
                    p = ops._resolve(q, sc, zp, iargs, None)
                    out_dtype = sc.dtype
                    out = torch.empty(q.shape, dtype=out_dtype, device=q.device)
                    d = ops._desc(p, None, p.scale.dtype, p.zp.dtype if p.zp is not None else None, None, q.dtype, out_dtype, N.Q_INT, 8)
                    src = q.contiguous()
            except (ValueError, NotImplementedError):
                decompress_module(m, force_format)
                continue
            probs.append((d, src, p.scale.contiguous(), p.zp.contiguous() if p.zp is not None else None, out))
            new["weight"] = out
            staged.append((m, new))
        if probs:
            ops.batched(N.OP_UNPACK_DEQUANTIZE if pack else N.OP_DEQUANTIZE, probs, dev)
        for m, new in staged:
            replace_direct_state_dict(m, new)
            m.quantization_status = QuantizationStatus.DECOMPRESSED

"""
Whole-model launches: every pack-quantized / naive-quantized module is compressed (or decompressed)
together with all other modules of the same (format, placement) group

  * tensors resident on one CUDA device  -> ONE multi-tensor kernel launch per kernel signature (ct_batched)
  * tensors resident in host memory      -> ONE pipelined H2D -> kernel -> D2H queue across all of them
                                            (ct_host_run_many); outputs land in host memory (pinned when
                                            the inputs are)

Results are bit-identical to the per-module plugin path (tests/test_gpu_compressors.py); modules that do
not qualify (other formats, activation ordering, N-D weights, meta tensors) take that path unchanged.
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
# fp4 nibble formats: one multi-tensor launch on CUDA placements, the cross-tensor host pipeline for host-resident modules
_FP4 = (CompressionFormat.nvfp4_pack_quantized, CompressionFormat.mxfp4_pack_quantized)


def _fp4_ok(module, fmt, place, key: str) -> bool:
    if fmt not in _FP4 or place is None:
        return False
    if fmt == CompressionFormat.nvfp4_pack_quantized:
        gs = module._parameters.get("weight_global_scale", None)
        t = module._parameters.get(key)
        return gs is not None and gs.device == t.device and gs.dtype == torch.float32 and gs.numel() == 1 and gs.ndim >= 1
    return True


def _compress_fp4_group(mods, fmt, place, force_format) -> None:
    comp = BaseCompressor.get_value_from_registry(fmt.value)
    probs, staged, keep = [], [], []
    for m in mods:
        scheme = m.quantization_scheme
        scheme.format = fmt
        args = scheme.weights
        sd = get_direct_state_dict(m)
        w, sc = sd["weight"], sd["weight_scale"]
        gs = sd.get("weight_global_scale", None) if fmt == CompressionFormat.nvfp4_pack_quantized else None
        try:
            if w.shape[1] % 2 != 0:
                raise ValueError("odd columns")
            scale, gsv, se = ops._global_scale(sc, gs)
            qtype, bits = ops._qparams(args)
            if qtype != N.Q_FP4 or (gs is not None and gsv is None):
                raise NotImplementedError
            cd = torch.result_type(w, scale if gsv is None else ops._like(scale, se))
            zp = sd.get("weight_zero_point", None)
            p = ops._resolve(w, scale, zp, args, None)
            d = ops._desc(p, w.dtype, p.scale.dtype, p.zp.dtype if p.zp is not None else None, cd, w.dtype, None, qtype, bits, se)
            if gsv is not None:
                d.global_scale = gsv.data_ptr()
                keep.append(gsv)
            out = _empty((p.rows, p.cols // 2), torch.uint8, w)
        except (ValueError, NotImplementedError):
            compress_module(m, force_format)
            continue
        probs.append((d, w.contiguous(), p.scale.contiguous(), p.zp.contiguous() if p.zp is not None else None, out))
        staged.append((m, sd, out))
    _run(N.OP_QUANTIZE_PACK_FP4, probs, place)
    for m, sd, out in staged:
        scheme = m.quantization_scheme
        new = sd.copy()
        new.pop("weight")
        new["weight_packed"] = out
        new["weight_scale"] = comp._compress_scale(sd["weight_scale"], scheme.weights)
        replace_direct_state_dict(m, comp._remove_symmetric_zp(new, scheme))
        m.quantization_status = QuantizationStatus.COMPRESSED


def _decompress_fp4_group(mods, fmt, place, force_format) -> None:
    comp = BaseCompressor.get_value_from_registry(fmt.value)
    dense = torch.bfloat16   # like the reference: unpack_fp4_from_uint8's default dtype (nvfp4/base.py:116)
    probs, staged, keep = [], [], []
    for m in mods:
        scheme = m.quantization_scheme
        scheme.format = fmt
        sd = get_direct_state_dict(m)
        packed, sc = sd["weight_packed"], sd["weight_scale"]
        gs = sd.get("weight_global_scale", None)
        try:
            if packed.dtype != torch.uint8 or packed.ndim != 2:
                raise ValueError
            stored = {torch.float8_e4m3fn: N.DT[torch.float8_e4m3fn], torch.uint8: N.DT_E8M0}.get(sc.dtype)
            if stored is None or (gs is not None and not (gs.dtype == torch.float32 and gs.numel() == 1 and gs.ndim >= 1)):
                raise NotImplementedError
            rows, cols = packed.shape[0], packed.shape[1] * 2
            like = torch.empty((rows, cols), dtype=torch.int8, device="meta")
            p = ops._resolve(like, sc, None, ops._infer_dequant_args(like, sc), None)
            se = torch.float32 if gs is not None else dense
            d = ops._desc(p, None, torch.float32, None, None, None, dense, N.Q_FP4, 4, se)
            d.scale_dtype = stored
            if gs is not None:
                gsv = gs.reshape(1).contiguous()
                d.global_scale = gsv.data_ptr()
                keep.append(gsv)
            out = _empty((rows, cols), dense, packed)
        except (ValueError, NotImplementedError):
            decompress_module(m, force_format)
            continue
        probs.append((d, packed.contiguous(), p.scale.contiguous(), None, out))
        new = sd.copy()
        new.pop("weight_packed")
        new["weight"] = out
        new["weight_scale"] = comp._decompress_scale(sc, dense)
        staged.append((m, new))
    _run(N.OP_UNPACK_DEQUANTIZE_FP4, probs, place)
    for m, new in staged:
        replace_direct_state_dict(m, new)
        m.quantization_status = QuantizationStatus.DECOMPRESSED


def _placement(module, key: str):
    """('cuda', index) / ('cpu', None) when the module's streamed tensor and qparams share a placement, else None"""
    sd = module._parameters
    t, sc = sd.get(key, None), sd.get("weight_scale", None)
    if t is None or sc is None or t.ndim != 2 or t.device != sc.device or t.device.type not in ("cuda", "cpu"):
        return None
    zp = sd.get("weight_zero_point", None)
    if zp is not None and zp.device != t.device:
        return None
    g = sd.get("weight_g_idx", None)
    if g is not None and bool((g != -1).all()):
        return None  # activation ordering: per-module path
    if t.device.type == "cpu" and not torch.cuda.is_available():
        return None  # the per-module path raises the loud "no CUDA device" error
    return (t.device.type, t.device.index)


def _run(op: int, probs, place):
    if not probs:
        return
    if place[0] == "cuda":
        ops.batched(op, probs, place[1])
    else:
        ops.host_batched(op, probs)


def _empty(shape, dtype, like: torch.Tensor):
    if like.is_cuda:
        return torch.empty(shape, dtype=dtype, device=like.device)
    return torch.empty(shape, dtype=dtype, pin_memory=like.is_pinned())


def compress_modules_batched(modules, force_format: Optional[CompressionFormat] = None) -> None:
    groups = defaultdict(list)
    for m in modules:
        scheme = getattr(m, "quantization_scheme", None)
        if not isinstance(scheme, QuantizationScheme):
            continue
        fmt = _resolve_format(m, scheme, force_format)
        place = _placement(m, "weight") if (fmt in _BATCHABLE + _FP4 and scheme.weights is not None) else None
        if place is not None and (fmt in _BATCHABLE or _fp4_ok(m, fmt, place, "weight")):
            groups[(fmt, place)].append(m)
        else:
            compress_module(m, force_format)

    for (fmt, place), mods in groups.items():
        if fmt in _FP4:
            _compress_fp4_group(mods, fmt, place, force_format)
            continue
        comp = BaseCompressor.get_value_from_registry(fmt.value)
        pack = fmt == CompressionFormat.pack_quantized
        probs, staged = [], []
        for m in mods:
            scheme = m.quantization_scheme
            scheme.format = fmt
            args = scheme.weights
            sd = get_direct_state_dict(m)
            w, sc = sd["weight"], sd["weight_scale"]
            # the zero point goes through like in the per-module plugin path (ops.quantize): for integer codes a symmetric
            # scheme's all-zero zero point cannot change a code (x/s + 0 only turns -0.0 into +0.0, which rounds to the same
            # integer), so it is skipped; for fp8 codes -0.0 and +0.0 are different bytes (0x80 / 0x00), so it is added
            zp = sd.get("weight_zero_point", None) if (not args.symmetric or args.type == "float") else None
            try:
                p = ops._resolve(w, sc, zp, args, None)
                qtype, bits = ops._qparams(args)
                cd = torch.result_type(w, sc)
                if pack:
                    out = _empty((p.rows, -(-p.cols * bits // 32)), torch.int32, w)
                    d = ops._desc(p, w.dtype, p.scale.dtype, p.zp.dtype if p.zp is not None else None, cd, torch.int8, None, qtype, bits)
                else:
                    qd = args.pytorch_dtype()
                    out = _empty(w.shape, qd, w)
                    d = ops._desc(p, w.dtype, p.scale.dtype, p.zp.dtype if p.zp is not None else None, cd, qd, qd, qtype, bits)
            except (ValueError, NotImplementedError):
                compress_module(m, force_format)
                continue
            probs.append((d, w.contiguous(), p.scale.contiguous(), p.zp.contiguous() if p.zp is not None else None, out))
            staged.append((m, sd, out))
        _run(N.OP_QUANTIZE_PACK if pack else N.OP_QUANTIZE, probs, place)
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
        key_t = "weight_packed" if fmt in (CompressionFormat.pack_quantized,) + _FP4 else "weight"
        place = _placement(m, key_t) if (fmt in _BATCHABLE + _FP4 and scheme.weights is not None) else None
        if place is not None and (fmt in _BATCHABLE or _fp4_ok(m, fmt, place, key_t)):
            groups[(fmt, place)].append(m)
        else:
            decompress_module(m, force_format)

    for (fmt, place), mods in groups.items():
        if fmt in _FP4:
            _decompress_fp4_group(mods, fmt, place, force_format)
            continue
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
                    out = _empty(shape, sc.dtype, packed)
                    d = ops._desc(p, None, p.scale.dtype, p.zp.dtype if p.zp is not None else None, None, torch.int8, sc.dtype, N.Q_INT, args.num_bits)
                    src = packed.contiguous()
                    new.pop("weight_packed")
                else:
                    q = sd["weight"]
                    iargs = ops._infer_dequant_args(q, sc)
                    p = ops._resolve(q, sc, zp, iargs, None)
                    out = _empty(q.shape, sc.dtype, q)
                    d = ops._desc(p, None, p.scale.dtype, p.zp.dtype if p.zp is not None else None, None, q.dtype, sc.dtype, N.Q_INT, 8)
                    src = q.contiguous()
            except (ValueError, NotImplementedError):
                decompress_module(m, force_format)
                continue
            probs.append((d, src, p.scale.contiguous(), p.zp.contiguous() if p.zp is not None else None, out))
            new["weight"] = out
            staged.append((m, new))
        _run(N.OP_UNPACK_DEQUANTIZE if pack else N.OP_DEQUANTIZE, probs, place)
        for m, new in staged:
            replace_direct_state_dict(m, new)
            m.quantization_status = QuantizationStatus.DECOMPRESSED

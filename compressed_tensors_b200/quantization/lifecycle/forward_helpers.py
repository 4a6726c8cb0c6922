"""
Per-op entry points of the reference's forward_helpers.py (:180-215 `_quantize_dequantize`, :523-546 `_quantize`, :549-572
`_dequantize`).  In the reference these run AFTER `_process_group` / `_process_block` reshaped x so that scale broadcasts against
it; here each is one launch of the generic CUDA kernels on exactly that broadcast layout (x [..., K], scale / zero point
broadcastable to x), so callers that reach below `quantize()` -- the reference's own tests and ImplBackend users -- get the same
bits from the B200.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from ... import ops
from ..quant_args import QuantizationArgs

__all__ = ["_quantize", "_dequantize", "_quantize_dequantize", "_is_fp8_supported"]


def _bcast_args(x: torch.Tensor, scale: torch.Tensor, args):
    """strategy that makes ops.* address `scale` the way plain broadcasting against x would"""
    bits, typ = (args.num_bits, args.type) if args is not None else (8, "int")
    if scale.numel() == 1:
        return SimpleNamespace(strategy="tensor", group_size=None, block_structure=None, num_bits=bits, type=typ)
    if scale.shape[-1] == 1:
        return SimpleNamespace(strategy="channel", group_size=None, block_structure=None, num_bits=bits, type=typ)
    raise NotImplementedError("forward_helpers entry points take per-tensor or per-row (trailing dimension 1) scales; use quantize() / dequantize()")


@torch.no_grad()
def _quantize(x, scale, zero_point, q_min, q_max, args: QuantizationArgs, dtype=None, global_scale=None):
    return ops.quantize(x, scale, zero_point, _bcast_args(x, scale, args), dtype=dtype, global_scale=global_scale)


@torch.no_grad()
def _dequantize(x_q, scale, zero_point=None, dtype=None, global_scale=None):
    out = ops.dequantize(x_q, scale, zero_point, args=_bcast_args(x_q, scale, None), global_scale=global_scale)
    return out.to(dtype) if dtype is not None else out


@torch.no_grad()
def _quantize_dequantize(x, scale, zero_point, q_min, q_max, args: QuantizationArgs, global_scale=None):
    return ops.fake_quantize(x, scale, zero_point, _bcast_args(x, scale, args), global_scale=global_scale)


def _is_fp8_supported(device: torch.device) -> bool:
    """native fp8 conversions: every device this engine runs on (sm_100a) has them (forward_helpers.py:346-354)"""
    device = torch.device(device)
    if device.type == "cuda":
        major, _ = torch.cuda.get_device_capability(device)
        return major >= 9
    return False


def adapt_scale_and_zp_for_triton(scale: torch.Tensor, zero_point: torch.Tensor | None, num_rows: int):
    """qparams as one contiguous row per weight row (forward_helpers.py:357-384): 0-d / 1-d / one-row tensors are broadcast over
    `num_rows`.  The reference's Triton kernels need this layout; the CUDA kernels here address the original qparams directly
    (DESIGN.md section 4), so only the reference's benchmark helpers call it."""
    def rows(t):
        if t is None:
            return None
        if t.ndim == 0:
            t = t.expand(num_rows, 1)
        elif t.ndim == 1:
            t = t.unsqueeze(1).expand(num_rows, 1)
        elif t.shape[0] == 1:
            t = t.expand(num_rows, -1)
        return t.contiguous()

    return rows(scale), rows(zero_point)


__all__.append("adapt_scale_and_zp_for_triton")

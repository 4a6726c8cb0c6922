"""
2:4 semi-structured (CUTLASS / marlin-24) conversions -- same functions and error behaviour as the
reference's utils/semi_structured_conversions.py (:66-197, :204-298, :301-330); the work is done by
ct_semi_structured_from_dense / ct_semi_structured_to_dense / the 2:4 select kernel.

The reference's fp32 to-dense path scatters through a float16 view on the CPU and thereby sets the
quiet bit of 16-bit halves that look like fp16 signalling NaNs; this implementation moves bits
unchanged (see tests/test_oracle_golden.py::test_semi_structured_golden).
"""
from __future__ import annotations

import torch

from .. import _native as N
from .. import ops

__all__ = ["sparse_semi_structured_from_dense_cutlass", "sparse_semi_structured_to_dense_cutlass", "mask_creator"]


def _meta_dtype(dtype: torch.dtype) -> torch.dtype:
    if dtype == torch.int8:
        return torch.int32
    if dtype in (torch.half, torch.bfloat16, torch.float):
        return torch.int16
    raise RuntimeError(f"Invalid datatype {dtype} of dense matrix")


def sparse_semi_structured_from_dense_cutlass(dense: torch.Tensor):
    """dense [m, k] (2:4 sparse; 1:2 for fp32) -> (sparse [m, k/2], reordered metadata)"""
    if dense.dim() != 2:
        raise RuntimeError(f"Expected 2-dimensional dense tensor, got {dense.dim()}-dimensional tensor")
    m, k = dense.shape
    meta_dtype = _meta_dtype(dense.dtype)
    qpe = meta_dtype.itemsize * 8 // 4
    if meta_dtype == torch.int32:
        if m % 16 != 0:
            raise RuntimeError(f"Number of rows of dense matrix {m} must be divisible by 16")
    elif m % 32 != 0:
        raise RuntimeError(f"Number of rows of dense matrix {m} must be divisible by 32")
    ks = 2 if dense.dtype == torch.float else 4
    if k % (4 * qpe) != 0:
        raise RuntimeError(f"Number of columns of dense matrix {k} must be divisible by {4 * qpe}")
    idx = ops._dev_index(dense)
    d = ops._to_dev(dense, idx)
    sparse = torch.empty((m, k // 2), dtype=dense.dtype, device=d.device)
    meta = torch.empty((m, k // (ks * qpe)), dtype=meta_dtype, device=d.device)
    N.check(N.lib().ct_semi_structured_from_dense(N.ptr(d), N.DT[dense.dtype], N.ptr(sparse), N.ptr(meta), m, k, idx, N.stream_ptr(idx)),
            "sparse_semi_structured_from_dense_cutlass")
    return ops._back(sparse, dense), ops._back(meta, dense)


def sparse_semi_structured_to_dense_cutlass(sparse: torch.Tensor, meta_reordered: torch.Tensor) -> torch.Tensor:
    """(sparse [m, k], metadata) -> dense [m, 2k]"""
    if sparse.dim() != 2:
        raise RuntimeError(f"Expected 2-dimensional sparse tensor, got {sparse.dim()}-dimensional tensor")
    m, k = sparse.shape
    if meta_reordered.dim() != 2:
        raise RuntimeError(f"Expected 2-dimensional meta tensor, got {meta_reordered.dim()}-dimensional tensor")
    if meta_reordered.device != sparse.device:
        raise RuntimeError(f"Expected meta matrix to be on {sparse.device} device, got matrix on {meta_reordered.device} device")
    if meta_reordered.dtype not in (torch.int16, torch.int32):
        raise RuntimeError(f"Invalid datatype {meta_reordered.dtype} of meta matrix")
    qpe = meta_reordered.dtype.itemsize * 8 // 4
    ks = 4 if sparse.dtype != torch.float else 2
    rows, cols = meta_reordered.shape
    if rows != m:
        raise RuntimeError(f"Number of rows of meta matrix {rows} must be equal to number of columns of spase matrix {m}")
    if cols * ks * qpe != 2 * k:
        raise RuntimeError(
            f"Number of columns of sparse matrix {k} different from the {cols * ks * qpe // 2}, "
            "expected according to the number of columns of meta matrix"
        )
    if _meta_dtype(sparse.dtype) != meta_reordered.dtype:
        raise RuntimeError(f"Invalid datatype {meta_reordered.dtype} of meta matrix")
    idx = ops._dev_index(sparse)
    s, mt = ops._to_dev(sparse, idx), ops._to_dev(meta_reordered, idx)
    dense = torch.empty((m, 2 * k), dtype=sparse.dtype, device=s.device)
    N.check(N.lib().ct_semi_structured_to_dense(N.ptr(s), N.DT[sparse.dtype], N.ptr(mt), N.ptr(dense), m, k, idx, N.stream_ptr(idx)),
            "sparse_semi_structured_to_dense_cutlass")
    return ops._back(dense, sparse)


def mask_creator(tensor: torch.Tensor) -> torch.Tensor:
    """2:4 mask (float32 ones / zeros, tensor's shape): the 2 largest |x| of every 4 consecutive
    elements are kept (semi_structured_conversions.py:301-330).  The reference drops the first two
    entries of an ascending argsort, which on ties drops the LOWER columns; the select kernel keeps
    the lower column on ties, so it is run on the column-reversed quads."""
    if tensor.numel() % 4 != 0:
        raise ValueError(f"Tensor of size {tensor.shape} can't be evenly divided into 4 groups")
    flat = tensor.detach().reshape(-1, 4).flip(-1)
    _, bitmask = ops.sparse24_compress(flat.contiguous())
    mask = ops.unpack_bitmasks(bitmask, flat.shape).flip(-1)
    return mask.to(torch.float32).reshape(tensor.shape)

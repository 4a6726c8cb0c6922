"""oracle front end for calculate_qparams (TEST INFRASTRUCTURE ONLY, see ct_oracle_qparams.c)"""
import ctypes

import torch

from . import DT, _check, _i64, _p, lib


def calculate_qparams(min_vals: torch.Tensor, max_vals: torch.Tensor, *, num_bits: int, qtype: str = "int", symmetric: bool = True):
    """reference quantization/utils/helpers.py:50-137; returns (scale in the min/max dtype, zero point
    in the default zp dtype: int8 for integer quantization, float8_e4m3fn for fp8)"""
    mn, mx = min_vals.contiguous(), max_vals.contiguous()
    zp_dtype = torch.int8 if qtype == "int" else torch.float8_e4m3fn
    scale = torch.empty(mn.shape, dtype=mn.dtype)
    zp = torch.empty(mn.shape, dtype=zp_dtype)
    L = lib()
    L.orc_calculate_qparams.restype = ctypes.c_int
    _check(L.orc_calculate_qparams(_p(mn), _p(mx), DT[mn.dtype], _p(scale), _p(zp), DT[zp_dtype], _i64(mn.numel()),
                                   0 if qtype == "int" else 1, int(num_bits), 1 if symmetric else 0), "calculate_qparams")
    if scale.ndim == 0:
        scale, zp = scale.reshape(1), zp.reshape(1)
    return scale, zp

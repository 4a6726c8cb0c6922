#!/usr/bin/env python
"""
Golden vectors for the FP4 / MX formats, produced by importing the reference
exactly like make_golden.py does:

  cast_to_fp4                     quantization/utils/fp4_utils.py:77-98
  pack_fp4_to_uint8 / unpack      compressors/nvfp4/helpers.py:108-193
  quantize / dequantize / fake_quantize with FP4 args and global_scale (tensor_group, group 16)
                                  quantization/lifecycle/forward_helpers.py:118-215, 523-572
  MXFP4 / MXFP8 group-32 power-of-two scales, compress_mx_scale / decompress_mx_scale
                                  compressors/mx_utils.py:18-44, quantization/utils/mxfp_utils.py
  NVFP4PackedCompressor, MXFP4PackedCompressor, MXFP8QuantizationCompressor  (state dict in -> out)
  calculate_qparams / generate_gparam for those schemes   quantization/utils/helpers.py:50-137, 308-337

    python tests/golden/make_golden_fp4.py   ->  tests/golden/fp4.pt.gz
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference from a temp copy)

from compressed_tensors.compressors.mx_utils import compress_mx_scale, decompress_mx_scale  # noqa: E402
from compressed_tensors.compressors.mxfp4.base import MXFP4PackedCompressor  # noqa: E402
from compressed_tensors.compressors.mxfp8.base import MXFP8QuantizationCompressor  # noqa: E402
from compressed_tensors.compressors.nvfp4.base import NVFP4PackedCompressor  # noqa: E402
from compressed_tensors.compressors.nvfp4.helpers import pack_fp4_to_uint8, unpack_fp4_from_uint8  # noqa: E402
from compressed_tensors.quantization import QuantizationArgs, QuantizationScheme  # noqa: E402
from compressed_tensors.quantization.lifecycle.forward import dequantize, fake_quantize, quantize  # noqa: E402
from compressed_tensors.quantization.quant_args import FP4_E2M1_DATA  # noqa: E402
from compressed_tensors.quantization.utils import calculate_qparams, generate_gparam  # noqa: E402

DTS = (torch.bfloat16, torch.float16, torch.float32)
FP4 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
NV = dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16, scale_dtype=torch.float8_e4m3fn, zp_dtype=torch.float8_e4m3fn)
MX4 = dict(num_bits=4, type="float", symmetric=True, strategy="group", group_size=32, scale_dtype=torch.uint8, zp_dtype=torch.uint8)
MX8 = dict(num_bits=8, type="float", symmetric=True, strategy="group", group_size=32, scale_dtype=torch.uint8, zp_dtype=torch.uint8)


def dump(args):
    return args.model_dump(mode="json")


def weights(g, rows, cols, dt, spread=True):
    w = torch.randn(rows, cols, generator=g) * 0.05
    if spread:  # per-group magnitudes over a few octaves, a dead group, an outlier
        w = w * torch.exp2(torch.randint(-3, 4, (rows, cols // 16, 1), generator=g).float()).expand(-1, -1, 16).reshape(rows, cols)
        w[0, :16] = 0
        w[-1, -1] = 3.0
    return w.to(dt)


def group_minmax(w, gsize):
    g = w.unflatten(-1, (-1, gsize))
    return g.amin(-1), g.amax(-1)


def main():
    g = torch.Generator().manual_seed(4242)
    out = {}

    # ---- cast_to_fp4 -------------------------------------------------------------------------
    edges = torch.tensor([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0])
    cast = []
    for dt in DTS:
        eps = torch.finfo(dt).eps
        sp = torch.cat([edges, edges * (1 + 2 * eps), edges * (1 - 2 * eps), FP4, torch.tensor([0.0, 1e-8, 6.5, 100.0, float("inf"), float("nan")])])
        x = torch.cat([sp, -sp, torch.randn(4096, generator=g) * 2.5]).to(dt)
        x[-1] = -0.0
        cast.append(dict(x=x, y=FP4_E2M1_DATA.cast_to_fp4(x.clone())))
    out["cast"] = cast

    # ---- pack / unpack -----------------------------------------------------------------------
    packs = []
    for shape in ((4, 8), (16, 64), (3, 10), (1, 2)):
        idx = torch.randint(0, 8, shape, generator=g)
        sgn = torch.where(torch.rand(shape, generator=g) < 0.5, -1.0, 1.0)
        for dt in DTS:
            x = (FP4[idx] * sgn).to(dt)  # includes -0.0
            p = pack_fp4_to_uint8(x)
            packs.append(dict(x=x, packed=p, unpacked={str(d): unpack_fp4_from_uint8(p, shape[0], shape[1], dtype=d) for d in DTS}))
    out["pack"] = packs

    # ---- NVFP4: quantize / dequantize / fake_quantize with a global scale -----------------------
    nv = []
    args = QuantizationArgs(**NV)
    for dt in DTS:
        for shape in ((8, 64), (32, 256)):
            w = weights(g, *shape, dt)
            gs = generate_gparam(w.min(), w.max())
            mn, mx = group_minmax(w, 16)
            scale, zp = calculate_qparams(mn, mx, args, global_scale=gs)
            for sdt in (None, torch.bfloat16, torch.float16):   # fp8 scales cannot be divided by the global scale in the reference (no fp8 promotion)
                s = scale if sdt is None else scale.to(sdt)
                q = quantize(w, s, zp, args, global_scale=gs)
                fq = fake_quantize(w, s, zp, args, global_scale=gs)
                dq = dequantize(q, s, global_scale=gs, dtype=dt)
                nv.append(dict(args=dump(args), x=w, scale=s, global_scale=gs, q=q, fq=fq, dq=dq, qparams_min=mn, qparams_max=mx, qparams_scale=scale, qparams_zp=zp))
    out["nvfp4"] = nv

    # ---- MXFP4 / MXFP8: power-of-two group-32 scales ---------------------------------------------
    mx_cases = []
    for kw in (MX4, MX8):
        args = QuantizationArgs(**kw)
        for dt in DTS:
            w = weights(g, 16, 128, dt)
            mn, mx = group_minmax(w, 32)
            scale, zp = calculate_qparams(mn, mx, args)
            q = quantize(w, scale, zp, args, dtype=(torch.float8_e4m3fn if kw["num_bits"] == 8 else None))
            fq = fake_quantize(w, scale, zp, args)
            dq = dequantize(q, scale, dtype=dt)
            mx_cases.append(dict(args=dump(args), x=w, scale=scale, q=q, fq=fq, dq=dq, qparams_min=mn, qparams_max=mx, qparams_zp=zp))
    out["mx"] = mx_cases

    # ---- E8M0 scale encode / decode -------------------------------------------------------------
    e8 = []
    for dt in DTS:
        s = torch.cat([torch.exp2(torch.randint(-20, 20, (256,), generator=g).float()), torch.rand(1024, generator=g) * 4 + 1e-4]).to(dt)
        enc = compress_mx_scale(s, torch.uint8)
        e8.append(dict(scale=s, enc=enc, dec=decompress_mx_scale(enc)))
    out["e8m0"] = e8

    # ---- compressors (state dict in -> state dict out) -------------------------------------------
    comp = []
    for name, cls, kw, gsz in (("nvfp4", NVFP4PackedCompressor, NV, 16), ("mxfp4", MXFP4PackedCompressor, MX4, 32), ("mxfp8", MXFP8QuantizationCompressor, MX8, 32)):
        args = QuantizationArgs(**kw)
        scheme = QuantizationScheme(targets=["Linear"], weights=args)
        for dt in (torch.bfloat16, torch.float16):
            w = weights(g, 32, 128, dt)
            mn, mx = group_minmax(w, gsz)
            sd = {"weight": w}
            if name == "nvfp4":
                gs = generate_gparam(w.min(), w.max())
                sd["weight_global_scale"] = gs
                scale, zp = calculate_qparams(mn, mx, args, global_scale=gs)
            else:
                scale, zp = calculate_qparams(mn, mx, args)
            sd["weight_scale"] = scale
            sd["weight_zero_point"] = zp
            c = cls.compress(sd, scheme)
            d = cls.decompress(c, scheme)
            comp.append(dict(format=name, args=dump(args), state=sd, compressed={k: v for k, v in c.items() if v is not None},
                             decompressed={k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in d.items() if v is not None}))
    out["compressors"] = comp
    mg.save("fp4.pt", out)


if __name__ == "__main__":
    main()

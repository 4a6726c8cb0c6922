"""
Differential fuzz of `ModelCompressor` against the reference's (TEST INFRASTRUCTURE; build container only): the same random small model and
quantization config go through apply_quantization_config -> (identical qparams) -> ModelCompressor.from_pretrained_model ->
compress_model -> update_config -> decompress_model in the reference (imported as `compressed_tensors` from a temporary copy) and in
this package (its tensor-level front end rebound to the CPU oracle).  Compared: every module's state dict after compress (keys, dtypes,
shapes, bits), the `quantization_config` written to config.json, and every module's state dict after decompress.

    python tests/reference_compat/fuzz_model_compressor.py [models]
"""
import copy
import json
import os
import random
import sys
import tempfile
import warnings

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, os.path.join(ROOT, "tests", "golden"), ROOT]
from loguru import logger  # noqa: E402

logger.remove()
import make_golden as mg  # noqa: E402,F401
import torch  # noqa: E402

import compressed_tensors as R  # noqa: E402
import compressed_tensors.quantization as RQ  # noqa: E402
from compressed_tensors.quantization.utils import calculate_qparams as r_qparams, generate_gparam as r_gparam  # noqa: E402
from compressed_tensors.utils import get_direct_state_dict as r_state  # noqa: E402

import oracle_patch  # noqa: E402

oracle_patch.apply("compressed_tensors_b200")
import compressed_tensors_b200 as M  # noqa: E402
import compressed_tensors_b200.quantization as MQ  # noqa: E402
from compressed_tensors_b200.utils import get_direct_state_dict as m_state  # noqa: E402

from fuzz_compressors import same_dict  # noqa: E402

PRESETS = ["W4A16", "W4A16_ASYM", "W8A16", "W8A8", "W4A8", "FP8", "FP8_DYNAMIC", "FP8_BLOCK", "NVFP4A16", "NVFP4", "MXFP4A16", "MXFP4"]


def build(rnd, seed):
    torch.manual_seed(seed)
    model = torch.nn.Sequential()
    n = rnd.randint(1, 5)
    for i in range(n):
        model.add_module(f"proj{i}", torch.nn.Linear(128 * rnd.choice([1, 2, 3]), 128 * rnd.choice([1, 2]), bias=rnd.random() < 0.3).to(torch.bfloat16))
    model.add_module("norm", torch.nn.LayerNorm(128).to(torch.bfloat16))
    model.add_module("lm_head", torch.nn.Linear(128, 64, bias=False).to(torch.bfloat16))
    return model


def calibrate(model):
    """memoryless min-max weights observer with the REFERENCE's rule; returns {module name: {param: tensor}} to load into both models"""
    out = {}
    for name, m in model.named_modules():
        scheme = getattr(m, "quantization_scheme", None)
        if scheme is None or scheme.weights is None:
            continue
        a, w = scheme.weights, m.weight.data
        s = a.strategy
        gs = None
        if s == "tensor":
            lo, hi = w.amin().reshape(1), w.amax().reshape(1)
        elif s == "channel":
            lo, hi = w.amin(-1, keepdim=True), w.amax(-1, keepdim=True)
        elif s in ("group", "tensor_group"):
            grp = w.unflatten(-1, (-1, a.group_size))
            lo, hi = grp.amin(-1), grp.amax(-1)
        else:
            bh, bw = a.block_structure
            blk = w.reshape(w.shape[0] // bh, bh, w.shape[1] // bw, bw)
            lo, hi = blk.amin((1, 3)), blk.amax((1, 3))
        q = {}
        if s == "tensor_group":
            gs = r_gparam(w.amin(), w.amax())
            q["weight_global_scale"] = gs
        sc, zp = r_qparams(lo, hi, a, global_scale=gs) if gs is not None else r_qparams(lo, hi, a)
        q["weight_scale"], q["weight_zero_point"] = sc, zp
        for base in ("input", "output"):        # static activation qparams are allocated uninitialised
            for suffix, val in (("scale", 0.5), ("zero_point", 0), ("global_scale", 2.0)):
                if hasattr(m, f"{base}_{suffix}"):
                    q[f"{base}_{suffix}"] = torch.full_like(getattr(m, f"{base}_{suffix}").data, val)
        out[name] = q
    return out


def load_qparams(model, q):
    for name, m in model.named_modules():
        for k, v in q.get(name, {}).items():
            if hasattr(m, k):
                getattr(m, k).data = v.clone().to(getattr(m, k).dtype).reshape(getattr(m, k).shape)


def states(model, fn):
    return {n: {k: v for k, v in fn(m).items()} for n, m in model.named_modules()}


def compare_models(a, b, what):
    sa, sb = states(a, m_state), states(b, r_state)
    for n in sb:
        err = same_dict(sa.get(n, {}), sb[n], f"{what} {n or '<root>'}")
        if err:
            return err
    return None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rnd = random.Random(31)
    checked = bad = 0
    for case in range(n):
        preset = rnd.choice(PRESETS)
        seed = rnd.randint(0, 10 ** 6)
        st = rnd.getstate()
        ref_model = build(rnd, seed)
        rnd.setstate(st)
        my_model = build(rnd, seed)
        cfg = dict(config_groups={preset: ["Linear"]}, ignore=["lm_head"])
        RQ.apply_quantization_config(ref_model, RQ.QuantizationConfig(**cfg))
        MQ.apply_quantization_config(my_model, MQ.QuantizationConfig(**cfg))
        q = calibrate(ref_model)
        load_qparams(ref_model, q)
        load_qparams(my_model, q)
        err = compare_models(my_model, ref_model, "initialized")
        if err is None:
            rc, mc = R.ModelCompressor.from_pretrained_model(ref_model), M.ModelCompressor.from_pretrained_model(my_model)
            rc.compress_model(ref_model)
            mc.compress_model(my_model)
            err = compare_models(my_model, ref_model, "compressed")
        if err is None:
            with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2:
                rc.update_config(d1)
                mc.update_config(d2)
                c1 = json.load(open(os.path.join(d1, "config.json")))["quantization_config"]
                c2 = json.load(open(os.path.join(d2, "config.json")))["quantization_config"]
                c1.pop("version", None), c2.pop("version", None)
                if c1 != c2:
                    err = "config.json: " + json.dumps({k: (c1.get(k), c2.get(k)) for k in set(c1) | set(c2) if c1.get(k) != c2.get(k)})[:600]
        if err is None:
            def back(mcomp, model):
                try:
                    mcomp.decompress_model(model)
                    return None
                except Exception as e:  # noqa: BLE001  (e.g. FP8_BLOCK with an [N, 1] scale grid: dequantize infers CHANNEL and the broadcast fails)
                    return type(e).__name__

            r_err, m_err = back(rc, ref_model), back(mc, my_model)
            if r_err or m_err:
                err = None if (r_err and m_err) else f"decompress_model: reference {r_err or 'ok'}, mirror {m_err or 'ok'}"
            else:
                err = compare_models(my_model, ref_model, "decompressed")
        checked += 1
        if err:
            bad += 1
            if bad <= 8:
                print(f"model {case} {preset}: {err}")
    print(f"model_compressor: {checked} models checked, {bad} mismatches", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

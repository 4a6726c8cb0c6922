"""
Pins the FP4 / MX part of the CPU oracle (oracle/ct_oracle_fp4.c and the global-scale / Q_FP4 paths of
oracle/ct_oracle.c) against golden vectors produced by running the reference (tests/golden/make_golden_fp4.py).
Bit-exact, including the sign of zero and NaN positions.
"""
import pytest
import torch

import oracle
from tests.golden import load
from tests.util import same, same_nan

G = load("fp4")


def kw(args, **extra):
    return dict(strategy=args["strategy"], group_size=args["group_size"], block_structure=args["block_structure"],
                num_bits=args["num_bits"], qtype=args["type"], **extra)


@pytest.mark.parametrize("i", range(len(G["cast"])))
def test_cast_to_fp4(i):
    c = G["cast"][i]
    same_nan(oracle.cast_to_fp4(c["x"]), c["y"], "cast_to_fp4")


@pytest.mark.parametrize("i", range(len(G["pack"])))
def test_pack_unpack_fp4(i):
    c = G["pack"][i]
    same(oracle.pack_fp4_to_uint8(c["x"]), c["packed"], "pack_fp4_to_uint8")
    m, n = c["x"].shape
    for name, want in c["unpacked"].items():
        dt = getattr(torch, name.split(".")[1])
        same(oracle.unpack_fp4_from_uint8(c["packed"], m, n, dt), want, f"unpack_fp4_from_uint8 -> {name}")


def test_pack_fp4_odd_columns():
    with pytest.raises(ValueError):
        oracle.pack_fp4_to_uint8(torch.zeros(3, 7))


@pytest.mark.parametrize("i", range(len(G["nvfp4"])))
def test_nvfp4_quantize_family(i):
    c = G["nvfp4"][i]
    a, gs = c["args"], c["global_scale"]
    zp = c["qparams_zp"]
    same(oracle.quantize(c["x"], c["scale"], zp, global_scale=gs, **kw(a)), c["q"], "nvfp4 quantize")
    same(oracle.fake_quantize(c["x"], c["scale"], zp, global_scale=gs, **kw(a)), c["fq"], "nvfp4 fake_quantize")
    same(oracle.dequantize(c["q"], c["scale"], None, global_scale=gs, dtype=c["x"].dtype), c["dq"], "nvfp4 dequantize")


@pytest.mark.parametrize("i", range(len(G["mx"])))
def test_mx_quantize_family(i):
    c = G["mx"][i]
    a = c["args"]
    zp = c["qparams_zp"]
    dt = torch.float8_e4m3fn if a["num_bits"] == 8 else None
    same(oracle.quantize(c["x"], c["scale"], zp, dtype=dt, **kw(a)), c["q"], "mx quantize")
    same(oracle.fake_quantize(c["x"], c["scale"], zp, **kw(a)), c["fq"], "mx fake_quantize")
    same(oracle.dequantize(c["q"], c["scale"], None, dtype=c["x"].dtype), c["dq"], "mx dequantize")


@pytest.mark.parametrize("i", range(len(G["e8m0"])))
def test_e8m0_scales(i):
    c = G["e8m0"][i]
    same(oracle.compress_mx_scale(c["scale"]), c["enc"], "compress_mx_scale")
    same(oracle.decompress_mx_scale(c["enc"]), c["dec"], "decompress_mx_scale")


# ---- host mirror: observer-side helpers of the FP4 / MX schemes (plain torch on qparam-sized tensors, run on CPU here) ----
def _args(d):
    from compressed_tensors_b200.quantization import QuantizationArgs

    return QuantizationArgs(**d)


@pytest.mark.parametrize("i", range(len(G["nvfp4"])))
def test_host_qparams_nvfp4(i):
    from compressed_tensors_b200.quantization.utils import calculate_qparams, generate_gparam

    c = G["nvfp4"][i]
    a = _args(c["args"])
    same(generate_gparam(c["x"].min(), c["x"].max()), c["global_scale"], "generate_gparam")
    s, z = calculate_qparams(c["qparams_min"], c["qparams_max"], a, global_scale=c["global_scale"])
    same(s, c["qparams_scale"], "calculate_qparams scale (fp8-rounded, global scale)")
    same(z, c["qparams_zp"], "calculate_qparams zero point")


@pytest.mark.parametrize("i", range(len(G["mx"])))
def test_host_qparams_mx(i):
    from compressed_tensors_b200.quantization.utils import calculate_qparams

    c = G["mx"][i]
    # the reference's model_dump drops zp_dtype of symmetric args; the generator used uint8 (make_golden_fp4.py MX4 / MX8)
    s, z = calculate_qparams(c["qparams_min"], c["qparams_max"], _args({**c["args"], "zp_dtype": torch.uint8}))
    same(s, c["scale"], "calculate_qparams MX scale (power of two)")
    same(z, c["qparams_zp"], "calculate_qparams MX zero point")


def test_fp4_formats_are_registered_and_inferred():
    from compressed_tensors_b200.compressors import BaseCompressor, MXFP4PackedCompressor, MXFP8QuantizationCompressor, NVFP4PackedCompressor
    from compressed_tensors_b200.compressors.format import infer_module_format
    from compressed_tensors_b200.quantization import preset_name_to_scheme

    for preset, cls, fmt in (("NVFP4A16", NVFP4PackedCompressor, "nvfp4-pack-quantized"), ("MXFP4A16", MXFP4PackedCompressor, "mxfp4-pack-quantized"),
                             ("MXFP8A16", MXFP8QuantizationCompressor, "mxfp8-quantized")):
        scheme = preset_name_to_scheme(preset, ["Linear"])
        assert infer_module_format(torch.nn.Linear, scheme).value == fmt
        assert BaseCompressor.get_value_from_registry(fmt) is cls
        assert cls.can_compress(torch.nn.Linear, scheme)
        assert cls.compression_param_names(scheme)[0] in ("weight_packed", "weight")
    nv = preset_name_to_scheme("NVFP4", ["Linear"])
    assert "weight_global_scale" in NVFP4PackedCompressor.compression_param_names(nv)
    assert "weight_global_scale" not in MXFP4PackedCompressor.compression_param_names(preset_name_to_scheme("MXFP4", ["Linear"]))

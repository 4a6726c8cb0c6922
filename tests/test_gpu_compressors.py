"""
GPU parity at the plugin boundary: the compressor classes (same names / call signatures as the
reference) against the state dicts the reference itself produced (tests/golden/compressors.pt.gz),
then ModelCompressor on a small model: whole-model multi-tensor launches == per-module plugin path.
"""
import copy
import json

import pytest
import torch

import oracle
from compressed_tensors_b200 import _native as N
from compressed_tensors_b200.compressors import BaseCompressor, ModelCompressor, PackedQuantizationCompressor, compress_module, decompress_module
from compressed_tensors_b200.quantization import (
    QuantizationArgs,
    QuantizationConfig,
    QuantizationScheme,
    QuantizationStatus,
    apply_quantization_config,
    calculate_qparams,
    fake_quantize,
    preset_name_to_scheme,
)
from compressed_tensors_b200.utils import get_direct_state_dict
from tests.golden import load
from tests.util import same, same_values

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_C = load("compressors")


def _to(sd, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}


@pytest.mark.parametrize("c", _C, ids=lambda c: f"{c['format']}-{c['tag']}")
def test_compressor_golden(c):
    scheme = QuantizationScheme.model_validate(c["scheme"])
    comp = BaseCompressor.get_value_from_registry(c["format"])
    sd = _to(c["state_dict"], DEV)
    before = {k: v.clone() for k, v in sd.items()}
    got = comp.compress(sd, scheme)
    assert set(sd) == set(before) and all(torch.equal(sd[k], before[k]) for k in sd), "input state dict was mutated"
    want = c["compressed"]
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in want:
        g = got[k].cpu()
        assert g.dtype == want[k].dtype and g.shape == want[k].shape, (k, g.dtype, g.shape, want[k].dtype, want[k].shape)
        if g.dtype == torch.float8_e4m3fn:
            same_values(g.view(torch.uint8), want[k].view(torch.uint8), f"compress[{k}]")
        else:
            same(g.contiguous(), want[k].contiguous(), f"compress[{k}]")
    back = comp.decompress(_to(want, DEV), scheme)
    assert set(back) == set(c["decompressed"]), (sorted(back), sorted(c["decompressed"]))
    for k, w in c["decompressed"].items():
        g = back[k].cpu()
        assert g.dtype == w.dtype and g.shape == w.shape, k
        if g.dtype == torch.float8_e4m3fn:
            same_values(g.view(torch.uint8), w.view(torch.uint8), f"decompress[{k}]")
        else:
            same(g.contiguous(), w.contiguous(), f"decompress[{k}]")


def test_cpu_state_dict_is_served_through_the_gpu():
    """a CPU-resident caller (the reference's usual flow) gets CPU tensors back, computed on the B200"""
    c = next(c for c in _C if c["tag"] == "W4A16")
    scheme = QuantizationScheme.model_validate(c["scheme"])
    before = N.launch_count()
    got = PackedQuantizationCompressor.compress(c["state_dict"], scheme)
    assert N.launch_count() > before
    assert not got["weight_packed"].is_cuda
    same_values(got["weight_packed"], c["compressed"]["weight_packed"], "cpu->gpu->cpu")


def _model(dtype=torch.bfloat16):
    torch.manual_seed(0)

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = torch.nn.Linear(512, 256, bias=False)
            self.up_proj = torch.nn.Linear(512, 1024, bias=False)
            self.down_proj = torch.nn.Linear(1024, 512, bias=True)

        def forward(self, x):
            return self.down_proj(torch.nn.functional.silu(self.up_proj(x))) + self.q_proj(x).repeat(1, 2)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.ModuleList([Block() for _ in range(3)])
            self.lm_head = torch.nn.Linear(512, 100, bias=False)

        def forward(self, x):
            for l in self.layers:
                x = l(x)
            return self.lm_head(x)

    return M().to(dtype).to(DEV)


def _calibrate(model):
    """min/max observer, like the reference's test fixtures (tests/conftest.py:21-102)"""
    for m in model.modules():
        scheme = getattr(m, "quantization_scheme", None)
        if scheme is None or scheme.weights is None:
            continue
        a, w = scheme.weights, m.weight.data.float()
        if a.strategy == "group":
            wr = w.unflatten(-1, (-1, a.group_size))
            mn, mx = wr.amin(-1), wr.amax(-1)
        elif a.strategy == "channel":
            mn, mx = w.amin(-1, keepdim=True), w.amax(-1, keepdim=True)
        else:
            mn, mx = w.min().reshape(1), w.max().reshape(1)
        s, z = calculate_qparams(mn, mx, a)
        m.weight_scale.data = s.to(m.weight.dtype)
        if hasattr(m, "weight_zero_point"):
            m.weight_zero_point.data = z.to(m.weight_zero_point.dtype)
        if hasattr(m, "input_scale"):           # static activation scales (FP8 preset) are created uninitialised
            m.input_scale.data.fill_(0.05)
        m.quantization_status = QuantizationStatus.FROZEN


@pytest.mark.parametrize("preset", ["W4A16", "W4A16_ASYM", "W8A16", "FP8_DYNAMIC", "W8A8", "FP8"])
def test_model_compressor_batched_equals_per_module(preset, tmp_path):
    model = _model()
    cfg = QuantizationConfig(config_groups={preset: ["Linear"]}, ignore=["lm_head"])
    apply_quantization_config(model, cfg)
    _calibrate(model)
    if preset.startswith("FP8"):
        # -0.0 weights (pruned checkpoints have them): x / s + zero_point turns -0.0 into +0.0, i.e. fp8 byte 0x00 and not 0x80 --
        # the float8 zero point of these presets must go through the whole-model path exactly as through the plugin path
        for l in model.layers:
            l.q_proj.weight.data[::3, ::5] = -0.0
    ref = copy.deepcopy(model)
    x = torch.randn(4, 512, device=DEV, dtype=torch.bfloat16)
    want_out = model(x)  # fake-quantized forward (weight QDQ on the fly)

    # per-module plugin path on the copy
    for m in ref.modules():
        compress_module(m)
    # whole-model path
    mc = ModelCompressor.from_pretrained_model(model)
    launches = N.launch_count()
    mc.compress_model(model)
    used = N.launch_count() - launches
    n_mod = sum(1 for m in model.modules() if getattr(m, "quantization_status", None) == QuantizationStatus.COMPRESSED)
    assert n_mod == 9
    assert used < n_mod or preset.endswith("ASYM"), f"{used} launches for {n_mod} modules: expected multi-tensor launches"
    for (n1, m1), (n2, m2) in zip(model.named_modules(), ref.named_modules()):
        s1, s2 = get_direct_state_dict(m1), get_direct_state_dict(m2)
        assert set(s1) == set(s2), (n1, sorted(s1), sorted(s2))
        for k in s1:
            if s1[k] is None:
                continue
            a, b = s1[k], s2[k]
            assert a.dtype == b.dtype and a.shape == b.shape, (n1, k)
            if a.dtype == torch.float8_e4m3fn:
                same_values(a.view(torch.uint8), b.view(torch.uint8), f"{n1}.{k}")   # bytes: -0.0 (0x80) != +0.0 (0x00)
            else:
                same(a, b, f"{n1}.{k}")
    assert mc.quantization_config.quantization_status == QuantizationStatus.COMPRESSED
    mc.update_config(str(tmp_path))
    cfgj = json.load(open(tmp_path / "config.json"))["quantization_config"]
    assert cfgj["quantization_status"] == "compressed"

    # the decompress-on-first-forward hook restores a runnable model whose weights equal fake_quantize(w)
    assert hasattr(model, "ct_decompress_hook")
    out = model(x)
    assert not hasattr(model, "ct_decompress_hook")
    for m in model.modules():
        if getattr(m, "quantization_scheme", None) is not None:
            assert m.quantization_status == QuantizationStatus.DECOMPRESSED and m.weight.dtype == torch.bfloat16
    same_values(out, want_out, "forward after decompress == fake-quantized forward")
    # and it can be compressed again to the same bytes (idempotence)
    mc.compress_model(model)
    for (n1, m1), (n2, m2) in zip(model.named_modules(), ref.named_modules()):
        for k, v in get_direct_state_dict(m2).items():
            if v is not None and v.dtype in (torch.int32, torch.int8, torch.float8_e4m3fn) and k != "weight_zero_point":
                a = get_direct_state_dict(m1)[k]
                same_values(a.view(torch.uint8) if a.dtype == torch.float8_e4m3fn else a,
                            v.view(torch.uint8) if v.dtype == torch.float8_e4m3fn else v, f"recompress {n1}.{k}")

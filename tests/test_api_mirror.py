"""
CPU tests of the host-side mirror of the reference's plugin / config API (no kernels run):
schema round trips against dumps produced by the reference, registry rules, format inference,
compressor key sets on meta tensors, ModelCompressor orchestration on meta models, and the
2-rank module-parallel recouple over gloo.
"""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from compressed_tensors_b200.compressors import (
    BaseCompressor,
    DenseCompressor,
    FloatQuantizationCompressor,
    IntQuantizationCompressor,
    ModelCompressor,
    NaiveQuantizationCompressor,
    PackedQuantizationCompressor,
    compress_module,
    decompress_module,
)
from compressed_tensors_b200.compressors.format import infer_model_format, infer_module_format
from compressed_tensors_b200.config import BitmaskConfig, CompressionFormat, Sparse24BitMaskConfig, SparsityCompressionConfig, SparsityStructure
from compressed_tensors_b200.distributed import greedy_bin_packing
from tests.util import free_port
from compressed_tensors_b200.quantization import (
    ActivationOrdering,
    QuantizationArgs,
    QuantizationConfig,
    QuantizationScheme,
    QuantizationStatus,
    QuantizationStrategy,
    apply_quantization_config,
    calculate_qparams,
    calculate_range,
    initialize_module_for_quantization,
    is_preset_scheme,
    preset_name_to_scheme,
)
from compressed_tensors_b200.registry import RegistryMixin
from compressed_tensors_b200.utils import ImplBackend, get_direct_state_dict, getattr_chain, replace_direct_state_dict
from tests.golden import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_C = load("compressors")


# ------------------------------------------------------------------------------------------------
# schema
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(len(_C)))
def test_scheme_dump_round_trips_reference_json(i):
    """QuantizationScheme.model_validate(reference dump).model_dump(json) == reference dump"""
    ref = _C[i]["scheme"]
    mine = QuantizationScheme.model_validate(ref).model_dump(mode="json")
    assert mine == ref, (mine, ref)


@pytest.mark.parametrize("i", range(len(_C)))
def test_param_names_and_format_match_reference(i):
    c = _C[i]
    scheme = QuantizationScheme.model_validate(c["scheme"])
    comp = BaseCompressor.get_value_from_registry(c["format"])
    assert list(comp.compression_param_names(scheme)) == c["param_names"]
    assert comp.compression_param_names(scheme)[0] in ("weight", "weight_packed")
    if c["tag"] in ("W4A16", "W4A16_ASYM", "W8A16", "FP8", "W8A8", "FP8_DYNAMIC", "FP8_BLOCK"):
        assert infer_module_format(torch.nn.Linear, preset_name_to_scheme(c["tag"], ["Linear"])).value == c["format"]


def test_quantization_args_validation():
    assert QuantizationArgs().strategy == "tensor"
    assert QuantizationArgs(group_size=128).strategy == "group"
    assert QuantizationArgs(group_size=-1).strategy == "channel"
    assert QuantizationArgs(num_bits=4).zp_dtype == torch.int8
    assert QuantizationArgs(num_bits=8, type="float").zp_dtype == torch.float8_e4m3fn
    assert QuantizationArgs(num_bits=8, type="float").pytorch_dtype() == torch.float8_e4m3fn
    assert QuantizationArgs(num_bits=4).pytorch_dtype() == torch.int8
    assert QuantizationArgs(block_structure="128x128", strategy="block").block_structure == [128, 128]
    assert QuantizationArgs(strategy="group", group_size=128, actorder=True).actorder == "group"
    static = QuantizationArgs(strategy="group", group_size=128, actorder="static").actorder
    assert static == "static" and static == ActivationOrdering.WEIGHT and ActivationOrdering.DYNAMIC == ActivationOrdering.GROUP   # aliases, as in the reference
    assert QuantizationArgs(dynamic=True, strategy="token").observer is None
    assert QuantizationArgs().observer == "memoryless_minmax"
    for bad in (dict(strategy="token"), dict(strategy="group"), dict(group_size=64, strategy="channel"), dict(group_size=-2),
                dict(strategy="block"), dict(block_structure=[1, 2]), dict(strategy="tensor", actorder="group"),
                dict(dynamic=True, strategy="channel")):
        with pytest.raises(ValueError):
            QuantizationArgs(**bad)
    with pytest.raises(ValueError):
        QuantizationArgs(foo=1)


def test_presets_and_config():
    assert is_preset_scheme("w4a16") and not is_preset_scheme("nope")
    with pytest.raises(KeyError):
        preset_name_to_scheme("nope", ["Linear"])
    w4 = preset_name_to_scheme("W4A16", ["Linear"])
    assert w4.weights.num_bits == 4 and w4.weights.group_size == 128 and w4.weights.symmetric and w4.input_activations is None
    cfg = QuantizationConfig(config_groups={"W8A8": ["Linear"]}, ignore=["lm_head"])
    assert isinstance(cfg.config_groups["W8A8"], QuantizationScheme)
    assert cfg.config_groups["W8A8"].input_activations.dynamic is True
    d = json.loads(json.dumps(cfg.model_dump(mode="json")))
    again = QuantizationConfig.model_validate(d)
    assert again.config_groups["W8A8"].weights.strategy == "channel"
    assert cfg.requires_calibration_data() is False
    assert QuantizationConfig(config_groups={"FP8": ["Linear"]}).requires_calibration_data() is True


def test_status_ordering():
    S = QuantizationStatus
    assert S.INITIALIZED < S.CALIBRATION < S.FROZEN < S.COMPRESSED < S.DECOMPRESSED
    assert S.COMPRESSED >= S.COMPRESSED and S.FROZEN >= None and not (S.FROZEN < None)
    assert S("compressed") is S.COMPRESSED


def test_calculate_range_and_qparams():
    lo, hi = calculate_range(QuantizationArgs(num_bits=4), "cpu")
    assert (lo.item(), hi.item()) == (-8.0, 7.0) and lo.dtype == torch.float32 and lo.ndim == 0
    lo, hi = calculate_range(QuantizationArgs(num_bits=8, type="float"), "cpu")
    assert (lo.item(), hi.item()) == (-448.0, 448.0)
    a = QuantizationArgs(num_bits=4, strategy="group", group_size=128)
    mn, mx = torch.tensor([[-0.3, -1.0]]), torch.tensor([[0.75, 0.5]])
    s, z = calculate_qparams(mn, mx, a)
    assert torch.equal(s, torch.tensor([[0.75 / 7.5, 1.0 / 7.5]])) and z.dtype == torch.int8 and not z.any()
    a = QuantizationArgs(num_bits=8, symmetric=False, strategy="channel")
    s, z = calculate_qparams(torch.tensor([[-1.0]]), torch.tensor([[3.0]]), a)
    assert torch.allclose(s, torch.tensor([[4.0 / 255]])) and z.item() == round(-128 + 1.0 / (4.0 / 255))
    s, z = calculate_qparams(torch.zeros(1), torch.zeros(1), QuantizationArgs(num_bits=8))
    assert s.item() == torch.finfo(torch.float32).eps


def test_compression_format_and_sparsity_config():
    assert CompressionFormat("pack-quantized") is CompressionFormat.pack_quantized
    assert SparsityStructure("2:4") is SparsityStructure.TWO_FOUR and SparsityStructure(None) is SparsityStructure.UNSTRUCTURED
    assert SparsityStructure("UNSTRUCTURED") is SparsityStructure.UNSTRUCTURED
    with pytest.raises(ValueError):
        SparsityStructure("3:7")
    c = SparsityCompressionConfig.load_from_registry("sparse-24-bitmask")
    assert isinstance(c, Sparse24BitMaskConfig) and c.sparsity_structure == "2:4"
    assert isinstance(SparsityCompressionConfig.load_from_registry("sparse_bitmask"), BitmaskConfig)


# ------------------------------------------------------------------------------------------------
# registry
# ------------------------------------------------------------------------------------------------
def test_registry_rules(tmp_path):
    class Base(RegistryMixin):
        pass

    @Base.register(name="Foo_bar baz", alias=["fb", "F B2"])
    class A(Base):
        pass

    assert Base.get_value_from_registry("foo-bar-baz") is A and Base.get_value_from_registry("FOO_BAR_BAZ") is A
    assert Base.get_value_from_registry("fb") is A and Base.get_value_from_registry("f-b2") is A
    assert "foo-bar-baz" in Base.registered_names() and set(Base.registered_aliases()) == {"fb", "f-b2"}
    Base.register_value(A, name="foo-bar-baz-again")  # same value under another name is fine
    with pytest.raises((RuntimeError, KeyError)):
        @Base.register(name="foo bar baz")
        class B(Base):
            pass
    with pytest.raises(KeyError):
        Base.get_value_from_registry("missing")
    plugin = tmp_path / "plug.py"
    plugin.write_text("class MyThing:\n    x = 7\n")
    assert Base.get_value_from_registry(f"{plugin}:MyThing").x == 7
    assert isinstance(Base.load_from_registry("fb"), A)


def test_compressor_registry_contents():
    names = set(BaseCompressor.registered_names())
    assert {"dense", "naive-quantized", "int-quantized", "float-quantized", "pack-quantized", "sparse-24-bitmask", "sparse-bitmask"} <= names
    assert BaseCompressor.get_value_from_registry("pack_quantized") is PackedQuantizationCompressor
    assert BaseCompressor.get_value_from_registry(CompressionFormat.float_quantized.value) is FloatQuantizationCompressor
    assert issubclass(IntQuantizationCompressor, NaiveQuantizationCompressor)
    with pytest.raises((RuntimeError, KeyError)):  # a name can be claimed once (registry.py:215-223, :296-303)
        @BaseCompressor.register(name="pack-quantized")
        class Impostor(BaseCompressor):
            pass


def test_impl_backend_surface():
    calls = []

    @ImplBackend.register("_ct_b200_test_op", lambda x: x > 0, "0")
    def fast(x):
        calls.append("fast")
        return x * 2

    @ImplBackend.register("_ct_b200_test_op", lambda x: True, "disable")
    def disabled(x):
        calls.append("disabled")
        return -1

    @ImplBackend.entrypoint("_ct_b200_test_op")
    def op(x):
        calls.append("body")
        return x

    assert op(3) == 6 and op(-3) == -3 and calls == ["fast", "body"]
    assert ImplBackend.call("fast", 2) == 4 and ImplBackend.call("op", 2) == 2     # by function name, bypassing the checks
    with pytest.raises(KeyError):
        ImplBackend.call("disabled", 1)
    with pytest.raises(ValueError, match="already registered"):
        ImplBackend.register("_ct_b200_other", lambda x: True, "0")(fast)
    os.environ["CT_ENFORCE_EAGER"] = "1"
    try:
        assert op(1) == 1, "CT_ENFORCE_EAGER: the entrypoint's own (eager) body runs even though a backend accepts the arguments"
    finally:
        del os.environ["CT_ENFORCE_EAGER"]


# ------------------------------------------------------------------------------------------------
# compressors on meta tensors / module plumbing
# ------------------------------------------------------------------------------------------------
def test_can_compress_matrix():
    L = torch.nn.Linear
    w4, fp8, w8a8, fp8d = (preset_name_to_scheme(n, ["Linear"]) for n in ("W4A16", "FP8", "W8A8", "FP8_DYNAMIC"))
    assert PackedQuantizationCompressor.can_compress(L, w4) and not PackedQuantizationCompressor.can_compress(L, fp8)
    assert not PackedQuantizationCompressor.can_compress(torch.nn.Conv2d, w4)
    assert PackedQuantizationCompressor.can_compress(L, w8a8)   # int weights, int activations
    assert IntQuantizationCompressor.can_compress(L, w8a8) and not IntQuantizationCompressor.can_compress(L, w4)
    assert FloatQuantizationCompressor.can_compress(L, fp8) and FloatQuantizationCompressor.can_compress(L, fp8d)
    w4afp8 = preset_name_to_scheme("W4AFP8", ["Linear"])
    assert not PackedQuantizationCompressor.can_compress(L, w4afp8)
    assert infer_module_format(L, w4afp8) == CompressionFormat.int_quantized
    fp8_w_only = QuantizationScheme(targets=["Linear"], weights=QuantizationArgs(num_bits=8, type="float"))
    assert infer_module_format(L, fp8_w_only) == CompressionFormat.naive_quantized
    assert infer_module_format(L, QuantizationScheme(targets=["Linear"])) == CompressionFormat.dense
    assert DenseCompressor.can_compress(L, w4)


# produced by running the reference's infer_module_format over its PRESET_SCHEMES (compressors/format.py:75-96)
_REF_PRESET_FORMATS = {
    'W4A16_ASYM': 'pack-quantized', 'W8A8': 'int-quantized', 'INT8': 'int-quantized', 'W4AFP8': 'int-quantized', 'FP8': 'float-quantized',
    'FP8_DYNAMIC': 'float-quantized', 'FP8_BLOCK': 'float-quantized', 'NVFP4A16': 'nvfp4-pack-quantized', 'NVFP4': 'nvfp4-pack-quantized',
    'MXFP4A16': 'mxfp4-pack-quantized', 'MXFP4': 'mxfp4-pack-quantized', 'MXFP8A16': 'mxfp8-quantized', 'MXFP8': 'mxfp8-quantized',
    'W2A4': 'int-quantized', 'W2A8': 'int-quantized', 'W2A16': 'pack-quantized', 'W3A4': 'int-quantized', 'W3A8': 'int-quantized',
    'W3A16': 'pack-quantized', 'W4A4': 'int-quantized', 'W4A8': 'int-quantized', 'W4A16': 'pack-quantized', 'W5A8': 'int-quantized',
    'W5A16': 'pack-quantized', 'W6A8': 'int-quantized', 'W6A16': 'pack-quantized', 'W7A8': 'int-quantized', 'W7A16': 'pack-quantized',
    'W8A16': 'pack-quantized',
}


def test_every_preset_infers_the_reference_format():
    for name, fmt in _REF_PRESET_FORMATS.items():
        scheme = preset_name_to_scheme(name, ["Linear"])
        assert infer_module_format(torch.nn.Linear, scheme).value == fmt, name


def test_pack_compressor_meta_path():
    scheme = preset_name_to_scheme("W4A16_ASYM", ["Linear"])
    sd = {"weight": torch.empty(64, 512, dtype=torch.bfloat16, device="meta"),
          "weight_scale": torch.empty(64, 4, dtype=torch.bfloat16, device="meta"),
          "weight_zero_point": torch.empty(64, 4, dtype=torch.int8, device="meta")}
    out = PackedQuantizationCompressor.compress(sd, scheme)
    assert set(sd) == {"weight", "weight_scale", "weight_zero_point"}, "input must not be mutated"
    assert set(out) == {"weight_packed", "weight_shape", "weight_scale", "weight_zero_point"}
    assert out["weight_packed"].shape == (64, 64) and out["weight_packed"].dtype == torch.int32
    assert out["weight_shape"].tolist() == [64, 512]
    assert out["weight_zero_point"].shape == (8, 4) and out["weight_zero_point"].dtype == torch.int32
    back = PackedQuantizationCompressor.decompress(out, scheme)
    assert back["weight_zero_point"].shape == (64, 4) and back["weight_zero_point"].dtype == torch.int8
    assert back["weight"].shape == (64, 512) and back["weight"].dtype == torch.bfloat16 and back["weight"].device.type == "meta"
    sym = PackedQuantizationCompressor.compress(sd, preset_name_to_scheme("W4A16", ["Linear"]))
    assert "weight_zero_point" not in sym


def _tiny_model(device="meta", dtype=torch.bfloat16):
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = torch.nn.Linear(256, 128, bias=False, device=device, dtype=dtype)
            self.up_proj = torch.nn.Linear(256, 512, bias=False, device=device, dtype=dtype)
            self.norm = torch.nn.LayerNorm(256, device=device, dtype=dtype)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.ModuleList([Block(), Block()])
            self.lm_head = torch.nn.Linear(256, 1000, bias=False, device=device, dtype=dtype)

    return M()


def test_apply_config_and_model_compressor_on_meta(tmp_path):
    model = _tiny_model()
    cfg = QuantizationConfig(config_groups={"W4A16": ["Linear"]}, ignore=["lm_head"])
    apply_quantization_config(model, cfg)
    q = model.layers[0].q_proj
    assert q.quantization_status == QuantizationStatus.INITIALIZED
    assert q.weight_scale.shape == (128, 2) and q.weight_scale.dtype == torch.bfloat16
    assert q.weight_zero_point.dtype == torch.int8
    assert not hasattr(model.lm_head, "quantization_scheme")
    mc = ModelCompressor.from_pretrained_model(model)
    assert mc.quantization_config.format == "pack-quantized"
    assert sorted(mc.quantization_config.ignore) == ["lm_head"]
    mc.compress_model(model)
    assert q.quantization_status == QuantizationStatus.COMPRESSED
    assert {k for k, v in get_direct_state_dict(q).items() if v is not None} == {"weight_packed", "weight_shape", "weight_scale"}
    assert q.weight_packed.shape == (128, 32) and not q.weight_packed.requires_grad
    assert mc.quantization_config.quantization_status == QuantizationStatus.COMPRESSED
    assert hasattr(model, "ct_decompress_hook")
    mc.update_config(str(tmp_path))
    data = json.load(open(tmp_path / "config.json"))["quantization_config"]
    assert data["quant_method"] == "compressed-tensors" and data["format"] == "pack-quantized"
    assert data["quantization_status"] == "compressed" and data["sparsity_config"] == {}
    assert data["config_groups"]["group_0"]["weights"]["num_bits"] == 4
    mc2 = ModelCompressor.from_compression_config({"quantization_config": data})
    assert mc2.quantization_config.config_groups["group_0"].weights.group_size == 128
    mc.decompress_model(model)
    assert q.quantization_status == QuantizationStatus.DECOMPRESSED and q.weight.shape == (128, 256)
    assert not hasattr(model, "ct_decompress_hook")


def test_state_dict_replacement_rules():
    m = torch.nn.Linear(4, 4)
    keep = m.bias.data
    replace_direct_state_dict(m, {"bias": keep, "weight_packed": torch.zeros(4, 1, dtype=torch.int32)})
    assert set(get_direct_state_dict(m)) == {"bias", "weight_packed"} and m.bias.data_ptr() == keep.data_ptr()
    assert isinstance(m.weight_packed, torch.nn.Parameter) and not m.weight_packed.requires_grad
    assert getattr_chain(m, "weight_packed.dtype") == torch.int32
    assert getattr_chain(m, "nope.deeper", None) is None
    with pytest.raises(AttributeError):
        getattr_chain(m, "nope.deeper")


def test_greedy_bin_packing():
    items = [("a", 5), ("b", 9), ("c", 3), ("d", 7), ("e", 1)]
    ordered, bins, owner = greedy_bin_packing(list(items), 2, lambda it: it[1])
    assert [i[0] for i in ordered] == ["b", "d", "a", "c", "e"]
    assert [[i[0] for i in b] for b in bins] == [["b", "c", "e"], ["d", "a"]]
    assert owner[("a", 5)] == 1


# ------------------------------------------------------------------------------------------------
# 2-rank module-parallel recouple over gloo (host logic of the multi-GPU path, no kernels)
# ------------------------------------------------------------------------------------------------
_WORKER = textwrap.dedent(
    """
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from compressed_tensors_b200.distributed import replace_module_parallel, is_distributed, greedy_bin_packing, module_size
    from compressed_tensors_b200.utils import get_direct_state_dict, replace_direct_state_dict
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    assert is_distributed()
    torch.manual_seed(0)
    mods = [torch.nn.Linear(16 * (i + 1), 8, bias=False) for i in range(5)]
    calls = []
    def fake_compress(m):
        sd = get_direct_state_dict(m)
        w = sd.pop("weight")
        if w.device.type == "meta":
            sd["weight_q"] = torch.empty(w.shape, dtype=torch.float8_e4m3fn, device="meta")
            sd["weight_sum"] = torch.empty(1, dtype=torch.float32, device="meta")
        else:
            calls.append(m)
            sd["weight_q"] = (w * 10 + rank * 0).to(torch.float8_e4m3fn)
            sd["weight_sum"] = w.sum().reshape(1) + 1000 * rank   # proves which rank produced it
        replace_direct_state_dict(m, sd)
    expect_owner = greedy_bin_packing(list(mods), 2, module_size)[2]
    replace_module_parallel(mods, fake_compress)
    mine = [m for m in mods if expect_owner[m] == rank]
    assert calls == [m for m in sorted(mods, key=module_size, reverse=True) if expect_owner[m] == rank], "only own modules are really compressed"
    for m in mods:
        sd = get_direct_state_dict(m)
        assert {k for k, v in sd.items() if v is not None} == {"weight_q", "weight_sum"} and sd["weight_q"].dtype == torch.float8_e4m3fn and sd["weight_q"].device.type == "cpu"
        assert int(sd["weight_sum"].item() // 500) in (2 * expect_owner[m], 2 * expect_owner[m] + 1, 2 * expect_owner[m] - 1)
        flat = torch.cat([sd["weight_q"].view(torch.uint8).flatten().float(), sd["weight_sum"].flatten()])
        other = flat.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(flat, other), "ranks disagree after recouple"
    dist.destroy_process_group()
    print("OK", rank)
    """
)


def test_module_parallel_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", free_port(), str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2


def test_load_pretrained_quantization_parameters(tmp_path):
    """activation qparams always, weight qparams on request, zeros for a missing zero point (apply.py:49-97, :195-236)"""
    from safetensors.torch import save_file

    from compressed_tensors_b200.quantization import QuantizationConfig, apply_quantization_config
    from compressed_tensors_b200.quantization.lifecycle.apply import load_pretrained_quantization_parameters

    model = torch.nn.Sequential(torch.nn.Linear(64, 32, bias=False), torch.nn.LayerNorm(32))
    apply_quantization_config(model, QuantizationConfig(config_groups={"FP8": ["Linear"]}))
    lin = model[0]
    save_file({"0.weight_scale": torch.full_like(lin.weight_scale.data, 0.25), "0.input_scale": torch.full_like(lin.input_scale.data, 0.5),
               "0.weight": torch.zeros(32, 64)}, str(tmp_path / "model.safetensors"))
    lin.input_zero_point.data.fill_(3)
    lin.weight_scale.data.fill_(9)
    load_pretrained_quantization_parameters(model, str(tmp_path))
    assert float(lin.input_scale) == 0.5 and float(lin.input_zero_point.float()) == 0.0 and float(lin.weight_scale.flatten()[0]) == 9.0
    load_pretrained_quantization_parameters(model, str(tmp_path), load_weight_qparams=True)
    assert float(lin.weight_scale.flatten()[0]) == 0.25


def test_apply_config_prefers_the_most_specific_target():
    """a module matching several targets takes the scheme of match_targets()[0]: exact name > regex on the name > class name
    (reference quantization/lifecycle/apply.py:149-151, :258-265), whatever the order of the config groups"""
    import torch
    from compressed_tensors_b200.quantization import QuantizationConfig, apply_quantization_config

    def model():
        m = torch.nn.Sequential()
        for n in ("fc1", "fc2", "fc3"):
            m.add_module(n, torch.nn.Linear(64, 64))
        return m

    groups = {
        "g0": {"targets": ["Linear"], "weights": {"num_bits": 8}},
        "g1": {"targets": ["re:.*fc2"], "weights": {"num_bits": 4, "strategy": "group", "group_size": 32}},
        "g2": {"targets": ["fc3"], "weights": {"num_bits": 2, "strategy": "channel"}},
        "g3": {"targets": ["re:fc3"], "weights": {"num_bits": 6}},
    }
    for order in (["g0", "g1", "g2", "g3"], ["g3", "g2", "g1", "g0"]):
        m = model()
        apply_quantization_config(m, QuantizationConfig(config_groups={k: groups[k] for k in order}))
        assert [getattr(m, n).quantization_scheme.weights.num_bits for n in ("fc1", "fc2", "fc3")] == [8, 4, 2], order
        assert m.fc2.weight_scale.shape == (64, 2) and m.fc3.weight_scale.shape == (64, 1)


def test_kv_cache_scheme_is_refused_loudly():
    import pytest
    import torch
    from compressed_tensors_b200.quantization import QuantizationConfig, apply_quantization_config

    cfg = QuantizationConfig(config_groups={"W8A8": ["Linear"]}, kv_cache_scheme={"num_bits": 8, "type": "float", "symmetric": True})
    with pytest.raises(NotImplementedError, match="KV-cache"):
        apply_quantization_config(torch.nn.Sequential(torch.nn.Linear(32, 32)), cfg)


def test_decompress_model_is_local_unless_asked():
    """reference behaviour: decompress_model (and the first-forward hook) never runs a collective (model_compressor.py:183-207)"""
    import inspect
    from compressed_tensors_b200.compressors import ModelCompressor

    assert inspect.signature(ModelCompressor.decompress_model).parameters["distributed"].default is False
    src = inspect.getsource(ModelCompressor.add_decompress_hook)
    assert "distributed=False" in src

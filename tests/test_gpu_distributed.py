"""
2-GPU test of the module-parallel compress path (one rank per GPU, NCCL): every rank ends with the
same compressed state, bit-identical to a single-process run.  Mirrors the intent of the reference's
tests/test_compressors/distributed/test_distributed_compression.py:111-149 (cross-rank checksums).
Parametrised over the world size: the 2-rank variant needs 2 GPUs, the 1-rank variant runs the same code path (NCCL process group,
bin packing, meta mirror, tensor broadcasts) on any box.  Also runs bench.py's 70B-sharded leg (BASELINE config 5) at a reduced
layer count under the same launches.
"""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch
from tests.util import free_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = textwrap.dedent(
    """
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from compressed_tensors_b200.distributed import init_dist, is_distributed
    from compressed_tensors_b200.compressors import ModelCompressor
    from compressed_tensors_b200.compressors.model_compressors.batched import compress_modules_batched, decompress_modules_batched
    from compressed_tensors_b200.quantization import QuantizationConfig, apply_quantization_config
    from compressed_tensors_b200.utils import get_direct_state_dict
    from compressed_tensors_b200 import _native as N
    sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
    from test_gpu_compressors import _model, _calibrate
    import test_gpu_compressors as T

    init_dist()
    rank = dist.get_rank()
    dev = f"cuda:{rank}"
    T.DEV = dev
    torch.cuda.set_device(rank)
    for preset, where in (("W4A16", dev), ("W4A16_ASYM", dev), ("FP8_DYNAMIC", dev), ("W4A16", "cpu")):   # the last one: host-resident model over NCCL
        T.DEV = where
        model = _model()
        apply_quantization_config(model, QuantizationConfig(config_groups={preset: ["Linear"]}, ignore=["lm_head"]))
        _calibrate(model)
        import copy
        single = copy.deepcopy(model)
        mods = [m for m in single.modules() if getattr(m, "quantization_scheme", None) is not None]
        compress_modules_batched(mods, None)                     # single-process result on this GPU
        launches = N.launch_count()
        mc = ModelCompressor.from_pretrained_model(model)
        mc.compress_model(model)   # distributed path
        mine = N.launch_count() - launches
        assert (0 < mine < 9 or where == "cpu") if dist.get_world_size() > 1 else mine > 0, f"rank {rank} launched {mine} kernels: work was not split"
        sums = []
        for (n1, m1), (n2, m2) in zip(model.named_modules(), single.named_modules()):
            s1, s2 = get_direct_state_dict(m1), get_direct_state_dict(m2)
            assert set(s1) == set(s2), (n1, sorted(s1), sorted(s2))
            for k, v in s1.items():
                if v is None:
                    continue
                assert v.device == torch.device(where) or k == "weight_shape", (n1, k, v.device)
                a = v.view(torch.uint8) if v.dtype == torch.float8_e4m3fn else v
                b = s2[k].view(torch.uint8) if s2[k].dtype == torch.float8_e4m3fn else s2[k]
                assert torch.equal(a, b), f"{preset} {n1}.{k} differs from the single-process result"
                sums.append(a.double().sum().to(dev))
        t = torch.stack(sums)
        other = t.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(t, other), "ranks disagree"
        # and back: distributed decompress_model == single-process decompress, on every rank
        decompress_modules_batched(mods, None)
        mc.decompress_model(model, distributed=True)
        for (n1, m1), (n2, m2) in zip(model.named_modules(), single.named_modules()):
            s1, s2 = get_direct_state_dict(m1), get_direct_state_dict(m2)
            assert set(s1) == set(s2), (n1, sorted(s1), sorted(s2))
            for k, v in s1.items():
                if v is not None:
                    assert v.dtype == s2[k].dtype and torch.equal(v, s2[k]), f"{preset} {n1}.{k}: distributed decompress differs"
                    assert v.device == torch.device(where) or k == "weight_shape", (n1, k, v.device)
    dist.barrier()
    dist.destroy_process_group()
    print("OK", rank)
    """
)


def _torchrun(world: int, args: list, timeout: int = 900):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                           "--master-port", free_port(), *args], capture_output=True, text=True, env=env, timeout=timeout, cwd=ROOT)


def _need(world: int):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")


@pytest.mark.parametrize("world", [1, 2])
def test_distributed_compress_nccl(tmp_path, world):
    _need(world)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    r = _torchrun(world, [str(script), ROOT])
    assert r.returncode == 0, r.stdout[-1500:] + "\n".join(l for l in r.stderr.splitlines() if "rank" in l or "Error" in l)[-6000:]
    assert r.stdout.count("OK") == world


@pytest.mark.parametrize("world", [1, 2])
def test_bench_cfg5_sharded_leg(world):
    """bench.py's BASELINE-config-5 leg at 2 of 80 layers: LPT-sharded 70B-shaped tensors generated on their owner rank only,
    ModelCompressor.compress_model(distributed=True), recouple, distributed decompress; every parity flag must be true"""
    _need(world)
    r = _torchrun(world, [os.path.join(ROOT, "bench.py"), "--gpus", str(world), "--steps", "3", "--warmup", "3", "--layers", "1", "--e2e-layers", "1",
                          "--no-cpu", "--no-extra", "--cfg5", "--cfg5-layers", "2", "--cfg5-reps", "1"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-5000:]
    lines = [line for line in r.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1, lines
    c = json.loads(lines[0])["ops"]["cfg5_70b_sharded"]
    assert c["world_size"] == world and c["tensors"] == 14 and len(c["per_rank_dense_GB"]) == world
    assert c["parity"]["ranks_agree_on_all_checksums"] and c["parity"]["sample_equals_single_rank"] and c["parity"]["oracle_tensor_1"] is True
    assert c["decompress"]["ranks_agree"] and c["decompress"]["sample_equals_fake_quantize"]
    assert c["compress_ms"] > 0 and (c["recouple_bytes_per_rank"] > 0)
    assert c["imbalance_max_over_mean"] < 1.35

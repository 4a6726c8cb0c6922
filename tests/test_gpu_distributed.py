"""
2-GPU test of the module-parallel compress path (one rank per GPU, NCCL): every rank ends with the
same compressed state, bit-identical to a single-process run.  Mirrors the intent of the reference's
tests/test_compressors/distributed/test_distributed_compression.py:111-149 (cross-rank checksums).
Skipped on boxes with fewer than 2 GPUs.
"""
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
        assert 0 < mine < 9 or where == "cpu", f"rank {rank} launched {mine} kernels: work was not split"
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_distributed_compress_two_gpus(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", free_port(), str(script), ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + "\n".join(l for l in r.stderr.splitlines() if "rank" in l or "Error" in l)[-6000:]
    assert r.stdout.count("OK") == 2

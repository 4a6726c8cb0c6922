"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the driver's contract (keys, units, the e2e / cpu_baseline
objects of the arm), standalone and as rank 0 of a 2-rank torchrun launch (the other rank exits 0 without work)."""
import json
import os
import subprocess
import sys
from tests.util import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def _check(line: str, n_gpus: int):
    d = json.loads(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["unit"] == "GB/s" and d["value"] > 0 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "compressed_tensors", "version.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    sys.path.insert(0, ROOT)
    import bench
    phys = bench.cpu_topology()[0]
    assert cb["cores"] == len(phys) == cb["port"]["threads"], "the CPU arm must use every physical core it may (torchrun exports OMP_NUM_THREADS=1)"
    assert cb["reps"] >= 5 and cb["cpu_model"] and cb["nproc"] == len(os.sched_getaffinity(0))
    if have_ref:
        assert cb["reference"]["equals_port_bit_for_bit"] is True and cb["reference"]["threads"] == len(phys)
        assert cb["value"] == cb["reference"]["GBps"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    return d


def test_reference_arm_standalone():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check(r.stdout.strip().splitlines()[-1], 1)
    assert "north_star" in BASE and d["metric"].startswith("weight_GBps")


def test_reference_arm_under_torchrun_prints_once():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", free_port(), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [line for line in r.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1, lines
    _check(lines[0], 2)

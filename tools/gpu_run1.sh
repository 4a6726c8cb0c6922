set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo smoke rc=$?
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2a.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest_r2a.log
python tools/sparse_bench.py > gpurun_out/sparse_r2a.jsonl 2> gpurun_out/sparse_r2a.err; echo sparse rc=$?; cat gpurun_out/sparse_r2a.jsonl; tail -3 gpurun_out/sparse_r2a.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err; echo ref rc=$?
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo bench rc=$?
cat gpurun_out/bench_r2a.json | head -c 6000
tail -5 gpurun_out/bench_r2a.err
lscpu | head -25 > gpurun_out/lscpu.txt; nvidia-smi topo -m > gpurun_out/topo.txt 2>&1

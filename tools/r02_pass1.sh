#!/bin/bash
# round 2, GPU pass 1: gpu tests, default bench (with all sub-records), smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r02_pytest_gpu.log
( time python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
nproc; nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1

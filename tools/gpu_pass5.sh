#!/bin/bash
# GPU pass: full gpu test-suite (incl. parameter gradients), phase profile of C2/C3, quick bench lines
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log
python tools/phase_probe.py c2 > $out/phase_c2.txt 2>&1; cat $out/phase_c2.txt
python tools/phase_probe.py c3 > $out/phase_c3.txt 2>&1; cat $out/phase_c3.txt
b() { name=$1; shift; timeout 400 python bench.py "$@" > $out/bench_$name.json 2> $out/bench_$name.err; cut -c1-400 $out/bench_$name.json; }
b c2q --steps 30 --warmup 5 --no-cpu
b c3q --workload c3 --steps 20 --warmup 5 --no-cpu

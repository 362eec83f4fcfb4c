#!/bin/bash
# GPU pass: full gpu test-suite, smoke, bench lines incl. the EEQ (q=None) variants.
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -3 $out/smoke.log
timeout 300 python bench.py --steps 30 --warmup 5 > $out/bench_c2.json 2> $out/bench_c2.err
timeout 300 python bench.py --steps 30 --warmup 5 --eeq > $out/bench_c2_eeq.json 2> $out/bench_c2_eeq.err
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 > $out/bench_c3.json 2> $out/bench_c3.err
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --eeq > $out/bench_c3_eeq.json 2> $out/bench_c3_eeq.err
timeout 300 python bench.py --workload c5 --steps 20 --warmup 5 > $out/bench_c5_f64.json 2> $out/bench_c5_f64.err
timeout 300 python bench.py --workload c5 --dtype f32 --steps 20 --warmup 5 > $out/bench_c5_f32.json 2> $out/bench_c5_f32.err
timeout 300 python bench.py --workload c1 --steps 50 --warmup 5 > $out/bench_c1.json 2> $out/bench_c1.err
timeout 300 python bench.py --workload c1 --steps 50 --warmup 5 --eeq > $out/bench_c1_eeq.json 2> $out/bench_c1_eeq.err
for f in $out/bench_*.json; do echo "== $f"; cut -c1-400 $f; done
tail -3 $out/*.err

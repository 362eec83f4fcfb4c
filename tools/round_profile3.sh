#!/bin/bash
# Refresh of the artefacts that depend on the gradient kernels (after the sweep rework).
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
b() { name=$1; shift; timeout 300 python bench.py "$@" > $out/bench_$name.json 2> $out/bench_$name.err; cut -c1-160 $out/bench_$name.json; }
b c3 --workload c3 --steps 20 --warmup 5
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.1, .bool.0, .int.100" -s 2 -c 1 -f -o $out/prof_c3_g100 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
b c5_f64 --workload c5 --steps 20 --warmup 5
b c5_f32 --workload c5 --dtype f32 --steps 20 --warmup 5
b c3_eeq --workload c3 --steps 20 --warmup 5 --eeq --no-cpu
python tools/phase_probe.py c3 > $out/phase_c3.txt 2>&1; cat $out/phase_c3.txt

#!/bin/bash
# Final GPU pass of the round, most important artefacts first (budget-bounded): gpu test log, smoke,
# bench lines, ncu launch lists and --set full captures of the dominant kernels.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
b() { name=$1; shift; timeout 300 python bench.py "$@" > $out/bench_$name.json 2> $out/bench_$name.err; cut -c1-160 $out/bench_$name.json; }
b c2 --steps 50 --warmup 5
b c3 --workload c3 --steps 20 --warmup 5
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.1, .bool.0, .int.100" -s 2 -c 1 -f -o $out/prof_c3_g100 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.0, .bool.0, .int.64" -s 2 -c 1 -f -o $out/prof_c2_e64 python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
b c5_f64 --workload c5 --steps 20 --warmup 5
b c5_f32 --workload c5 --dtype f32 --steps 20 --warmup 5
b reference_arm --impl reference --steps 2 --warmup 1
b c2_f32 --dtype f32 --steps 50 --warmup 5 --no-cpu
b c3_eeq --workload c3 --steps 20 --warmup 5 --eeq --no-cpu
b c2_eeq --steps 30 --warmup 5 --eeq --no-cpu
b c1 --workload c1 --steps 50 --warmup 5 --no-cpu
b c1_eeq --workload c1 --steps 50 --warmup 5 --eeq --no-cpu
ls -la $out | tail -30

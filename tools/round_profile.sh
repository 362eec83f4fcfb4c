#!/bin/bash
# One GPU-box pass that regenerates the measured artefacts kept under profiles/:
# gpu test log, bench lines (C1, C2 FP64/FP32, C3, C5, the q=None/EEQ variants, reference arm),
# ncu launch lists, ncu --set full captures of the dominant kernels.
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > $out/bench_$name.json 2> $out/bench_$name.err; cut -c1-200 $out/bench_$name.json; }
b c2 --steps 50 --warmup 5
b c3 --workload c3 --steps 20 --warmup 5
b c2_f32 --dtype f32 --steps 50 --warmup 5 --no-cpu
b c2_eeq --steps 30 --warmup 5 --eeq
b c3_eeq --workload c3 --steps 20 --warmup 5 --eeq
b c5_f64 --workload c5 --steps 20 --warmup 5
b c5_f32 --workload c5 --dtype f32 --steps 20 --warmup 5
b c1 --workload c1 --steps 50 --warmup 5
b c1_eeq --workload c1 --steps 50 --warmup 5 --eeq
b reference_arm --impl reference --steps 2 --warmup 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3_eeq.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --eeq > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.0, .bool.0, .int.64" -s 2 -c 1 -f -o $out/prof_c2_e64 python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.1, .bool.0, .int.100" -s 2 -c 1 -f -o $out/prof_c3_g100 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:eeq_kernel -s 3 -c 1 -f -o $out/prof_eeq_c2 python bench.py --workload c2 --steps 1 --warmup 3 --eeq --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:eeq_kernel -s 6 -c 2 -f -o $out/prof_eeq_c3 python bench.py --workload c3 --steps 1 --warmup 3 --eeq --no-cpu > /dev/null 2>&1
ls -la $out | tail -40

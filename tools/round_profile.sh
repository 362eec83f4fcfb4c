#!/bin/bash
# One GPU-box pass that regenerates the measured artefacts kept under profiles/:
# bench lines (C2 FP64/FP32, C3, reference arm), ncu launch lists, ncu --set full captures.
set -x
out=gpurun_out
python bench.py --steps 50 --warmup 5 > $out/bench_c2.json 2> $out/bench_c2.err
python bench.py --workload c3 --steps 20 --warmup 5 > $out/bench_c3.json 2> $out/bench_c3.err
python bench.py --dtype f32 --steps 50 --warmup 5 --no-cpu > $out/bench_c2_f32.json 2> /dev/null
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.0, .bool.0, .int.64" -s 2 -c 1 -f -o $out/prof_c2_e64 python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.1, .bool.0, .int.100" -s 2 -c 1 -f -o $out/prof_c3_g100 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la $out

#!/bin/bash
# One 1-GPU box pass that regenerates the measured artefacts kept under profiles/: gpu test log, smoke,
# same-box A/B of library variants built with tools/build_variant.sh (build_ab/*.so), bench lines, ncu launch
# lists and --set full captures of the dominant kernels, compute-sanitizer memcheck.
#   tools/gpu_retry.sh 2400 'bash tools/round_profile.sh'      (retries while the pod is busy)
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/r02_smoke.log 2>&1; tail -2 $out/r02_smoke.log
rm -f $out/r02_ab_r01_final.txt
for w in "c3" "c2" "c5" "c5 --dtype f32" "c3 --dtype f32" "c2 --dtype f32"; do
  bash tools/ab.sh "--workload $w --steps 30 --warmup 5 --no-subs --no-e2e" build_ab/r01.so build_ab/final.so 2>&1 | tee -a $out/r02_ab_r01_final.txt
done
b() { name=$1; shift; timeout 600 python bench.py "$@" > $out/r02_bench_$name.json 2> $out/r02_bench_$name.err; cut -c1-200 $out/r02_bench_$name.json; }
b default --steps 20 --warmup 3
b reference --impl reference --steps 3 --warmup 1
b c3_f32 --workload c3 --dtype f32 --steps 20 --warmup 5 --no-subs
b c3_eeq --workload c3 --eeq --steps 20 --warmup 5 --no-subs --no-cpu
b c2_eeq --workload c2 --eeq --steps 30 --warmup 5 --no-subs --no-cpu
b c1 --workload c1 --steps 50 --warmup 5 --no-subs --no-cpu
b c4g --workload c4g --steps 3 --warmup 1 --no-cpu
python tools/eeq_time.py 2>/dev/null | tail -3 > $out/r02_eeq_time_factor.txt; cat $out/r02_eeq_time_factor.txt
N="--kernel-name-base demangled --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu $N -k "regex:small_kernel<double, .bool.1, .bool.0, .int.100" -s 2 -c 1 -f -o $out/prof_c3_g100 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-subs > /dev/null 2>&1
ncu $N -k "regex:small_kernel<double, .bool.0, .bool.0, .int.64" -s 2 -c 1 -f -o $out/prof_c2_e64 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-subs > /dev/null 2>&1
ncu $N -k "regex:small_kernel<float, .bool.1, .bool.0, .int.128" -s 2 -c 1 -f -o $out/prof_c3_g128_f32 python bench.py --workload c3 --dtype f32 --steps 1 --warmup 3 --no-cpu --no-subs > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/r02_launches_c4g.csv python bench.py --workload c4g --steps 1 --warmup 1 > /dev/null 2>&1
ncu $N -k "regex:large_atm_grad" -c 1 -f -o $out/prof_c4_atm_grad python bench.py --workload c4g --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
ncu $N -k "regex:eeq_kernel<double, .bool.1" -s 2 -c 1 -f -o $out/prof_eeq_vjp_c3 python bench.py --workload c3 --eeq --steps 1 --warmup 3 --no-cpu --no-subs --no-e2e > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > $out/r02_sanitizer.txt 2>&1; echo "sanitizer rc=$?"; tail -3 $out/r02_sanitizer.txt

#!/bin/bash
# validation of the D4S gradient rework (in-tree = final3), same-box A/B, exit-path diagnosis
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r02_pytest_gpu.log
for w in "c5" "c5 --dtype f32"; do
  bash tools/ab.sh "--workload $w --steps 30 --warmup 5 --no-subs --no-e2e" build_ab/final2.so build_ab/final3.so 2>&1 | tee -a $out/r02_ab_d4s.txt
done
for i in 1 2 3 4 5 6; do
  D4_BENCH_EXIT=normal timeout 300 python bench.py --steps 5 > /dev/null 2> $out/r02_exit_normal_$i.err; echo "normal exit run $i rc=$?"
done
for i in 1 2 3; do
  timeout 300 python bench.py --steps 5 > $out/r02_exit_hooks_$i.json 2> $out/r02_exit_hooks_$i.err; echo "explicit-hooks exit run $i rc=$? lines=$(wc -l < $out/r02_exit_hooks_$i.json)"
done
D4_BENCH_EXIT=normal timeout 300 python bench.py --steps 5 --no-cpu > /dev/null 2> $out/r02_exit_normal_nocpu.err; echo "normal exit --no-cpu rc=$?"
D4_BENCH_EXIT=normal timeout 300 python bench.py --steps 5 --no-subs > /dev/null 2> $out/r02_exit_normal_nosubs.err; echo "normal exit --no-subs rc=$?"
tail -n 30 $out/r02_exit_normal_*.err | grep -v "^$" | head -80
b() { name=$1; shift; timeout 600 python bench.py "$@" > $out/r02_bench_$name.json 2> $out/r02_bench_$name.err; cut -c1-120 $out/r02_bench_$name.json; }
b default --steps 20 --warmup 3
N="--kernel-name-base demangled --set full --clock-control none --import-source on"
ncu $N -k "regex:small_kernel<double, .bool.1, .bool.1, .int.120" -s 2 -c 1 -f -o $out/prof_c5_g120_d4s python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu --no-subs > /dev/null 2>&1

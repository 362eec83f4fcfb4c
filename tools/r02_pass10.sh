#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r02_pytest_gpu.log
for w in "c3" "c3 --dtype f32" "c5"; do
  bash tools/ab.sh "--workload $w --steps 30 --warmup 5 --no-subs --no-e2e" build_ab/final3.so build_ab/final4.so 2>&1 | tee -a $out/r02_ab_weights4.txt
done

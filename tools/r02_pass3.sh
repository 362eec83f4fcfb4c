#!/bin/bash
# round 2, GPU pass 3: FP64 issue-rate probe; parity of the current in-tree-equivalent variant; A/B of sweep variants
out=gpurun_out
mkdir -p $out
./build_ab/fp64_probe | tee $out/r02_fp64_probe.txt
D4B200_LIBRARY=build_ab/t512.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $out/r02_pytest_t512.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02_pytest_t512.log
D4B200_LIBRARY=build_ab/t384.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/r02_pytest_t384.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02_pytest_t384.log
bash tools/ab.sh "--workload c3 --steps 30 --warmup 5 --no-subs" build_ab/base0.so build_ab/untiled.so build_ab/tiled0.so build_ab/t512.so build_ab/t512u2.so build_ab/t384.so build_ab/t384u2.so 2>&1 | tee $out/r02_ab2_c3.txt

#!/bin/bash
out=gpurun_out
mkdir -p $out
./build_ab/fp64_probe | tee $out/r02_fp64_probe.txt
timeout 1200 python -m pytest tests/test_gpu_large.py tests/test_gpu_install.py -m gpu -x -q > $out/r02_pytest_large.log 2>&1; echo "pytest rc=$?"; tail -15 $out/r02_pytest_large.log

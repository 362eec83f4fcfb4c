#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_hessian.py tests/test_gpu_eeq.py tests/test_gpu_parity.py -m gpu -x -q > $out/r02_pytest_hessian.log 2>&1; echo "pytest rc=$?"; tail -30 $out/r02_pytest_hessian.log

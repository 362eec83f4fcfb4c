#!/bin/bash
# A/B pass: parity tests with the in-tree library, then timing of library variants (tools/ab.sh)
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_ab.log 2>&1; tail -3 $out/pytest_ab.log
bash tools/ab.sh "${AB_ARGS:---steps 40 --warmup 5}" build_ab/base.so tad_dftd4_b200/libd4b200.so "$@" 2>&1 | tee $out/ab.txt

#!/bin/bash
# ncu --set full captures of the EEQ kernels
out=gpurun_out
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eeq_kernel -s 3 -c 1 -f -o $out/prof_eeq_c2 python bench.py --workload c2 --steps 1 --warmup 3 --eeq --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eeq_kernel -s 6 -c 2 -f -o $out/prof_eeq_c3 python bench.py --workload c3 --steps 1 --warmup 3 --eeq --no-cpu > /dev/null 2>&1
ls -la $out/*.ncu-rep

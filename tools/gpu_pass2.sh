#!/bin/bash
# EEQ kernel timings (ncu launch list) + bench lines of the q=None variants.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_eeq.py -x -q > $out/pytest_eeq.log 2>&1; tail -3 $out/pytest_eeq.log
for w in c2 c3 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --eeq --no-cpu > $out/bench_${w}_eeq.json 2> $out/bench_${w}_eeq.err
  cut -c1-330 $out/bench_${w}_eeq.json
done
for w in c2 c3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:eeq_kernel -c 12 --csv --log-file $out/launches_eeq_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --eeq --no-cpu > /dev/null 2>&1
  grep eeq_kernel $out/launches_eeq_$w.csv | awk -F'","' '{print $5, $NF}' | tail -6
done

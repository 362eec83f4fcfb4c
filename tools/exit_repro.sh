#!/bin/bash
# 20 runs of each mode of tools/exit_repro.py; prints the exit-code histogram per mode.
out=gpurun_out/r02_exit_repro.txt
: > $out
for mode in race idle; do
  for i in $(seq 15); do
    python tools/exit_repro.py $mode > /dev/null 2> /tmp/e.txt; rc=$?
    echo "$mode run $i rc=$rc" >> $out
  done
done
for mode in race idle; do echo "$mode: $(grep "^$mode" $out | grep -c 'rc=0') ok, $(grep "^$mode" $out | grep -vc 'rc=0') failed" | tee -a $out; done

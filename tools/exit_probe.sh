#!/bin/bash
# Which part of bench.py makes a PLAIN interpreter exit abort now and then?  Runs each variant several times with
# D4_BENCH_EXIT=normal and the SIGABRT stack tracer, and records exit codes + the stack of any abort.
out=gpurun_out/${EXIT_PROBE_OUT:-r02_exit_probe.txt}
: > $out
export D4_BENCH_EXIT=normal D4_BENCH_ABORT_TRACE=$PWD/build_ab/abort_trace.so
run() { # label, repeats, args...
    label=$1; reps=$2; shift 2
    for i in $(seq $reps); do
        timeout 300 python bench.py --steps 5 --warmup 3 "$@" > /tmp/o.txt 2> /tmp/e.txt
        rc=$?
        echo "$label run $i rc=$rc json=$(grep -c '"metric"' /tmp/o.txt)" >> $out
        if [ $rc -ne 0 ]; then grep -A60 "abort_trace" /tmp/e.txt | head -70 >> $out; tail -5 /tmp/e.txt >> $out; fi
    done
}
run default 3
run with-cpu 5 --no-subs --no-e2e
cat $out | grep -c "rc=0" ; grep "rc=" $out | grep -v "rc=0"

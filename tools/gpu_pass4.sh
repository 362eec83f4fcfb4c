#!/bin/bash
out=gpurun_out
mkdir -p $out
python tools/phase_probe.py c2 > $out/phase_c2.txt 2>&1; cat $out/phase_c2.txt
python tools/phase_probe.py c3 > $out/phase_c3.txt 2>&1; cat $out/phase_c3.txt

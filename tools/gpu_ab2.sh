#!/bin/bash
# timing of library variants only (tools/ab.sh): gpu_ab2.sh "<bench args>" lib...
args="$1"; shift
bash tools/ab.sh "$args" "$@" 2>&1 | tee gpurun_out/ab.txt

#!/bin/bash
# A/B timing of library builds inside ONE gpurun call (box-to-box variation is ~10 %):
#   tools/ab.sh "<bench args>" build_ab/base.so build_ab/v1.so ...
args="$1"; shift
for r in 1 2; do
  for lib in "$@"; do
    D4B200_LIBRARY=$lib python bench.py $args --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'step %.4f ms' % d['ms_per_step'], 'e2e %.4f' % (d['e2e']['ms_per_step'] if d['e2e'] else 0), 'classes', [round(x,4) for x in d['roofline']['all_class_ms']])"
  done
done

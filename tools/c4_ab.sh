# same-box A/B of the C4 energy+forces step (+ the large-system tests per library):
#   [C4_AB_OUT=name.txt] tools/c4_ab.sh lib1.so lib2.so ...
for lib in "$@"; do D4B200_LIBRARY=$PWD/$lib python -m pytest tests/test_gpu_large.py -m gpu -q -x 2>&1 | tail -1; done
for r in 1 2; do for lib in "$@"; do
  D4B200_LIBRARY=$PWD/$lib python bench.py --workload c4g --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'step %.1f ms' % d['ms_per_step'])"
done; done | tee gpurun_out/${C4_AB_OUT:-r02_ab_c4.txt}

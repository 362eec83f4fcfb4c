# same-box A/B of the C4 energy+forces step: tools/c4_ab.sh lib1.so lib2.so ...
for r in 1 2; do for lib in "$@"; do
  D4B200_LIBRARY=$PWD/$lib python bench.py --workload c4g --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'step %.1f ms' % d['ms_per_step'], d.get('sections', d.get('config', {})).get('sections', ''))"
done; done | tee gpurun_out/${C4_AB_OUT:-r02_ab_c4.txt}
python -m pytest tests/test_gpu_large.py -m gpu -q -x 2>&1 | tail -2

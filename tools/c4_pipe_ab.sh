# large_atm_grad (0) against large_atm_grad_pipe (1) on one box: the large-system tests with the pipeline, then timing
D4B200_LARGE_PIPE=1 timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_param.py -m gpu -q -x 2>&1 | tail -3
for r in 1 2; do for v in 0 1; do
  D4B200_LARGE_PIPE=$v timeout 300 python bench.py --workload c4g --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe=$v', 'step %.1f ms' % d['ms_per_step'], 'E %.12f' % d['energy_sum'])"
done; done | tee gpurun_out/r02_ab_c4_pipe.txt

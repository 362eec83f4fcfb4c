python -m pytest tests/test_gpu_eeq.py tests/test_gpu_hessian.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_pytest_eeq_factor.log
python tools/eeq_time.py 2>&1 | tail -3 | tee gpurun_out/r02_eeq_time_factor.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_eeq.py > gpurun_out/r02_sanitizer_eeq.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_eeq.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_eeq.py >> gpurun_out/r02_sanitizer_eeq.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_eeq.txt
for w in c3 c2; do for lib in build_ab/pre_eeq.so tad_dftd4_b200/libd4b200.so; do
  D4B200_LIBRARY=$PWD/$lib python bench.py --workload $w --eeq --steps 20 --warmup 5 --no-subs --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $lib', 'step %.4f ms' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'])"
done; done | tee gpurun_out/r02_ab_eeq_factor.txt

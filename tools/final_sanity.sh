python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final_default.json 2> gpurun_out/r02_bench_final_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_final_default.json
python bench.py --impl reference > gpurun_out/r02_bench_final_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02_bench_final_reference.json
for i in 1 2 3; do python bench.py --workload c1 --steps 50 --warmup 5 --no-subs --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c1', d['ms_per_step'], d['e2e']['value'])"; done

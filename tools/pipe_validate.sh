python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r02_racecheck_all.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_racecheck_all.txt; grep -E "RACECHECK SUMMARY|^large" gpurun_out/r02_racecheck_all.txt
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r02_sanitizer.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer.txt; grep -E "ERROR SUMMARY" gpurun_out/r02_sanitizer.txt
python bench.py --workload c4g --steps 3 --warmup 1 --no-cpu > gpurun_out/r02_bench_c4g.json 2>/dev/null; cut -c1-120 gpurun_out/r02_bench_c4g.json
ncu --kernel-name-base demangled --set full --clock-control none --import-source on -k "regex:large_atm_grad_pipe" -c 1 -f -o gpurun_out/prof_c4_atm_grad_pipe python bench.py --workload c4g --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1; ls -la gpurun_out/prof_c4_atm_grad_pipe.ncu-rep

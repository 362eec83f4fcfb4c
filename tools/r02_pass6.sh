#!/bin/bash
out=gpurun_out
mkdir -p $out
D4B200_LIBRARY=build_ab/cur2.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_model.py tests/test_gpu_eeq.py tests/test_gpu_hessian.py tests/test_gpu_param.py -m gpu -x -q > $out/r02_pytest_cur.log 2>&1; echo "pytest rc=$?"; tail -8 $out/r02_pytest_cur.log
D4B200_LIBRARY=build_ab/e_lowocc.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/r02_pytest_elowocc.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02_pytest_elowocc.log
bash tools/ab.sh "--workload c2 --steps 40 --warmup 5 --no-subs" build_ab/e_untiled.so build_ab/nostage.so build_ab/cur.so build_ab/e_lowocc.so 2>&1 | tee $out/r02_ab_c2.txt
bash tools/ab.sh "--workload c3 --steps 30 --warmup 5 --no-subs" build_ab/nostage.so build_ab/cur.so build_ab/cur2.so 2>&1 | tee $out/r02_ab_c3b.txt
D4B200_LIBRARY=build_ab/cur2.so timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > $out/r02_sanitizer.txt 2>&1; echo "sanitizer rc=$?"; tail -4 $out/r02_sanitizer.txt

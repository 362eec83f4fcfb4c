#!/bin/bash
# round 2, GPU pass 2: parity of the register-tiled gradient sweep (build_ab/tiled.so), A/B timing against the
# untiled build, ncu --set full (with source) of the C2 class-64 energy kernel and the tiled C3 kernel
out=gpurun_out
mkdir -p $out
D4B200_LIBRARY=build_ab/tiled.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_install.py tests/test_gpu_param.py -m gpu -x -q > $out/r02_pytest_tiled.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02_pytest_tiled.log
bash tools/ab.sh "--workload c3 --steps 30 --warmup 5 --no-subs" build_ab/base.so build_ab/tiled.so 2>&1 | tee $out/r02_ab_c3.txt
bash tools/ab.sh "--workload c5 --steps 20 --warmup 5 --no-subs" build_ab/base.so build_ab/tiled.so 2>&1 | tee $out/r02_ab_c5.txt
bash tools/ab.sh "--workload c5 --dtype f32 --steps 20 --warmup 5 --no-subs" build_ab/base.so build_ab/tiled.so 2>&1 | tee $out/r02_ab_c5f32.txt
D4B200_LIBRARY=build_ab/tiled.so ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.1, .bool.0, .int.100" -s 2 -c 1 -f -o $out/prof_c3_g100_tiled python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-subs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:small_kernel<double, .bool.0, .bool.0, .int.64" -s 2 -c 1 -f -o $out/prof_c2_e64 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-subs > /dev/null 2>&1
ls -la $out

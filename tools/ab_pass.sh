# same-box A/B of build_ab/final.so against the in-tree library on the gradient workloads + the GPU test suite
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for w in "c3" "c5" "c5 --dtype f32" "c3 --dtype f32"; do
  bash tools/ab.sh "--workload $w --steps 30 --warmup 5 --no-subs --no-e2e" build_ab/final.so tad_dftd4_b200/libd4b200.so 2>&1
done | tee gpurun_out/${AB_OUT:-r02_ab_pass.txt}

#!/bin/bash
# Multi-GPU pass (gpurun --gpus N -- 'bash tools/round_multi.sh N'): bench under torchrun (the driver's launch
# line), reference arm under torchrun, cross-rank consistency of the large-system path
N=${1:-2}
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
( time run 29501 bench.py --gpus $N --steps 20 --warmup 3 ) > $out/r02_bench_n$N.json 2> $out/r02_bench_n$N.err; echo "bench rc=$?"; tail -c 600 $out/r02_bench_n$N.err
[ -n "$SKIP_REF" ] || run 29502 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $out/r02_bench_ref_n$N.json 2> $out/r02_bench_ref_n$N.err; echo "ref rc=$?"; wc -l $out/r02_bench_ref_n$N.json
run 29503 tools/c4_crossrank.py 6667 > $out/r02_crossrank_n$N.json 2> $out/r02_crossrank_n$N.err; echo "crossrank rc=$?"; cat $out/r02_crossrank_n$N.json
python -m pytest tests/test_gpu_large.py -m gpu -x -q -k cross_rank 2>&1 | tail -3
nvidia-smi topo -m > $out/r02_topo_n$N.txt 2>&1

#!/bin/bash
# multi-GPU pass (gpurun --gpus N): strong scaling of the 20 001-atom system (C4 energy, C4 energy+gradient)
N=${1:-4}
out=gpurun_out
mkdir -p $out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > $out/bench_${name}_n$N.json 2> $out/bench_${name}_n$N.err; cut -c1-300 $out/bench_${name}_n$N.json; tail -2 $out/bench_${name}_n$N.err; }
run c4 --workload c4 --steps 3 --warmup 3 --no-cpu
run c4g --workload c4g --steps 2 --warmup 3 --no-cpu

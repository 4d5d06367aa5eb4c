#!/bin/bash
# Multi-GPU visit: bench.py under torchrun at N GPUs (default line + embedded sharded configs). Usage: bash tools/gpu_multi.sh <tag> <N>
tag=${1:-multi}; n=${2:-2}
out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
echo "rc=$?"; tail -5 $out/bench_n$n.err; cut -c1-600 $out/bench_n$n.json

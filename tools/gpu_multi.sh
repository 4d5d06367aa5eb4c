#!/bin/bash
# Multi-GPU visit: bench.py under torchrun at N GPUs (default line + embedded sharded configs), then the strip-sharded config
# alone with a sweep over the frames in flight. Usage: bash tools/gpu_multi.sh <tag> <N> [sweep]
tag=${1:-multi}; n=${2:-2}; sweep=${3:-}
out=gpurun_out/$tag; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
timeout 1500 bash -c "$(declare -f run); n=$n; run --steps 20 --warmup 3" > $out/bench_n$n.json 2> $out/bench_n$n.err
echo "rc=$?"; tail -3 $out/bench_n$n.err; cut -c1-300 $out/bench_n$n.json
if [ -n "$sweep" ]; then
  timeout 900 bash -c "$(declare -f run); n=$n; run --workload synthetic --gather copy --steps 20 --warmup 3 --strip-frames-in-flight 4 --strip-sweep $sweep" > $out/strips_n$n.json 2> $out/strips_n$n.err
  echo "rc=$?"; tail -3 $out/strips_n$n.err; cut -c1-1200 $out/strips_n$n.json
fi

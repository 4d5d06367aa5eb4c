#!/bin/bash
# strip-sharded config alone at N GPUs with a sweep over frames in flight. Usage: bash tools/gpu_strips.sh <tag> <N> <gather> <sweep>
tag=${1:-strips}; n=${2:-2}; gather=${3:-copy}; sweep=${4:-1,3}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload synthetic --gather $gather --steps 20 --warmup 3 --strip-frames-in-flight 4 --strip-sweep $sweep > $out/strips_${gather}_n$n.json 2> $out/strips_${gather}_n$n.err
echo "rc=$?"; tail -3 $out/strips_${gather}_n$n.err; cut -c1-1500 $out/strips_${gather}_n$n.json

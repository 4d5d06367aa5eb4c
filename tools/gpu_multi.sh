#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the default bench (scene per rank), the strip-sharded synthetic canvas with both
# gathers, and the scene-sharded batch. Usage: bash tools/gpu_multi.sh <tag> <N>
tag=${1:-multi}; N=${2:-2}
out=gpurun_out/$tag; mkdir -p $out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 100 --warmup 10 > $out/bench_n$N.json 2> $out/bench_n$N.err; tail -c 600 $out/bench_n$N.json; echo
run --workload synthetic --paths 200000 --size 8192 --steps 10 --warmup 3 --gather nccl > $out/synth_nccl_n$N.json 2> $out/synth_nccl_n$N.err; tail -c 900 $out/synth_nccl_n$N.json; echo
run --workload synthetic --paths 200000 --size 8192 --steps 10 --warmup 3 --gather p2p > $out/synth_p2p_n$N.json 2> $out/synth_p2p_n$N.err; tail -c 900 $out/synth_p2p_n$N.json; echo
run --workload tiger512 --frames 4096 --steps 2 --warmup 1 > $out/batch_n$N.json 2> $out/batch_n$N.err; tail -c 700 $out/batch_n$N.json; echo
for f in $out/*.err; do echo "== $f"; tail -n 3 $f; done

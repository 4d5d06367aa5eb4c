#!/bin/bash
# Multi-GPU visit (gpurun --gpus N), the two sharded BASELINE configs only: strip-sharded synthetic canvas (peer-store
# gather) and the scene-sharded batch. Usage: bash tools/gpu_multi2.sh <tag> <N>
tag=${1:-multi}; N=${2:-4}
out=gpurun_out/$tag; mkdir -p $out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
date +%s > $out/t0
run --workload synthetic --paths 200000 --size 8192 --steps 10 --warmup 3 --gather p2p > $out/synth_p2p_n$N.json 2> $out/synth_p2p_n$N.err; tail -c 900 $out/synth_p2p_n$N.json; echo
run --workload tiger512 --frames 4096 --steps 2 --warmup 1 > $out/batch_n$N.json 2> $out/batch_n$N.err; tail -c 700 $out/batch_n$N.json; echo
date +%s > $out/t1
for f in $out/*.err; do echo "== $f"; tail -n 3 $f; done

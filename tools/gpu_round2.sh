#!/bin/bash
# GPU-box visit: parity tests, bench line, stage times, other BASELINE configs (features 2048, synthetic 8192, batch 512).
# Usage (under gpurun): bash tools/gpu_round2.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 3000 $out/bench.json
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1; cat $out/stage_times.txt
timeout 300 python bench.py --workload features2048 --no-cpu-baseline > $out/bench_features2048.json 2> $out/bench_features.err; tail -c 1500 $out/bench_features2048.json
timeout 600 python bench.py --workload synthetic --paths 200000 --size 8192 --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_synth.json 2> $out/bench_synth.err; tail -c 2500 $out/bench_synth.json; tail -3 $out/bench_synth.err
timeout 600 python bench.py --workload tiger512 --frames 4096 --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_batch.json 2> $out/bench_batch.err; tail -c 1500 $out/bench_batch.json; tail -3 $out/bench_batch.err
ls -la $out

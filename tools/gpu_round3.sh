#!/bin/bash
# GPU-box visit: parity tests, bench line (streamed e2e), e2e vs frames in flight, features / demo configs.
# Usage (under gpurun): bash tools/gpu_round3.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -8 $out/pytest.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 3000 $out/bench.json; tail -3 $out/bench.err
timeout 300 python tools/e2e_stream.py tiger4096 > $out/e2e_stream.txt 2>&1; cat $out/e2e_stream.txt
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1; cat $out/stage_times.txt
timeout 300 python bench.py --workload features2048 --no-cpu-baseline > $out/bench_features2048.json 2> $out/bench_features.err; tail -c 1200 $out/bench_features2048.json; tail -3 $out/bench_features.err
timeout 300 python bench.py --workload demo2048 --no-cpu-baseline > $out/bench_demo2048.json 2> $out/bench_demo.err; tail -c 1200 $out/bench_demo2048.json; tail -3 $out/bench_demo.err
ls -la $out

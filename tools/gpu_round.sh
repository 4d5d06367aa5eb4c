#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-stage times, ncu launch list, ncu --set full of the raster kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 3000 $out/bench.json
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1; cat $out/stage_times.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $out/launches.csv python tools/prof_frame.py --split --frames 8 > $out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fill|k_composite' -s 6 -c 2 -o $out/raster python tools/prof_frame.py --split --frames 6 > $out/ncu_raster.log 2>&1
ls -la $out

#!/bin/bash
# round 2, visit a: baseline of the round-1 build + store floors
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 120 tools/ubench/store_floor > $out/store_floor.txt 2>&1; cat $out/store_floor.txt
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1; cat $out/stage_times.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -3 $out/pytest.log

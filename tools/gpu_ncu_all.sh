#!/bin/bash
# ncu --set full of every kernel of one warm frame (split fill + tile), source counters included.
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 ncu --set full --clock-control none --import-source on -s 44 -c 11 -o $out/frame python tools/prof_frame.py --split --frames 6 > $out/ncu_frame.log 2>&1
tail -3 $out/ncu_frame.log
ls -la $out

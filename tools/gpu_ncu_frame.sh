#!/bin/bash
# ncu --set full of EVERY kernel of one warm frame (11 launches) + launch list of 3 frames. Usage: bash tools/gpu_ncu_frame.sh <tag> [fixture]
tag=${1:-ncu_frame}; fixture=${2:-tiger_4096_scene}
out=gpurun_out/$tag; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 55 -c 33 --csv --log-file $out/launches.csv python tools/prof_frame.py --fixture $fixture --frames 8 > $out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 44 -c 11 -o $out/frame python tools/prof_frame.py --fixture $fixture --frames 6 > $out/ncu_frame.log 2>&1
tail -2 $out/ncu_frame.log; ls -la $out

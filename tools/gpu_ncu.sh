#!/bin/bash
# ncu --set full of the raster kernels of one warm frame + launch list. Usage (under gpurun): bash tools/gpu_ncu.sh <tag> [fixture] [kernel regex]
tag=${1:-ncu}; fixture=${2:-tiger_4096_scene}; rx=${3:-^(k_fill|k_composite)$}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 55 -c 33 --csv --log-file $out/launches.csv python tools/prof_frame.py --fixture $fixture --frames 8 > $out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 6 -c 2 -o $out/raster python tools/prof_frame.py --fixture $fixture --frames 6 > $out/ncu_raster.log 2>&1
tail -3 $out/ncu_raster.log
ls -la $out

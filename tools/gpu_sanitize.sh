#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over whole frames of the scenes that reach every kernel path: tiger 512
# (plain colours), the demo's primitives scene (clips, render targets, blur, image) and features 2048 (gradients, even-odd).
# Usage (under gpurun): bash tools/gpu_sanitize.sh <tag>
tag=${1:-sanitize}; out=gpurun_out/$tag; mkdir -p $out
for tool in memcheck racecheck synccheck; do
  for fx in tiger_512 demo_full_512 features_2048; do
    [ $tool != memcheck ] && [ $fx = features_2048 ] && continue
    echo "== compute-sanitizer --tool $tool  prof_frame.py --fixture $fx --frames 2" >> $out/sanitizer.txt
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_frame.py --fixture $fx --frames 2 2>&1 | grep -v "^[0-9] {" | tail -8 >> $out/sanitizer.txt
  done
  # the fill stage inside the tile kernel (PFCU_OPT_FUSED_FILL): shared-memory accumulators, masks and work queues
  for fx in tiger_512 demo_full_512; do
    echo "== compute-sanitizer --tool $tool  prof_frame.py --fixture $fx --frames 2 --fused" >> $out/sanitizer.txt
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_frame.py --fixture $fx --frames 2 --fused 2>&1 | grep -v "^[0-9] {" | tail -8 >> $out/sanitizer.txt
  done
done
cat $out/sanitizer.txt

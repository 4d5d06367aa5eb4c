#!/bin/bash
# scratch experiments on the GPU box
out=gpurun_out/exp; mkdir -p $out
for d in 1 2 3; do PFCU_EXP_DRAWS=$d python tools/stage_times.py 2>&1 | sed 's/init.*fill /fill /' | sed "s/^/draws=$d /"; done | tee $out/draws.txt

#!/bin/bash
# GPU-box visit for kernel A/B: parity tests on the default library, then stage times for every variant given.
# Usage (under gpurun): bash tools/gpu_ab.sh <tag> [variant ...]
tag=${1:-ab}; shift
out=gpurun_out/$tag
mkdir -p $out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -15 $out/pytest.log
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1
for v in "$@"; do
  PFCU_LIB=$PWD/pathfinder-cpp_b200/lib/libpfcu_$v.so timeout 300 python tools/stage_times.py >> $out/stage_times.txt 2>&1
done
cat $out/stage_times.txt

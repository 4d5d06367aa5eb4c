#!/bin/bash
# Quick GPU visit: parity tests (optionally filtered with -k "$PYTEST_K"), then stage times of the default library and of every
# variant named on the command line (tools/variants.sh). Usage (under gpurun): bash tools/gpu_quick.sh <tag> [variant ...]
tag=${1:-quick}; shift
out=gpurun_out/$tag; mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
  tail -25 $out/pytest.log
fi
for f in ${FIXTURES:-tiger_4096_scene}; do
  timeout 300 python tools/stage_times.py $f >> $out/stage_times.txt 2>&1
  for v in "$@"; do
    PFCU_LIB=$PWD/pathfinder-cpp_b200/lib/libpfcu_$v.so timeout 300 python tools/stage_times.py $f >> $out/stage_times.txt 2>&1
  done
done
cat $out/stage_times.txt

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
if [ -n "$SYNTH" ]; then
  for v in "" "$@"; do
    lib=$PWD/pathfinder-cpp_b200/lib/libpfcu${v:+_$v}.so
    PFCU_LIB=$lib timeout 300 python bench.py --workload synthetic --paths 200000 --size 8192 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('synthetic ${v:-default}', round(d['ms_per_step'],3), {k:round(x*1e3) for k,x in d['config']['rank0_stage_ms'].items()})"
  done | tee $out/synth.txt
fi

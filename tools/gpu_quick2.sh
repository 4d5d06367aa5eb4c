#!/bin/bash
# tests + stage times (tiger 4096) + the synthetic config-4 line on one GPU. Usage: bash tools/gpu_quick2.sh <tag>
tag=${1:-q2}; out=gpurun_out/$tag; mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -6 $out/pytest.log
fi
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1; cat $out/stage_times.txt
timeout 600 python bench.py --workload synthetic --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('synthetic', round(d['ms_per_step'],3), {k:round(x*1e3) for k,x in d['config']['rank0_stage_ms'].items()})" | tee $out/synth.txt

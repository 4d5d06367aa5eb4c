#!/bin/bash
# Final GPU visit of a round: parity tests, every bench line, ncu launch list of the bench command, ncu --set full of
# every kernel of one frame, stroke / incremental-frame timings, compute-sanitizer. Usage (under gpurun): bash tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; cut -c1-400 $out/bench.json; tail -2 $out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_reference.json 2>&1
timeout 300 python tools/stage_times.py > $out/stage_times.txt 2>&1; cat $out/stage_times.txt
timeout 300 python tools/gpu_throughput.py tiger4096 > $out/gpu_throughput.txt 2>&1; cat $out/gpu_throughput.txt
timeout 300 python tools/timeline.py > $out/timeline.txt 2>&1; head -16 $out/timeline.txt
timeout 300 python bench.py --workload features2048 --no-cpu-baseline --no-sharded --steps 20 > $out/bench_features2048.json 2> $out/bench_features.err; cut -c1-200 $out/bench_features2048.json
timeout 300 python bench.py --workload demo2048 --no-cpu-baseline --no-sharded --steps 4 --frames-per-step 64 > $out/bench_demo2048.json 2> $out/bench_demo.err; cut -c1-200 $out/bench_demo2048.json
timeout 300 python tools/stroke_time.py > $out/stroke_time.txt 2>&1; cat $out/stroke_time.txt
timeout 600 python tools/incremental_cost.py 50000 4096 > $out/incremental_cost.txt 2>&1; cat $out/incremental_cost.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_bench.csv python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline --no-sharded > $out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 44 -c 11 -o $out/frame python tools/prof_frame.py --fixture tiger_4096_scene --frames 6 > $out/ncu_frame.log 2>&1; tail -2 $out/ncu_frame.log
bash tools/gpu_sanitize.sh $tag > /dev/null 2>&1; tail -30 $out/sanitizer.txt | grep -E "==|SUMMARY"
ls -la $out

for v in "" fp0 all_bp0 all_both0; do
  lib=$PWD/pathfinder-cpp_b200/lib/libpfcu${v:+_$v}.so
  for c in "" 1; do
  PFCU_EXP_FILL_CULLED=$c PFCU_LIB=$lib timeout 300 python bench.py --workload tiger512 --frames 4096 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch ${v:-default} culled=$c', round(d['value']))"
  done
done

#!/bin/bash
# A/B of the issue-loop changes (one barrier wait per slot in pair mode; lane-0 polling)
cd /root/repo
V=/root/repo/nerf_mlp_b200/csrc/variants
OUT=gpurun_out/r67_ab.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r67_tests.txt 2>&1; tail -n 4 gpurun_out/r67_tests.txt
for lib in "" $V/libnerf_b200_base.so $V/libnerf_b200_relayonly.so; do
  export NERF_B200_LIB=$lib; [ -z "$lib" ] && unset NERF_B200_LIB
  timeout 120 python tests/tc_bench.py 16384 192 9 0 >> $OUT 2>&1
  timeout 120 python tests/tc_bench.py 1024 192 15 1 >> $OUT 2>&1
  timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench train ms_per_step', d['ms_per_step'], d['stage_ms'])" >> $OUT 2>&1
done
unset NERF_B200_LIB
timeout 300 python bench.py --workload render --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench render ms_per_step', d['ms_per_step'], d['value'], d['roofline']['frac'], d['density_only_coarse']['value'])" >> $OUT 2>&1
cat $OUT
for mode in "4096 192 1 0" "1024 192 1 1"; do echo "== $mode" >> gpurun_out/r67_trace.txt; NERF_B200_LIB=$V/libnerf_b200_trace.so timeout 120 python tests/tc_bench.py $mode >> gpurun_out/r67_trace.txt 2>&1; done

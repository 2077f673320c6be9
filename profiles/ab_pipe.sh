#!/bin/bash
# pipelined prologue + early next-unit arrives of the inference forward kernel: tests, forward micro-bench, both bench workloads, timeline
cd /root/repo
V=/root/repo/nerf_mlp_b200/csrc/variants
OUT=gpurun_out/r66_ab.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r66_tests.txt 2>&1; tail -n 4 gpurun_out/r66_tests.txt
timeout 120 python tests/tc_bench.py 16384 192 9 0 >> $OUT 2>&1
timeout 120 python tests/tc_bench.py 16384 64 9 0 >> $OUT 2>&1
timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench train ms_per_step', d['ms_per_step'], d['stage_ms'])" >> $OUT 2>&1
timeout 300 python bench.py --workload render --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench render ms_per_step', d['ms_per_step'], d['value'], d['roofline']['frac'], d['density_only_coarse']['value'])" >> $OUT 2>&1
cat $OUT
for mode in "4096 192 1 0"; do echo "== pipe $mode" >> gpurun_out/r66_trace.txt; NERF_B200_LIB=$V/libnerf_b200_trace.so timeout 120 python tests/tc_bench.py $mode >> gpurun_out/r66_trace.txt 2>&1; done

#!/bin/bash
# A/B: weight slots of 4 K-steps (K = 64 per slot, SWIZZLE_128B B operand, ring of 2 x 32 KB) against the default 2 K-steps x ring of 4
cd /root/repo
V=/root/repo/nerf_mlp_b200/csrc/variants
OUT=gpurun_out/r68_ab.txt
: > $OUT
NERF_B200_LIB=$V/libnerf_b200_nk4_r2_s1.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r68_tests_nk4.txt 2>&1; tail -n 4 gpurun_out/r68_tests_nk4.txt
for lib in "" $V/libnerf_b200_nk4_r2_s1.so; do
  export NERF_B200_LIB=$lib; [ -z "$lib" ] && unset NERF_B200_LIB
  timeout 120 python tests/tc_bench.py 16384 192 9 0 >> $OUT 2>&1
  timeout 120 python tests/tc_bench.py 1024 192 15 1 >> $OUT 2>&1
  timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench train ms_per_step', d['ms_per_step'], d['stage_ms'])" >> $OUT 2>&1
done
cat $OUT

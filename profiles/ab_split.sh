#!/bin/bash
# A/B of the per-tile weight streams ("split" mode) of the fused MLP kernels
cd /root/repo
V=/root/repo/nerf_mlp_b200/csrc/variants
OUT=gpurun_out/r63_ab.txt
: > $OUT
for lib in "" $V/libnerf_b200_nosplit.so $V/libnerf_b200_splitinfer.so $V/libnerf_b200_sharedtrain.so; do
  export NERF_B200_LIB=$lib; [ -z "$lib" ] && unset NERF_B200_LIB
  timeout 120 python tests/tc_bench.py 16384 192 9 0 >> $OUT 2>&1
  timeout 120 python tests/tc_bench.py 1024 192 15 1 >> $OUT 2>&1
  timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench train ms_per_step', d['ms_per_step'], d['stage_ms'])" >> $OUT 2>&1
done
unset NERF_B200_LIB
NERF_B200_LIB=$V/libnerf_b200_dbg.so timeout 200 python tests/tc_bench.py 16384 192 2 0 >> $OUT 2>&1
NERF_B200_LIB=$V/libnerf_b200_dbg.so timeout 200 python tests/tc_bench.py 1024 192 2 1 >> $OUT 2>&1
cat $OUT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r63_tests.txt 2>&1; tail -n 8 gpurun_out/r63_tests.txt
NERF_B200_LIB=$V/libnerf_b200_splitinfer.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r63_tests_splitinfer.txt 2>&1; tail -n 8 gpurun_out/r63_tests_splitinfer.txt
NERF_B200_LIB=$V/libnerf_b200_sharedtrain.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r63_tests_sharedtrain.txt 2>&1; tail -n 8 gpurun_out/r63_tests_sharedtrain.txt

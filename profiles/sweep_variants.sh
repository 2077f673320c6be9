#!/bin/bash
# variant sweep of the fused MLP kernels (development; timing only)
cd /root/repo
V=nerf_mlp_b200/csrc/variants
OUT=gpurun_out/r60_sweep.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> $OUT
for lib in nk2_r4_s1 nk2_r4_s3 nk2_r4_s5 nk1_r7_s1 nk1_r7_s3 alias_r5_s1 alias_r5_s3; do
  export NERF_B200_LIB=/root/repo/$V/libnerf_b200_$lib.so
  timeout 120 python tests/tc_bench.py 16384 192 9 0 >> $OUT 2>&1
  case $lib in alias*) ;; *)
    timeout 120 python tests/tc_bench.py 1024 192 15 1 >> $OUT 2>&1
    timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench train ms_per_step', d['ms_per_step'], d['stage_ms'])" >> $OUT 2>&1
  ;; esac
done
cat $OUT

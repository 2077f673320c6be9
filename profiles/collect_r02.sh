#!/bin/bash
# Round-2 evidence collection (final state of the round):  gpurun --timeout 2400 -- bash profiles/collect_r02.sh
# Every ncu pass runs only after the same command exited 0 without ncu (the `&&` directly before it).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
python -m pytest tests -m gpu -q > $O/tests_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
# the driver's command (20 timed steps: burst clocks), the default command (200 steps: power-capped), the reference arm
python bench.py --steps 20 --warmup 5 > $O/bench_train_20steps.json 2> $O/bench_train_20steps.err
python bench.py > $O/bench_train.json 2> $O/bench_train.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
# the two backward paths side by side (stage times inside the replayed step)
NERF_BWD_FUSED=0 python bench.py --steps 20 --warmup 5 --no-render --no-cpu-baseline > $O/bench_train_two_kernel_bwd.json 2>/dev/null
# the other BASELINE configs, one line each
python bench.py --workload render --steps 3 --warmup 3 > $O/bench_render.json 2> $O/bench_render.err
python bench.py --workload render --rays 10000 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_render_100x100_bf16.json 2>/dev/null
python bench.py --workload render --rays 10000 --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_render_100x100_fp32.json 2>/dev/null
python bench.py --rays 4096 --steps 20 --warmup 5 --no-render --no-cpu-baseline > $O/bench_train_4096.json 2>/dev/null
python bench.py --samples 256 --importance 256 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train_stress_256_256.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --autograd --no-render --no-cpu-baseline > $O/bench_train_autograd.json 2>/dev/null
# per-kernel micro-benchmarks (CUDA events, L2 flushed): MLP training kernels, ray kernels, fused vs two-kernel backward
python tests/bwd_bench.py 1024 192 10 > $O/bwd_bench.txt 2>&1
python tests/fused_check.py 1024 192 10 > $O/fused_check.txt 2>&1
NERF_FZ_MODE=4 python tests/fused_check.py 1024 192 6 > $O/fused_roles.txt 2>&1
python tests/ray_bench.py > $O/ray_bench.txt 2>&1
./tests/l2_bw > $O/l2_bw.txt 2>&1
# launch list of the driver's bench command (CUDA-graph replays: ncu lists the kernel nodes); SHARES, not absolutes
python bench.py --steps 3 --warmup 3 --no-render --no-cpu-baseline > $O/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/train_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-render --no-cpu-baseline > $O/ncu_train_launches.log 2>&1
# full captures, one kernel family each
python tests/fused_check.py 1024 192 1 > $O/plain_fused.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bwd_fused -s 1 -c 1 -f -o $O/bwd_fused \
    python tests/fused_check.py 1024 192 1 > $O/ncu_fused.log 2>&1
python tests/bwd_bench.py 1024 192 1 > $O/plain_bwd_bench.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_kernel|wgrad_tc|heads_wgrad" -s 10 -c 4 -f -o $O/train_kernels \
    python tests/bwd_bench.py 1024 192 1 > $O/ncu_bwd_bench.log 2>&1
python tests/tc_bench.py 16384 192 2 0 > $O/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -f -o $O/mlp_fwd \
    python tests/tc_bench.py 16384 192 2 0 > $O/ncu_fwd.log 2>&1
python tests/ray_bench.py 1 1 262144 > $O/plain_ray.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"composite_|sample_pdf" -s 1 -c 5 -f -o $O/ray_kernels \
    python tests/ray_bench.py 1 1 262144 > $O/ncu_ray.log 2>&1
ls -la $O
# multi-GPU lines: gpurun --gpus N -- bash profiles/run_multi.sh N [--with-4096]; tests/test_gpu_multi.py on >= 2 GPUs

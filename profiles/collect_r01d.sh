#!/bin/bash
# Round-1 evidence collection (final state of the round, after the pipelined-prologue forward): gpurun -- bash profiles/collect_r01d.sh
# Every ncu pass runs only after the same command exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r01d
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/tests_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python bench.py > $O/bench_train.json 2> $O/bench_train.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train_20steps.json 2> /dev/null
python bench.py --workload render --steps 3 --warmup 3 > $O/bench_render.json 2> $O/bench_render.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 20 --warmup 5 --autograd --no-cpu-baseline > $O/bench_train_autograd.json 2> /dev/null
# launch list of the default bench command (CUDA-graph replays: ncu lists the kernel nodes); shares, not absolutes
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/train_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_train_launches.log 2>&1
# full capture: fused MLP forward (inference, 16384 rays x 192 samples)
python tests/tc_bench.py 16384 192 2 0 > $O/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -f -o $O/mlp_fwd \
    python tests/tc_bench.py 16384 192 2 0 > $O/ncu_fwd.log 2>&1
# full capture: one training step's MLP kernels at the BASELINE size (1024 rays): coarse fwd, fwd-save, dgrad, wgrad, heads
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/plain_train_nograph.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_kernel|wgrad_tc|heads_wgrad|composite|sample_pdf|adam|train_prepare|pack_kernel|mse" -s 38 -c 22 -f -o $O/train_step \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_train_step.log 2>&1
# (the HBM-bound ray kernels did not change since r01c: profiles/r01c_ray_kernels_ncu_summary.csv)
ls -la $O
# the other BASELINE configs (one line each)
python bench.py --samples 256 --importance 256 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train_stress_256_256.json 2>/dev/null
python bench.py --workload render --rays 262144 --samples 256 --importance 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_render_stress_256_256.json 2>/dev/null
python bench.py --workload render --rays 10000 --steps 5 --warmup 3 > $O/bench_render_100x100_bf16.json 2>/dev/null
python bench.py --workload render --rays 10000 --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_render_100x100_fp32.json 2>/dev/null
python bench.py --rays 4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train_4096.json 2>/dev/null
python bench.py --workload render --impl reference --steps 3 --warmup 1 > $O/bench_reference_render.json 2>/dev/null
ls -la $O

#!/bin/bash
# N-GPU bench lines (torchrun, one rank per GPU):  gpurun --gpus N --timeout 900 -- bash profiles/run_multi.sh N [extra bench flags]
cd "$(dirname "$0")/.."
N=${1:-2}; shift
O=gpurun_out/r02m; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" > $O/bench_train_${N}gpu.json 2> $O/bench_train_${N}gpu.err
echo "rc=$?"
tail -c 1500 $O/bench_train_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_train_${N}gpu.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d.get("grad_exchange"), d.get("grad_exchange_note"), d.get("stage_ms"))
    print({k: d.get(k) for k in ("dp_parity",)})
    for k in ("train_4096","render"):
        if k in d: print(k, d[k].get("value"), d[k].get("ms_per_step"))
except Exception as e: print("no json", e)
PY

#!/usr/bin/env python
"""Benchmark of the NeRF hot path (BASELINE.json metric: rays/s render & train, 64 coarse + 128 importance
samples, at 1/2/4/8 B200; MLP % of bf16 tensor-core peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|render] [--impl ours|reference]

ONE JSON line on stdout (rank 0).  The default line carries BOTH halves of the metric:

  headline   BASELINE.json configs[1]: one training step on a 1024-ray batch per GPU (render coarse+fine -> MSE ->
             analytic backward -> Adam), bf16 tensor-core mode, synthetic rays, random-init 8x256 NeRF.  N > 1
             (torchrun, one rank per GPU) is data-parallel with ONE flat-gradient exchange per step, fused with Adam into one
             kernel over NVLink peer memory (`grad_exchange` says which path ran; NCCL is the fall-back) (weak scaling).
             `value` = whole-job rays/s with the batch resident in HBM; `e2e` = the same step through the public API
             (TrainStep.submit/result) with the batch copied from pinned host memory and the metrics read back every step.
  `render`   BASELINE.json configs[2]: the 800x800 (640 000-ray) render, rays sharded over the N ranks (strong
             scaling), same keys (value, ms_per_step, e2e, roofline, clocks).
  `train_4096` (N > 1, or --with-4096)  BASELINE.json configs[3]: 4096 rays per GPU, data-parallel.
  `dp_parity`  (N > 1)  numerical check on the real ranks: k data-parallel steps == k single-rank steps on the
             concatenated batch (peer-memory exchange and NCCL); sharded render == unsharded render bit for bit.

`roofline` is the TENSOR roofline of the dominant kernel of the step (SURVEY.md section 8d bounds every MLP kernel by
the tensor cores): algorithmic FLOP per launch / CUDA-event time of that kernel inside the replayed step / the measured
bf16 matmul peak of MEASURED_PEAKS.json; its HBM view (algorithmic bytes, ncu DRAM traffic) is `roofline.hbm`.
`kernels` lists every MLP kernel of the step the same way, `mlp` is their FLOP-weighted total ("MLP % of peak").
`cpu_baseline` = the reference's own CPU path on this box's host cores (oracle/_ref = the unmodified reference,
compiled by oracle/build_ref.py; the bit-identical torch port oracle/nerf_oracle_torch.py if that is absent).
`--impl reference` times that CPU path alone, on the same `config`, honouring --steps / --warmup.
`--workload render` makes the render the headline (and drops the training records).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The CPU legs use every host thread: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin the
# reference arm to one core (set before numpy / torch are first imported).
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        if os.environ.get(_v, "") in ("", "1"):
            os.environ[_v] = str(os.cpu_count())
# stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner with printf)
# get stderr as their fd 1 for the whole run; emit() writes the result line to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# ---- algorithmic work per sample row (SURVEY.md section 8d / BASELINE.md section 3) ----------------------------
FLOP_FWD = 1186816            # un-padded MACs x 2 of one forward row
FLOP_FWD_DENSITY = 982528     # layers 0-7 + sigma head only
FLOP_BWD = 2302208            # backward of one fine row: wgrad (= FLOP_FWD) + dgrad
FLOP_WGRAD = FLOP_FWD
FLOP_DGRAD = FLOP_BWD - FLOP_FWD
# distinct bf16 operand bytes per row that a wgrad fed from HBM has to read:
# dY = d(pre-act) of 8 trunk layers + d_bottleneck (9 x 512 B) + d_hv (256 B); X = h0..h7 + bottleneck (9 x 512 B) + x_enc + dir_enc
WGRAD_HBM_BYTES_PER_ROW = 9 * 512 + 256 + 9 * 512 + 128 + 128
N_SAMPLES, N_IMPORTANCE = 64, 128


def ncu_traffic():
    """DRAM bytes per launch of the MLP kernels from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per sample row)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 200 train / 5 render; reference arm: 5)")
    ap.add_argument("--warmup", type=int, default=None, help="untimed warm-up steps (default: 10 train / 3 render; reference arm: 1)")
    ap.add_argument("--workload", choices=["train", "render"], default="train")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (train) / total rays (render)")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--samples", type=int, default=64, help="coarse samples per ray (BASELINE configs[4] stress: 256)")
    ap.add_argument("--importance", type=int, default=128, help="importance samples per ray (stress: 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", choices=["auto", "reference", "torch", "numpy"], default="auto",
                    help="CPU legs: oracle/_ref (the unmodified reference) when built, else the torch restatement (auto); "
                         "or force one of them / the numpy checker")
    ap.add_argument("--no-graph", action="store_true", help="train: enqueue the step eagerly instead of replaying the CUDA graph")
    ap.add_argument("--autograd", action="store_true", help="train: the reference's loop on the drop-in classes (autograd + FlatAdam)")
    ap.add_argument("--no-render", action="store_true", help="train workload: skip the `render` record (configs[2])")
    ap.add_argument("--with-4096", action="store_true", help="train workload: add the `train_4096` record (configs[3]) at N=1 too")
    ap.add_argument("--render-steps", type=int, default=3, help="timed steps of the `render` record")
    ap.add_argument("--dp-parity-only", action="store_true", help="N > 1: run only the `dp_parity` numerical check and print it")
    args = ap.parse_args()
    args.steps_given, args.warmup_given = args.steps is not None, args.warmup is not None
    if args.steps is None:
        args.steps = 200 if args.workload == "train" else 5
    if args.warmup is None:
        args.warmup = 10 if args.workload == "train" else 3
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines, self.skip = gpu_index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples delivered so far are from before the region of interest (the sampler is started early because
        nvidia-smi needs a few hundred ms to deliver its first line)."""
        self.skip = len(self.lines)
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.skip:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own CPU path on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
# kind "reference": oracle/_ref = the UNMODIFIED reference package (nerfmlp.NeRFMLP / NeRFRenderer on device 'cpu',
#   torch.optim.Adam, the loop body of scripts/train.py:374-388), byte-compiled by oracle/build_ref.py where
#   /root/reference exists; it travels to the GPU box like the built .so.
# kind "port": oracle/nerf_oracle_torch.py, the same torch CPU kernels in the same order (bit-identical to the reference
#   on the golden vectors, tests/test_oracle_golden.py) -- used when oracle/_ref is absent.
# `--cpu-port numpy` times the numpy/OpenBLAS checker (oracle/nerf_oracle.py) instead (~2-3x slower).
CPU_RAYS = {"train": 1024, "render": 2048}


def cpu_kind(port):
    if port in ("auto", "reference"):
        from oracle import build_ref
        ok, why = build_ref.available()
        if ok:
            return "reference", ""
        if port == "reference":
            raise SystemExit(f"--cpu-port reference: {why}")
        return "torch", why
    return port, ""


def cpu_step_fn(workload, rays, port):
    import numpy as np
    from oracle import nerf_oracle as O
    p = O.init_params(0)
    o, d = O.random_rays(rays, 1)
    tgt = np.random.default_rng(2).uniform(0, 1, (rays, 3)).astype(np.float32)
    if port in ("reference", "torch"):
        import torch
        torch.set_num_threads(os.cpu_count())
        to, td_, tt = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(tgt)
    if port == "reference":
        from oracle import build_ref
        ref = build_ref.load()
        model = ref.NeRFMLP()
        model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        if workload == "train":
            renderer = ref.NeRFRenderer(model, "cpu", N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=1.0)
            opt = torch.optim.Adam(model.parameters(), lr=5e-4)              # scripts/train.py:258

            def train_ref():                                                  # scripts/train.py:374-387
                loss = torch.mean((renderer._render_rays(to, td_)["rgb_map"] - tt) ** 2)
                opt.zero_grad()
                loss.backward()
                opt.step()
            return train_ref
        renderer = ref.NeRFRenderer(model, "cpu", N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0)
        return lambda: renderer.render(to, td_, rays, 1, 1.0)                 # renderer.py:23-45 (chunk 16384)
    if port == "torch":
        from oracle import nerf_oracle_torch as T
        if workload == "train":
            tr = T.Trainer(p, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=1.0)
            return lambda: tr.step(to, td_, tt)
        pt = T.params_from_numpy(p)

        def render_t():
            with torch.no_grad():
                T.render_rays(pt, to, td_, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0)
        return render_t
    cfg = O.RenderConfig(N_samples=N_SAMPLES, N_importance=N_IMPORTANCE)
    t_vals = np.linspace(0, 1, N_SAMPLES, dtype=np.float32)
    u = np.linspace(0, 1, N_IMPORTANCE, dtype=np.float32)
    rng = np.random.default_rng(3)
    state = {"flat": O.flatten_params(p), "m": None, "v": None, "step": 0}

    def train():
        t_rand = rng.uniform(0, 1, (rays, N_SAMPLES)).astype(np.float32)
        u_r = rng.uniform(0, 1, (rays, N_IMPORTANCE)).astype(np.float32)
        prm = O.unflatten_params(state["flat"])
        _, grads, _ = O.train_grads(prm, o, d, tgt, cfg, t_vals, u_r, t_rand)
        if state["m"] is None:
            state["m"] = np.zeros_like(state["flat"]); state["v"] = np.zeros_like(state["flat"])
        state["step"] += 1
        state["flat"], state["m"], state["v"] = O.adam_step(state["flat"], O.flatten_params(grads), state["m"],
                                                            state["v"], state["step"])

    def render():
        O.render_rays(p, o, d, cfg, t_vals, u)

    return train if workload == "train" else render


def time_torch_eager_gpu(workload, dev):
    """The reference's arithmetic (oracle/nerf_oracle_torch.py, bit-identical on the golden vectors) run as eager fp32
    PyTorch on this GPU: what the reference itself would do with device=cuda.  An extra BASELINE (`torch_eager_gpu`),
    measured after the timed regions; not the product path, never part of `value`."""
    import numpy as np
    import torch
    from oracle import nerf_oracle as O
    from oracle import nerf_oracle_torch as T
    p = O.init_params(0)
    rays = 1024 if workload == "train" else 16384
    o, d = O.random_rays(rays, 1)
    to, td_ = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    if workload == "train":
        tgt = torch.from_numpy(np.random.default_rng(2).uniform(0, 1, (rays, 3)).astype(np.float32)).to(dev)
        tr = T.Trainer(p, device=dev, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=1.0)
        fn = lambda: tr.step(to, td_, tgt)
    else:
        pt = T.params_from_numpy(p, device=dev)

        def fn():
            with torch.no_grad():
                T.render_rays(pt, to, td_, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    reps = 10 if workload == "train" else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    what = ("1024-ray train step (autograd + torch.optim.Adam)" if workload == "train" else "16384-ray render chunk")
    return {"value": rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
            "what": f"{what}, {N_SAMPLES}+{N_IMPORTANCE} samples, eager fp32 PyTorch (TF32 off) on the same B200: the reference's own "
                    "arithmetic (oracle/nerf_oracle_torch.py) with device=cuda; an extra baseline, not the product path"}


def cpu_sample_text(workload, rays, port, reps):
    impl = {"reference": "the unmodified reference (oracle/_ref: nerfmlp.NeRFMLP + NeRFRenderer on device 'cpu', torch.optim.Adam), fp32",
            "torch": "torch CPU restatement of the reference (same torch ops: F.linear / autograd / torch.optim.Adam), fp32",
            "numpy": "numpy/OpenBLAS fp32 oracle port"}[port]
    what = "train step (render+MSE+backward+Adam) = configs[1] itself" if workload == "train" else \
        "render: a bounded sample of the 640000 rays of configs[2]"
    return f"{rays}-ray {what}, {N_SAMPLES}+{N_IMPORTANCE} samples, {impl}, median of {reps} steps"


def time_cpu(workload, rays, steps, warmup, port):
    fn = cpu_step_fn(workload, rays, port)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def cpu_rays(args, port):
    if port == "numpy":
        return 256 if args.workload == "train" else 512
    if args.workload == "train":
        return args.rays or CPU_RAYS["train"]
    return CPU_RAYS["render"]


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation on the host cores (all threads), same metric /
    config as the product arm, each step a bounded sample of the workload, EXACTLY --steps timed steps after --warmup
    (defaults 5 / 1 when the flags are absent: a 1024-ray CPU step takes ~1 s).  Rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    port, why = cpu_kind(args.cpu_port)
    rays = cpu_rays(args, port)
    steps = args.steps if args.steps_given else 5
    warmup = args.warmup if args.warmup_given else 1
    sec = time_cpu(args.workload, rays, steps, warmup, port)
    val = rays / sec
    line = {"impl": "reference", "metric": f"{args.workload}_rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.workload == "train" else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": os.cpu_count(),
                             "kind": "reference" if port == "reference" else "port",
                             "sample": cpu_sample_text(args.workload, rays, port, steps)},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if why:
        line["cpu_baseline"]["note"] = f"oracle/_ref unavailable ({why}); timed the bit-identical torch port"
    emit(line)


# ------------------------------------------------------------------------------------------------
# Synthetic inputs of the product arm.  Kept here so that the timed product path never imports
# anything under oracle/ (the oracle is only executed by the CPU legs above and by the torch-eager baseline).
# ------------------------------------------------------------------------------------------------
def synthetic_rays(n, seed):
    """i.i.d. origins ~ N((0,0,4), 0.1^2), non-unit directions ~ N(0,I) with d_z <- -|d_z| - 1 (exercises the |d| scaling)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    o = (np.array([0, 0, 4], np.float32) + 0.1 * rng.standard_normal((n, 3))).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d[:, 2] = -np.abs(d[:, 2]) - 1.0
    return o, d


def synthetic_pinhole_rays(H, W, fov=0.6911):
    """Pinhole camera at (0,0,4), identity rotation, looking down -z: the view scripts/render_example.py:245-250 builds."""
    import math
    import numpy as np
    focal = 0.5 * W / math.tan(0.5 * fov)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    d = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1).reshape(-1, 3).astype(np.float32)
    o = np.broadcast_to(np.array([0, 0, 4], np.float32), d.shape).copy()
    return o, d, focal


def workload_config(args, workload=None, rays=None):
    """The `config` object: a function of the command line only, so the product arm and `--impl reference` emit the
    SAME dict for the same flags (the reference arm's bounded sample is described in its cpu_baseline.sample)."""
    workload = workload or args.workload
    if workload == "train":
        rays = rays or args.rays or 1024
        return {"workload": f"train step: {rays}-ray batch per GPU, {N_SAMPLES}+{N_IMPORTANCE} samples, "
                            "render+MSE+backward+Adam (BASELINE.json configs[1]; DP all-reduce for N>1)",
                "rays_per_gpu": rays, "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE, "perturb": 1.0,
                "parallelism": f"dp{args.gpus}",
                "coarse_pass": "full network in both passes, as the reference (the product default evaluates the coarse pass "
                               "for its densities only: reported separately as density_only_coarse)",
                "l2": "per-step working set (~1 GB of saved activations) exceeds the 126 MB L2; a 256 MB buffer is also rewritten between timed steps (untimed)"}
    rays = rays or args.rays or 640000
    what = f"render {rays} rays (800x800)" if rays == 640000 else f"render {rays} rays"
    return {"workload": f"{what}, {N_SAMPLES}+{N_IMPORTANCE} samples, chunk 16384, rays sharded over ranks "
                        "(BASELINE.json configs[2])",
            "rays_total": rays, "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE, "perturb": 0.0,
            "parallelism": f"rays/{args.gpus}",
            "coarse_pass": "full network in both passes, as the reference",
            "l2": "640k rays x 256 samples stream >10 GB of intermediates per frame; a 256 MB buffer is also rewritten between timed steps (untimed)"}


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def tc_peak_for(pk, region_s):
    """Task contract: the BURST matmul figure for a kernel timed in a short region, the SUSTAINED one for a kernel timed
    inside a long back-to-back region (>= 0.2 s of device time sits in the power cap, as the 4 s sustained matmul does)."""
    long_run = region_s >= 0.2 and pk.get("bf16_tflops_sustained")
    if long_run:
        return pk["bf16_tflops_sustained"], pk["source"] + ", sustained bf16 matmul (kernel timed inside a %.2f s back-to-back region)" % region_s
    return pk["bf16_tflops"], pk["source"] + ", burst bf16 matmul (short timed region)"


def tensor_roofline(pk, kernel, flop, ms, region_s, **extra):
    tf = flop / (ms * 1e-3) / 1e12
    peak, src = tc_peak_for(pk, region_s)
    r = {"kernel": kernel, "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
         "peak_source": src, "frac_of_burst": tf / pk["bf16_tflops"],
         "frac_of_sustained": (tf / pk["bf16_tflops_sustained"]) if pk.get("bf16_tflops_sustained") else None,
         "flop_per_launch": flop, "avg_launch_ms": ms, "traffic": None}
    r.update(extra)
    return r


def bench_train(cx, args, rays, K, W, full=True):
    """One training configuration: returns the record (value, ms_per_step, e2e, stage_ms, roofline, ...)."""
    import numpy as np
    import torch
    import torch.distributed as td
    nb, ops, dev, world, rank = cx.nb, cx.ops, cx.dev, cx.world, cx.rank
    dll = nb._lib.dll()
    sampler = ClockSampler(cx.local).start()                    # started early: see ClockSampler.mark
    torch.manual_seed(0)
    model = nb.NeRFMLP(precision=args.precision).to(dev)       # random-init 8x256 (torch default init)
    if world > 1:
        nb.dist.broadcast_params(model)
    o_np, d_np = synthetic_rays(rays, 100 + rank)
    tgt_np = np.random.default_rng(200 + rank).uniform(0, 1, (rays, 3)).astype(np.float32)
    o, d, tgt = (torch.from_numpy(a).to(dev) for a in (o_np, d_np, tgt_np))
    renderer = nb.NeRFRenderer(model, dev, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=1.0,
                               coarse_density_only=False)      # headline: the whole network in both passes, as the reference
    opt = nb.FlatAdam(model, lr=5e-4)

    def step_autograd(o_, d_, tgt_):
        out = renderer._render_rays(o_, d_)
        loss = ops.mse_loss(out["rgb_map"], tgt_)               # scripts/train.py:376
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    units_per_step = rays * world
    rows_c, rows_f = rays * N_SAMPLES, rays * (N_SAMPLES + N_IMPORTANCE)
    flop_step = rows_c * FLOP_FWD + rows_f * (FLOP_FWD + FLOP_BWD)          # per GPU
    ho, hd, ht = (torch.from_numpy(a).pin_memory() for a in (o_np, d_np, tgt_np))
    if args.autograd:
        train_step = None
        run = lambda: step_autograd(o, d, tgt)

        def run_e2e():
            o_ = ho.to(dev, non_blocking=True); d_ = hd.to(dev, non_blocking=True); t_ = ht.to(dev, non_blocking=True)
            return float(step_autograd(o_, d_, t_).detach())    # loss read back: D2H + sync every step
    else:
        # the public training API: one CUDA-graph replay per step (nerf_mlp_b200.TrainStep)
        train_step = nb.TrainStep(renderer, opt, rays, graph=not args.no_graph)
        train_step.load_batch(o, d, tgt)                        # inputs resident in HBM for `value`
        run = lambda: train_step()

        def run_e2e():
            # pipelined public API: H2D of this step's batch (pinned host -> device) + replay + D2H of this
            # step's [loss, psnr, grad_norm]; the host reads the metrics of the PREVIOUS step while this one runs
            t_new = train_step.submit(ho, hd, ht)
            if run_e2e.ticket is not None:
                run_e2e.last = train_step.result(run_e2e.ticket)["loss"]
            run_e2e.ticket = t_new
            return run_e2e.last
        run_e2e.ticket, run_e2e.last = None, None
    h2d, d2h = 3 * rays * 12, (24 if not args.autograd else 4)

    # --- warm-up, then EXACTLY K timed steps (device time, per-step events, L2 flushed in between) ---
    # (the clock sampler starts before the warm-up: nvidia-smi needs a few hundred ms to deliver its first sample and a
    #  20-step timed region lasts ~20 ms; the warm-up steps run the same kernels at the same clocks)
    sampler.mark()
    for _ in range(W):
        run()
    cx.barrier()
    launches0 = dll.nerf_launch_count()
    evs = []
    cx.barrier()
    wall0 = time.perf_counter()
    for _ in range(K):
        cx.flush.zero_()                                             # L2 flush, outside the event pair
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        evs.append((e0, e1))
    cx.barrier()
    wall = time.perf_counter() - wall0
    launches = dll.nerf_launch_count() - launches0
    if train_step is not None and train_step.use_graph:
        launches = K * train_step.launches_per_step                 # replayed launches are not seen by the host-side counter
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    ms_total = torch.tensor([sum(ms_steps)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(ms_total, op=td.ReduceOp.MAX)                 # max over ranks
    ms_per_step = float(ms_total) / K
    value = units_per_step / (ms_per_step * 1e-3)
    region_s = ms_per_step * K * 1e-3

    # --- e2e: public API with host buffers, H2D + D2H inside the timed region (wall clock) ----------
    def drain_e2e():
        if getattr(run_e2e, "ticket", None) is not None:            # the last submitted step's metrics are read too
            run_e2e.last = train_step.result(run_e2e.ticket)["loss"]
            run_e2e.ticket = None

    for _ in range(2):
        run_e2e()
    drain_e2e()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        run_e2e()
    cx.barrier()
    drain_e2e()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
    e2e_val = units_per_step * K / float(e2e_s)
    # sampled over the warm-up, the timed region and the e2e loop.  If nvidia-smi delivered nothing in that window (the
    # 20-step region lasts ~20 ms and, with 8 ranks starting their samplers at once, the first sample can take longer),
    # keep replaying the same step -- untimed -- until one arrives (all ranks the same number of replays: the
    # data-parallel step is collective)
    extra = 0
    for _ in range(40):
        got = torch.tensor([1.0 if (len(sampler.lines) > sampler.skip or sampler.proc is None) else 0.0], device=dev)
        if world > 1:
            td.all_reduce(got, op=td.ReduceOp.MIN)
        if float(got) > 0:
            break
        for _ in range(50):
            run()
        torch.cuda.synchronize()
        extra += 50
    clocks = sampler.stop()
    if extra:
        clocks["note"] = (f"no nvidia-smi sample fell into the timed region; sampled over {extra} further untimed replays of "
                          "the same step directly after it")

    pk = peaks()
    rec = {"value": value, "unit": "rays/s", "ms_per_step": ms_per_step, "steps": K, "warmup": W,
           "step_tflops": flop_step / (ms_per_step * 1e-3) / 1e12,
           "wall_s_timed_region": wall, "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "step_api": ("nerf_mlp_b200.TrainStep (" + ("CUDA graph replay" if train_step.use_graph else "eager launches") + ")")
           if train_step is not None else "NeRFRenderer._render_rays + loss.backward() + FlatAdam.step (autograd)"}
    rec["step_frac_of_burst"] = rec["step_tflops"] / pk["bf16_tflops"]
    if train_step is None:
        return rec

    # --- per-stage device times: K more replays of the same step re-captured with an event record between the
    # kernels (the records cost ~4 us each, so they stay out of the timed region above)
    stage_acc = {}
    train_step.stage_events = True
    if train_step.use_graph:
        train_step.recapture()
    for _ in range(2):
        run()
    for _ in range(K):
        cx.flush.zero_()
        run()
        torch.cuda.synchronize()
        for k_, v_ in train_step.stage_times().items():
            stage_acc[k_] = stage_acc.get(k_, 0.0) + v_ / K
    cx.barrier()
    rec["stage_ms"] = {k_: round(v_, 5) for k_, v_ in stage_acc.items()}
    if cx.world > 1:
        if getattr(train_step, "_peer", None) is not None:
            rec["grad_exchange"] = (("two-shot (reduce-scatter + all-gather)" if train_step._peer.get("two_shot") else "one-shot")
                                    + " all-reduce over NVLink peer memory fused into the Adam kernel "
                                    "(nerf_adam_step_fused_peer; no NCCL call in the step)")
        else:
            rec["grad_exchange"] = ("NCCL all_reduce " + ("captured in the step graph" if train_step.allreduce_in_graph
                                                          else "between two graph replays"))
            if getattr(train_step, "peer_allreduce_error", None):
                rec["grad_exchange_note"] = "peer-memory path unavailable: " + train_step.peer_allreduce_error
    rec["stage_ms_note"] = ("mean device time per stage over %d instrumented replays after the timed region "
                            "(CUDA event records on the launching stream between the kernels of the replayed graph)" % K)

    # --- roofline: every MLP kernel of the step against the tensor peak; `roofline` = the dominant one ----------
    traffic = ncu_traffic()
    stage_kernels = [   # stage name, kernel, rows, FLOP per row
        ("mlp_fwd_coarse", "mlp_tc_kernel<fwd> coarse pass (fused PE + 8x256 MLP + heads)", rows_c, FLOP_FWD),
        ("mlp_fwd_fine_save", "mlp_tc_kernel<fwd,save> fine pass (+ saved activations / ReLU masks)", rows_f, FLOP_FWD),
        ("mlp_bwd_dgrad", "mlp_tc_kernel<dgrad> (9-GEMM chain over the transposed weights)", rows_f, FLOP_DGRAD),
        ("mlp_bwd_wgrad", "wgrad_tc_kernel (dW = dY^T X over sample rows, tcgen05; + heads_wgrad beside it)", rows_f, FLOP_WGRAD),
        ("mlp_bwd", "mlp_bwd_fused_kernel (dgrad chain + dW = dY^T X in one launch)", rows_f, FLOP_BWD),
    ]
    kernels, mlp_ms, mlp_flop = [], 0.0, 0.0
    for st_name, kname, rows, fpr in stage_kernels:
        if st_name in stage_acc and args.precision == "bf16":
            r = tensor_roofline(pk, kname, rows * fpr, stage_acc[st_name], region_s, stage=st_name, rows_per_launch=rows,
                                flop_per_row=fpr, share_of_step=stage_acc[st_name] / (sum(ms_steps) / K))
            tr = traffic.get(st_name)
            if tr:
                r["traffic"] = tr["dram_bytes_per_row"] * rows
                r["traffic_source"] = tr.get("source")
            kernels.append(r)
            mlp_ms += stage_acc[st_name]
            mlp_flop += rows * fpr
    if kernels:
        dom = max(kernels, key=lambda r: r["avg_launch_ms"])
        roofline = dict(dom)
        if dom["stage"] == "mlp_bwd_wgrad":
            gbs = rows_f * WGRAD_HBM_BYTES_PER_ROW / (dom["avg_launch_ms"] * 1e-3) / 1e9
            roofline["hbm"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                               "bytes_per_row": WGRAD_HBM_BYTES_PER_ROW,
                               "note": "secondary view: the bf16 operands a wgrad fed from HBM must read (dY written by the dgrad "
                                       "kernel + X saved by the forward); this is spill induced by the three-kernel design, not "
                                       "algorithmic work of the path (SURVEY.md 8d bounds the MLP backward by the tensor cores)"}
        rec["roofline"] = roofline
        rec["kernels"] = kernels
        peak, src = tc_peak_for(pk, region_s)
        rec["mlp"] = {"what": "all MLP kernels of the step (FLOP-weighted): north_star's 'MLP % of bf16 tensor-core peak'",
                      "ms_per_step": mlp_ms, "tflops": mlp_flop / (mlp_ms * 1e-3) / 1e12,
                      "frac": mlp_flop / (mlp_ms * 1e-3) / 1e12 / peak, "peak": peak, "peak_source": src,
                      "frac_of_burst": mlp_flop / (mlp_ms * 1e-3) / 1e12 / pk["bf16_tflops"]}
    if not full:
        return rec

    # --- the product default: coarse pass evaluated for its densities only (identical outputs of render() and of
    # the training step; 17 % fewer coarse-pass FLOPs).  Reported separately; the headline above does the
    # reference's full work in both passes.
    renderer.coarse_density_only = True
    train_step.stage_events = False
    if train_step.use_graph:
        train_step.recapture()
    for _ in range(3):
        run()
    cx.barrier()
    evs2 = []
    for _ in range(K):
        cx.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        evs2.append((e0, e1))
    cx.barrier()
    ms2 = torch.tensor([sum(a.elapsed_time(b) for a, b in evs2)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(ms2, op=td.ReduceOp.MAX)
    rec["density_only_coarse"] = {"value": units_per_step / (float(ms2) / K * 1e-3), "unit": "rays/s", "ms_per_step": float(ms2) / K,
                                  "note": "coarse pass in NERF_FWD_DENSITY_ONLY mode (product default of render() / TrainStep): same pixels / "
                                          "same loss, bottleneck+view+rgb layers of the coarse pass not evaluated"}
    return rec


def bench_render(cx, args, total, K, W):
    """BASELINE configs[2]: `total` rays sharded over the ranks (strong scaling), chunk loop of renderer.render."""
    import torch
    import torch.distributed as td
    nb, ops, dev, world, rank = cx.nb, cx.ops, cx.dev, cx.world, cx.rank
    dll = nb._lib.dll()
    torch.manual_seed(0)
    model = nb.NeRFMLP(precision=args.precision).to(dev)
    if world > 1:
        nb.dist.broadcast_params(model)
    Himg = int(round(total ** 0.5))
    if Himg * Himg != total:
        Himg, Wimg = total, 1
    else:
        Wimg = Himg
    o_np, d_np, focal = synthetic_pinhole_rays(Himg, Wimg) if Wimg > 1 else (*synthetic_rays(total, 1), 1.0)
    lo, hi = nb.dist.shard_range(total, rank, world)
    o, d = torch.from_numpy(o_np[lo:hi]).to(dev), torch.from_numpy(d_np[lo:hi]).to(dev)
    renderer = nb.NeRFRenderer(model, dev, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0,
                               coarse_density_only=False)      # the whole network in both passes, as the reference
    n_local = hi - lo

    def run():
        return renderer.render(o, d, n_local, 1, focal)          # chunk loop of renderer.py:40-44 on this rank's block

    ho, hd = torch.from_numpy(o_np[lo:hi]).pin_memory(), torch.from_numpy(d_np[lo:hi]).pin_memory()
    hout = torch.empty((n_local, 1, 3), dtype=torch.float32).pin_memory()

    def run_e2e():
        img = renderer.render(ho.to(dev, non_blocking=True), hd.to(dev, non_blocking=True), n_local, 1, focal)
        hout.copy_(img, non_blocking=True)
        torch.cuda.synchronize()
        return hout

    # per-kernel timing hook: CUDA events (current stream = the launching stream) around the fused MLP forward launches
    mlp_events = []
    orig_fwd = ops.mlp_fwd_rays

    def timed_fwd(model_, rays_o, rays_d, z_vals, *a, **k):
        if not timed_fwd.on:
            return orig_fwd(model_, rays_o, rays_d, z_vals, *a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_fwd(model_, rays_o, rays_d, z_vals, *a, **k)
        e1.record()
        mlp_events.append((z_vals.numel(), e0, e1))
        return out

    timed_fwd.on = False
    ops.mlp_fwd_rays = timed_fwd
    try:
        sampler = ClockSampler(cx.local).start()
        for _ in range(W):
            run()
        cx.barrier()
        launches0 = dll.nerf_launch_count()
        timed_fwd.on = True
        evs = []
        cx.barrier()
        wall0 = time.perf_counter()
        for _ in range(K):
            cx.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            evs.append((e0, e1))
        cx.barrier()
        wall = time.perf_counter() - wall0
        timed_fwd.on = False
        launches = dll.nerf_launch_count() - launches0
        ms_steps = [a.elapsed_time(b) for a, b in evs]
        ms_total = torch.tensor([sum(ms_steps)], device=dev, dtype=torch.float64)
        if world > 1:
            td.all_reduce(ms_total, op=td.ReduceOp.MAX)
        ms_per_step = float(ms_total) / K
        value = total / (ms_per_step * 1e-3)
        for _ in range(2):
            run_e2e()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            run_e2e()
        cx.barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
        clocks = sampler.stop()
        # density-only coarse pass (product default of render())
        renderer.coarse_density_only = True
        for _ in range(3):
            run()
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cx.flush.zero_()
        e0.record()
        for _ in range(K):
            run()
        e1.record()
        cx.barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
        if world > 1:
            td.all_reduce(ms2, op=td.ReduceOp.MAX)
    finally:
        ops.mlp_fwd_rays = orig_fwd
    pk = peaks()
    flop_per_ray = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_FWD
    rows_max = max(n for n, _, _ in mlp_events)
    fine = [a.elapsed_time(b) for n, a, b in mlp_events if n == rows_max]
    avg_ms = sum(fine) / len(fine)
    mlp_ms_per_step = sum(a.elapsed_time(b) for _, a, b in mlp_events) / K
    step_ms = sum(ms_steps) / K
    region_s = step_ms * K * 1e-3
    roofline = tensor_roofline(pk, "mlp_tc_kernel<fwd> (fused PE + 8x256 MLP + heads), fine pass of one 16384-ray chunk"
                               if args.precision == "bf16" else "fp32 check-mode SGEMM chain",
                               rows_max * FLOP_FWD, avg_ms, region_s, rows_per_launch=rows_max, flop_per_row=FLOP_FWD,
                               share_of_step=mlp_ms_per_step / step_ms)
    tr = ncu_traffic().get("mlp_fwd_infer")
    if tr and args.precision == "bf16":
        roofline["traffic"] = tr["dram_bytes_per_row"] * rows_max
        roofline["traffic_source"] = tr.get("source")
    return {"metric": "render_rays_per_s", "value": value, "unit": "rays/s", "ms_per_step": ms_per_step, "steps": K, "warmup": W,
            "scaling": "strong", "higher_is_better": True, "config": workload_config(args, "render", total),
            "step_tflops": total * flop_per_ray / (ms_per_step * 1e-3) / 1e12 / world,
            "step_frac_of_burst": total * flop_per_ray / (ms_per_step * 1e-3) / 1e12 / world / pk["bf16_tflops"],
            "wall_s_timed_region": wall, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total * K / float(e2e_s), "unit": "rays/s", "h2d_bytes_per_step": 2 * n_local * 12,
                    "d2h_bytes_per_step": n_local * 12},
            "roofline": roofline,
            "density_only_coarse": {"value": total / (float(ms2) * 1e-3), "unit": "rays/s", "ms_per_step": float(ms2),
                                    "note": "product default of render(): coarse pass evaluated for its densities only (same pixels)"}}


def dp_parity(cx, args):
    """Numerical check on the REAL ranks (scripts/train.py:376 global-mean semantics):
      * k data-parallel TrainStep steps (R rays per rank; gradient exchange over peer memory fused into the Adam kernel,
        and again with NCCL's all-reduce) leave bit-identical parameters on every rank, and they equal k single-rank
        steps on the concatenated N*R-ray batch up to summation order;
      * render_sharded over the ranks == the unsharded render, bit for bit."""
    import numpy as np
    import torch
    import torch.distributed as td
    nb, dev, world, rank = cx.nb, cx.dev, cx.world, cx.rank
    R, k = 256, 3
    batches = [synthetic_rays(R, 900 + r) for r in range(world)]
    tgts = [np.random.default_rng(950 + r).uniform(0, 1, (R, 3)).astype(np.float32) for r in range(world)]
    res = {}
    exchange = {}
    for kind in ("dp", "dp_nccl", "single"):
        torch.manual_seed(5)
        m = nb.NeRFMLP(precision=args.precision).to(dev)
        nb.dist.broadcast_params(m)
        p0 = m.flat_params.detach().clone()
        r_ = nb.NeRFRenderer(m, dev, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0)
        if kind in ("dp", "dp_nccl"):
            opt = nb.FlatAdam(m, lr=5e-4)
            o, d, t = batches[rank][0], batches[rank][1], tgts[rank]
            n = R
        else:
            opt = nb.FlatAdam(m, lr=5e-4, world_size=1)                    # no all-reduce: the whole batch on this rank
            o = np.concatenate([b[0] for b in batches]); d = np.concatenate([b[1] for b in batches]); t = np.concatenate(tgts)
            n = R * world
        step = nb.TrainStep(r_, opt, n, peer_allreduce=(kind != "dp_nccl"))   # "dp": peer-memory exchange if available
        exchange[kind] = "peer" if getattr(step, "_peer", None) is not None else "nccl"
        o, d, t = (torch.from_numpy(a).to(dev) for a in (o, d, t))
        losses = []
        for _ in range(k):
            step(o, d, t)
            losses.append(step.read_metrics()["loss"])
        torch.cuda.synchronize()
        res[kind] = (m.flat_params.detach().clone(), losses, p0)
    p_dp, l_dp, p0 = res["dp"]
    p_1, l_1, _ = res["single"]
    p_nccl = res["dp_nccl"][0]
    # the two gradient-exchange paths differ in the order of the `world` summands (and, like any two runs, in the
    # split-K reduction order of the weight-gradient kernels)
    rel_paths = float((p_dp - p_nccl).norm() / (p_1 - p0).norm())
    ref_n = p_nccl.clone()
    td.broadcast(ref_n, src=0)
    nccl_ranks_differ = torch.tensor([float((ref_n != p_nccl).sum())], device=dev, dtype=torch.float64)
    td.all_reduce(nccl_ranks_differ, op=td.ReduceOp.SUM)
    # (a) identical across ranks: compare with rank 0's copy
    ref0 = p_dp.clone()
    td.broadcast(ref0, src=0)
    ranks_differ = torch.tensor([float((ref0 != p_dp).sum())], device=dev, dtype=torch.float64)
    td.all_reduce(ranks_differ, op=td.ReduceOp.SUM)
    # (b) DP == single-rank on the concatenated batch
    diff = (p_dp - p_1).abs()
    upd = (p_1 - p0)
    rel_update = float((p_dp - p_1).norm() / upd.norm())
    # the DP loss of a rank is its own batch mean; the mean over ranks is the single-rank loss
    l_mean = torch.tensor(l_dp, device=dev, dtype=torch.float64)
    td.all_reduce(l_mean, op=td.ReduceOp.SUM)
    l_mean = (l_mean / world).tolist()
    loss_rel = max(abs(a - b) / abs(b) for a, b in zip(l_mean, l_1))
    # (c) sharded render == unsharded, bit for bit
    Hh = 96
    o_np, d_np, focal = synthetic_pinhole_rays(Hh, Hh)
    torch.manual_seed(6)
    m = nb.NeRFMLP(precision=args.precision).to(dev)
    nb.dist.broadcast_params(m)
    rr = nb.NeRFRenderer(m, dev, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0)
    oo, dd = torch.from_numpy(o_np).to(dev), torch.from_numpy(d_np).to(dev)
    full = rr.render(oo, dd, Hh, Hh, focal)
    shard = nb.dist.render_sharded(nb.dist.make_render_fn(rr), oo, dd, Hh, Hh, focal)
    render_mismatch = torch.tensor([float((full != shard).sum())], device=dev, dtype=torch.float64)
    td.all_reduce(render_mismatch, op=td.ReduceOp.SUM)
    out = {"world": world, "rays_per_rank": R, "steps": k, "perturb": 0.0,
           "grad_exchange": exchange["dp"], "params_differ_across_ranks": int(ranks_differ),
           "nccl_path_params_differ_across_ranks": int(nccl_ranks_differ), "peer_vs_nccl_rel_l2_of_update": rel_paths,
           "dp_vs_single_max_abs": float(diff.max()),
           "dp_vs_single_frac_le_1e-6": float((diff <= 1e-6).float().mean()),
           "dp_vs_single_rel_l2_of_update": rel_update, "loss_rel_diff_max": loss_rel,
           "sharded_render_mismatching_values": int(render_mismatch),
           "ok": bool(int(ranks_differ) == 0 and int(nccl_ranks_differ) == 0 and rel_update <= 2e-2 and rel_paths <= 2e-2
                      and loss_rel <= 1e-4 and int(render_mismatch) == 0),
           "note": "params after 3 Adam steps; Adam's first steps are sign-like (lr*g/(|g|+eps)), so elements whose gradient is "
                   "near 0 amplify summation-order noise: the gate is on the relative L2 of the whole update, max/fraction are reported"}
    return out


def main():
    global N_SAMPLES, N_IMPORTANCE
    args = parse()
    N_SAMPLES, N_IMPORTANCE = args.samples, args.importance
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as td
    import nerf_mlp_b200 as nb
    from nerf_mlp_b200 import ops

    cx = Ctx()
    cx.nb, cx.ops = nb, ops
    cx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = rank = int(os.environ.get("RANK", "0"))
    cx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    cx.dev = dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    cx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
    cx.barrier = barrier
    K, W = args.steps, max(args.warmup, 3)
    if args.dp_parity_only:
        if world < 2:
            raise SystemExit("--dp-parity-only needs N > 1 ranks (torchrun)")
        rec = dp_parity(cx, args)
        if rank == 0:
            emit({"dp_parity": rec})
        finish(world)
        return

    common = {"n_gpus": world, "higher_is_better": True, "vs_baseline": None, "dtype": args.precision, "data": "synthetic"}
    if args.workload == "train":
        rec = bench_train(cx, args, args.rays or 1024, K, W)
        line = {"metric": "train_rays_per_s", "value": rec.pop("value"), "unit": rec.pop("unit"), **common,
                "steps": K, "warmup": W, "ms_per_step": rec.pop("ms_per_step"), "scaling": "weak",
                "config": workload_config(args)}
        rec.pop("steps"); rec.pop("warmup")
        line.update(rec)
        if not args.no_render and (N_SAMPLES, N_IMPORTANCE) == (64, 128):
            line["render"] = bench_render(cx, args, 640000, args.render_steps, 3)
        if (world > 1 or args.with_4096) and not args.autograd and (args.rays or 1024) != 4096:
            r4 = bench_train(cx, args, 4096, min(K, 50), 5, full=False)
            r4["config"] = workload_config(args, "train", 4096)
            r4["config"]["workload"] = r4["config"]["workload"].replace("configs[1]", "configs[3]: 4096 rays/GPU data-parallel")
            r4.update({"metric": "train_rays_per_s", "scaling": "weak"})
            line["train_4096"] = r4
        if world > 1 and not args.autograd:
            try:
                line["dp_parity"] = dp_parity(cx, args)
            except Exception as exc:  # the check must never take the bench line down; a failure is reported as such
                line["dp_parity"] = {"ok": False, "error": repr(exc)[:300]}
    else:
        rec = bench_render(cx, args, args.rays or 640000, K, W)
        line = {"metric": "render_rays_per_s", "value": rec.pop("value"), "unit": rec.pop("unit"), **common,
                "steps": K, "warmup": W, "ms_per_step": rec.pop("ms_per_step"), "scaling": "strong",
                "config": workload_config(args)}
        for k_ in ("steps", "warmup", "metric", "scaling", "higher_is_better", "config"):
            rec.pop(k_, None)
        line.update(rec)

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        port, why = cpu_kind(args.cpu_port)
        rays_cpu = cpu_rays(args, port)
        sec = time_cpu(args.workload, rays_cpu, 3, 1, port)
        line["cpu_baseline"] = {"value": rays_cpu / sec, "unit": "rays/s", "cores": os.cpu_count(),
                                "kind": "reference" if port == "reference" else "port",
                                "sample": cpu_sample_text(args.workload, rays_cpu, port, 3)}
        if why:
            line["cpu_baseline"]["note"] = f"oracle/_ref unavailable ({why}); timed the bit-identical torch port"
    if rank == 0 and not args.no_cpu_baseline and world == 1 and (N_SAMPLES, N_IMPORTANCE) == (64, 128):
        try:
            line["torch_eager_gpu"] = time_torch_eager_gpu(args.workload, dev)
        except Exception as exc:  # a baseline must never take the bench line down
            line["torch_eager_gpu"] = {"unavailable": repr(exc)[:200]}
    if rank == 0:
        emit(line)
    finish(world)


def finish(world):
    """Leave a multi-rank run without tearing the process group down: destroy_process_group() with CUDA graphs that hold
    captured NCCL kernels still alive was seen to hang after the result line had been printed (the line is out, every
    rank has passed the last barrier, nothing is left to flush), so the ranks exit hard with status 0."""
    if world > 1:
        import torch
        import torch.distributed as td
        td.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()

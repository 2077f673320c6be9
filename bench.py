#!/usr/bin/env python
"""Benchmark of the NeRF hot path (BASELINE.json metric: rays/s, 64 coarse + 128 importance samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|render] [--impl ours|reference]

Default workload = BASELINE.json configs[1]: one training step on a 1024-ray batch per GPU
(render coarse+fine -> MSE -> backward -> Adam), bf16 tensor-core mode, synthetic rays, random-init
8x256 NeRF.  N > 1 (launched by torchrun, one rank per GPU) is data-parallel with ONE flat gradient
all-reduce per step (weak scaling: 1024 rays per GPU).  `--workload render` times BASELINE.json
configs[2], the 800x800 (640k-ray) render sharded over the ranks (strong scaling).

One JSON line on stdout (rank 0).  `value` = whole-job rays/s with inputs resident in HBM;
`e2e` = the same step through the public API with rays/targets copied from pinned host memory and
the loss read back every step; `roofline` = the fused MLP forward kernel (fine pass) against the
measured bf16 tensor peak; `cpu_baseline` = the torch-CPU restatement of the reference timed on this box's host cores.
`--impl reference` times that CPU path alone (the reference's own algorithm on the host cores).
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
# numpy/OpenBLAS reference arm to one core (set before numpy is first imported).
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


FLOP_PER_ROW_FWD = 1186816          # un-padded MACs x 2
FLOP_PER_ROW_BWD = 2302208
# wgrad: every distinct bf16 operand once per sample row: dY = d(pre-act) of 8 trunk
# layers + d_bottleneck (9 x 512 B) + d_hv (256 B); X = h0..h7 + bottleneck (9 x 512 B) + x_enc (128 B) + dir_enc (128 B)
WGRAD_ALG_BYTES_PER_ROW = 9 * 512 + 256 + 9 * 512 + 128 + 128
# measured DRAM traffic (ncu dram__bytes_read.sum + dram__bytes_write.sum), see profiles/
NCU_BYTES_PER_ROW_FWD_SAVE = 5162.0      # 1.0149 GB on the 196 608-row save-mode launch
NCU_BYTES_PER_ROW_WGRAD = 10442.0        # 2.0530 GB on the same rows
NCU_BYTES_FWD_INFER_3145728 = 27.07e6    # whole 3 145 728-row inference launch (22.71 MB read + 4.36 MB written)
N_SAMPLES, N_IMPORTANCE = 64, 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 200 train / 5 render)")
    ap.add_argument("--warmup", type=int, default=None, help="untimed warm-up steps (default: 10 train / 3 render)")
    ap.add_argument("--workload", choices=["train", "render"], default="train")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (train) / total rays (render)")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--samples", type=int, default=64, help="coarse samples per ray (BASELINE configs[4] stress: 256)")
    ap.add_argument("--importance", type=int, default=128, help="importance samples per ray (stress: 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", choices=["torch", "numpy"], default="torch",
                    help="CPU legs: torch-CPU restatement of the reference (default) or the numpy checker")
    ap.add_argument("--no-graph", action="store_true", help="train: enqueue the step eagerly instead of replaying the CUDA graph")
    ap.add_argument("--autograd", action="store_true", help="train: the reference's loop on the drop-in classes (autograd + FlatAdam)")
    args = ap.parse_args()
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
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

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
        for ln in self.lines:
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
# CPU baseline: the reference's algorithm on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
# Timed port = oracle/nerf_oracle_torch.py: the same torch CPU kernels, in the same order, as the reference's
# nerfmlp/model.py + renderer.py + the loop body of scripts/train.py (autograd backward, torch.optim.Adam);
# bit-identical to the reference on the golden vectors (tests/test_oracle_golden.py::test_torch_port_*).
# `--cpu-port numpy` times the numpy/OpenBLAS checker (oracle/nerf_oracle.py) instead (~2-3x slower).
CPU_RAYS = {"train": 1024, "render": 2048}


def cpu_step_fn(workload, rays, port="torch"):
    import numpy as np
    from oracle import nerf_oracle as O
    p = O.init_params(0)
    o, d = O.random_rays(rays, 1)
    tgt = np.random.default_rng(2).uniform(0, 1, (rays, 3)).astype(np.float32)
    if port == "torch":
        import torch
        from oracle import nerf_oracle_torch as T
        torch.set_num_threads(os.cpu_count())
        to, td_, tt = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(tgt)
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
    """The same restatement (= the reference's arithmetic, bit-identical on the golden vectors) run as eager fp32
    PyTorch on this GPU: what the reference itself would do if its device were set to cuda.  An extra BASELINE
    (reported as `torch_eager_gpu`), measured after the timed regions; not the product path, never part of `value`."""
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
    impl = ("torch CPU restatement of the reference (same torch ops: F.linear / autograd / torch.optim.Adam), fp32"
            if port == "torch" else "numpy/OpenBLAS fp32 oracle port")
    what = "train step (render+MSE+backward+Adam)" if workload == "train" else "render"
    return f"{rays}-ray {what}, {N_SAMPLES}+{N_IMPORTANCE} samples, {impl}, median of {reps}"


def time_cpu(workload, rays, steps, warmup, port="torch"):
    fn = cpu_step_fn(workload, rays, port)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def run_reference(args):
    """`--impl reference`: the reference's algorithm on the host cores (the torch-CPU restatement; the
    reference checkout itself does not exist on the GPU box), same metric/config, each step a bounded
    sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rays = args.rays or (CPU_RAYS[args.workload] if args.cpu_port == "torch" else (256 if args.workload == "train" else 512))
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    sec = time_cpu(args.workload, rays, steps, warmup, args.cpu_port)
    val = rays / sec
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": f"{args.workload}_rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, rays_override=rays),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": cpu_sample_text(args.workload, rays, args.cpu_port, steps)},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


def workload_config(args, rays_override=None):
    if args.workload == "train":
        rays = rays_override or args.rays or 1024
        return {"workload": f"train step: {rays}-ray batch per GPU, {N_SAMPLES}+{N_IMPORTANCE} samples, "
                            "render+MSE+backward+Adam (BASELINE.json configs[1]; DP all-reduce for N>1)",
                "rays_per_gpu": rays, "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE, "perturb": 1.0,
                "parallelism": f"dp{args.gpus}",
                "l2": "per-step working set (~1 GB of saved activations) exceeds the 126 MB L2; a 256 MB buffer is also rewritten between timed steps (untimed)"}
    rays = rays_override or args.rays or 640000
    what = f"render {rays} rays (800x800)" if rays == 640000 else f"render, bounded sample of {rays} of the 640000 rays of an 800x800 view"
    return {"workload": f"{what}, {N_SAMPLES}+{N_IMPORTANCE} samples, chunk 16384, rays sharded over ranks "
                        "(BASELINE.json configs[2])",
            "rays_total": rays, "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE, "perturb": 0.0,
            "parallelism": f"rays/{args.gpus}",
            "l2": "640k rays x 256 samples stream >10 GB of intermediates per frame; a 256 MB buffer is also rewritten between timed steps (untimed)"}


# ------------------------------------------------------------------------------------------------
def main():
    global N_SAMPLES, N_IMPORTANCE
    args = parse()
    N_SAMPLES, N_IMPORTANCE = args.samples, args.importance
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as td
    import nerf_mlp_b200 as nb
    from nerf_mlp_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = nb.NeRFMLP(precision=args.precision).to(dev)       # random-init 8x256 (torch default init)
    if world > 1:
        nb.dist.broadcast_params(model)
    dll = nb._lib.dll()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K, W = args.steps, max(args.warmup, 3)

    # --- per-kernel timing hook: CUDA events around the fused MLP forward launches -----------------
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

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    if args.workload == "train":
        rays = args.rays or 1024
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
        flop_per_unit = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_PER_ROW_FWD + (N_SAMPLES + N_IMPORTANCE) * FLOP_PER_ROW_BWD
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
    else:
        total = args.rays or 640000
        Himg = int(round(total ** 0.5))
        if Himg * Himg != total:
            Himg, Wimg = total, 1
        else:
            Wimg = Himg
        o_np, d_np, focal = synthetic_pinhole_rays(Himg, Wimg) if Wimg > 1 else (*synthetic_rays(total, 1), 1.0)
        lo, hi = nb.dist.shard_range(total, rank, world)
        o, d = torch.from_numpy(o_np[lo:hi]).to(dev), torch.from_numpy(d_np[lo:hi]).to(dev)
        renderer = nb.NeRFRenderer(model, dev, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0.0,
                                   coarse_density_only=False)      # headline: the whole network in both passes, as the reference
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
        units_per_step = total
        flop_per_unit = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_PER_ROW_FWD
        h2d, d2h = 2 * n_local * 12, n_local * 12

    # --- warm-up, then EXACTLY K timed steps (device time, per-step events, L2 flushed in between) ---
    for _ in range(W):
        run()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = dll.nerf_launch_count()
    timed_fwd.on = True
    evs = []
    stage_acc = {} if (args.workload == "train" and train_step is not None) else None
    barrier()
    wall0 = time.perf_counter()
    for _ in range(K):
        flush.zero_()                                                # L2 flush, outside the event pair
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    timed_fwd.on = False
    launches = dll.nerf_launch_count() - launches0
    if stage_acc is not None and train_step.use_graph:
        launches = K * train_step.launches_per_step                 # replayed launches are not seen by the host-side counter
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    ms_total = torch.tensor([sum(ms_steps)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(ms_total, op=td.ReduceOp.MAX)                 # max over ranks
    ms_per_step = float(ms_total) / K
    value = units_per_step / (ms_per_step * 1e-3)

    # --- e2e: public API with host buffers, H2D + D2H inside the timed region (wall clock) ----------
    def drain_e2e():
        if getattr(run_e2e, "ticket", None) is not None:            # the last submitted step's metrics are read too
            run_e2e.last = train_step.result(run_e2e.ticket)["loss"]
            run_e2e.ticket = None

    for _ in range(2):
        run_e2e()
    drain_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        run_e2e()
    barrier()
    drain_e2e()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
    e2e_val = units_per_step * K / float(e2e_s)
    clocks = sampler.stop()                                         # sampled over the timed region and the e2e loop

    # --- per-stage device times: K more replays of the same step re-captured with an event record between the
    # kernels (the records cost ~4 us each, so they stay out of the timed region above)
    if stage_acc is not None:
        train_step.stage_events = True
        if train_step.use_graph:
            train_step.recapture()
        for _ in range(2):
            run()
        for _ in range(K):
            flush.zero_()
            run()
            torch.cuda.synchronize()
            for k_, v_ in train_step.stage_times().items():
                stage_acc[k_] = stage_acc.get(k_, 0.0) + v_ / K
        barrier()

    # --- the product default: coarse pass evaluated for its densities only (identical outputs of render() and of
    # the training step; 17 % fewer coarse-pass FLOPs).  Reported separately; the headline above does the
    # reference's full work in both passes.
    timed_fwd.on = False
    renderer.coarse_density_only = True
    if args.workload == "train" and train_step is not None:
        train_step.stage_events = False
        if train_step.use_graph:
            train_step.recapture()
    for _ in range(3):
        run()
    barrier()
    evs2 = []
    for _ in range(K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        evs2.append((e0, e1))
    barrier()
    ms2 = torch.tensor([sum(a.elapsed_time(b) for a, b in evs2)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(ms2, op=td.ReduceOp.MAX)
    density_only = {"value": units_per_step / (float(ms2) / K * 1e-3), "unit": "rays/s", "ms_per_step": float(ms2) / K,
                    "note": "coarse pass in NERF_FWD_DENSITY_ONLY mode (product default of render() / TrainStep): same pixels / "
                            "same loss, bottleneck+view+rgb layers of the coarse pass not evaluated"}

    # --- roofline -----------------------------------------------------------------------------------
    # render: the dominant kernel is the fused MLP forward (tensor-bound).  train: the largest launch
    # of the step is the weight-gradient kernel, which -- like the save-mode forward and the dgrad
    # kernel around it -- is bound by HBM bytes; its roofline is `roofline`, and
    # the tensor-core view of the fused forward (north_star's "% of bf16 peak") is `roofline_tensor`.
    pk = peaks()
    if stage_acc is not None:
        rows_max = rays * (N_SAMPLES + N_IMPORTANCE)
        avg_ms = stage_acc["mlp_fwd_fine_save"]
        mlp_ms_per_step = stage_acc["mlp_fwd_fine_save"] + stage_acc["mlp_fwd_coarse"]
    else:
        rows_max = max(n for n, _, _ in mlp_events)
        fine = [(n, a.elapsed_time(b)) for n, a, b in mlp_events if n == rows_max]
        avg_ms = sum(ms for _, ms in fine) / len(fine)
        mlp_ms_per_step = sum(a.elapsed_time(b) for _, a, b in mlp_events) / K
    achieved = rows_max * FLOP_PER_ROW_FWD / (avg_ms * 1e-3) / 1e12
    step_ms = sum(ms_steps) / K
    # Denominator (task contract): the BURST matmul figure for a kernel timed in a short region, the SUSTAINED one for a
    # kernel timed inside a long back-to-back region (>= 0.2 s of device time: the 200-step training run and the 800x800
    # render sit in the power cap, as the 4 s sustained matmul measurement does).  Both fractions are always reported.
    long_run = (step_ms * K * 1e-3) >= 0.2 and pk.get("bf16_tflops_sustained")
    tc_peak = pk["bf16_tflops_sustained"] if long_run else pk["bf16_tflops"]
    tc_peak_kind = (", sustained bf16 matmul (kernel timed inside a %.2f s back-to-back region)" % (step_ms * K * 1e-3)) if long_run \
        else ", burst bf16 matmul (short timed region)"
    roofline_tensor = {"kernel": ("mlp_tc_kernel<fwd" + (", save" if args.workload == "train" else "") + "> (fused PE + 8x256 MLP + heads), fine pass")
                       if args.precision == "bf16" else "fp32 check-mode SGEMM chain",
                       "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                       "frac": achieved / tc_peak, "peak_source": pk["source"] + tc_peak_kind,
                       "frac_of_burst": achieved / pk["bf16_tflops"],
                       "frac_of_sustained": (achieved / pk["bf16_tflops_sustained"]) if pk.get("bf16_tflops_sustained") else None,
                       "rows_per_launch": rows_max, "flop_per_row": FLOP_PER_ROW_FWD, "avg_launch_ms": avg_ms,
                       "share_of_step": mlp_ms_per_step / step_ms, "traffic": None}
    # DRAM traffic per launch from the committed ncu --set full captures (profiles/r01d_*_ncu_summary.csv), per row x rows
    if args.precision == "bf16":
        if args.workload == "render" and rows_max == 3145728:
            roofline_tensor["traffic"] = NCU_BYTES_FWD_INFER_3145728 if (N_SAMPLES, N_IMPORTANCE) == (64, 128) else None
            roofline_tensor["traffic_source"] = "ncu dram__bytes_read+write, profiles/r01d_mlp_fwd_ncu_summary.csv"
        elif args.workload == "train":
            roofline_tensor["traffic"] = NCU_BYTES_PER_ROW_FWD_SAVE * rows_max
            roofline_tensor["traffic_source"] = ("ncu dram__bytes_read+write per row of the 196 608-row save-mode launch "
                                                 "(profiles/r01d_train_step_ncu_summary.csv) x rows of this launch")
    roofline = roofline_tensor
    if stage_acc is not None and args.precision == "bf16" and "mlp_bwd_wgrad" in stage_acc:
        wg_ms = stage_acc["mlp_bwd_wgrad"]
        gbs = rows_max * WGRAD_ALG_BYTES_PER_ROW / (wg_ms * 1e-3) / 1e9
        roofline = {"kernel": "wgrad_tc_kernel (split-K dW = dY^T X over sample rows, tcgen05, TMA-fed; + heads_wgrad beside it)",
                    "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                    "peak_source": pk["source"] + ", copy bandwidth (read+write)",
                    "rows_per_launch": rows_max, "bytes_per_row": WGRAD_ALG_BYTES_PER_ROW, "avg_launch_ms": wg_ms,
                    "share_of_step": wg_ms / step_ms,
                    "traffic": NCU_BYTES_PER_ROW_WGRAD * rows_max,
                    "traffic_source": "ncu dram__bytes_read+write per row (profiles/r01d_train_step_ncu_summary.csv) x rows"}

    line = {"metric": f"{args.workload}_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.workload == "train" else "strong",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args),
            "step_tflops": units_per_step * flop_per_unit / (ms_per_step * 1e-3) / 1e12 / world,
            "wall_s_timed_region": wall, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roofline}
    if roofline is not roofline_tensor:
        line["roofline_tensor"] = roofline_tensor
    line["density_only_coarse"] = density_only
    line["config"]["coarse_pass"] = "full network (as the reference); see density_only_coarse for the product default"

    if stage_acc is not None:
        line["stage_ms"] = {k_: round(v_, 5) for k_, v_ in stage_acc.items()}
        line["stage_ms_note"] = ("mean device time per stage over %d instrumented replays after the timed region "
                                 "(CUDA event records between the kernels of the replayed graph)" % K)
        line["config"]["step_api"] = "nerf_mlp_b200.TrainStep (" + ("CUDA graph replay" if train_step.use_graph else "eager launches") + ")"
    elif args.workload == "train":
        line["config"]["step_api"] = "NeRFRenderer._render_rays + loss.backward() + FlatAdam.step (autograd)"
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        rays_cpu = CPU_RAYS[args.workload] if args.cpu_port == "torch" else (256 if args.workload == "train" else 512)
        sec = time_cpu(args.workload, rays_cpu, 3, 1, args.cpu_port)
        line["cpu_baseline"] = {"value": rays_cpu / sec, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": cpu_sample_text(args.workload, rays_cpu, args.cpu_port, 3)}
    if rank == 0 and not args.no_cpu_baseline and world == 1 and (N_SAMPLES, N_IMPORTANCE) == (64, 128):
        try:
            line["torch_eager_gpu"] = time_torch_eager_gpu(args.workload, dev)
        except Exception as exc:  # a baseline must never take the bench line down
            line["torch_eager_gpu"] = {"unavailable": repr(exc)[:200]}
    if rank == 0:
        emit(line)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()

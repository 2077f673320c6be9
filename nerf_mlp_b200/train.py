"""One optimisation step of the reference's training loop, kept entirely on the device
(SURVEY.md section 8f row 1).

The reference's loop body (scripts/train.py:368-388)

    rgb_pred = renderer._render_rays(ray_o, ray_d)['rgb_map']          # :374
    loss = torch.mean((rgb_pred - target_rgb) ** 2)                    # :376
    batch_psnr = calculate_psnr(rgb_pred, target_rgb)                  # :379  (.cpu().numpy() -> skimage)
    optimizer.zero_grad(); loss.backward()                             # :381-382
    grad_norm = get_gradient_norm(model)                               # :385  (24x .item())
    optimizer.step(); scheduler.step()                                 # :387-388

runs here as a fixed sequence of libnerf_b200 launches with no autograd graph and no host sync:
loss, PSNR and the gradient norm are device scalars (``TrainStep.metrics``), the Adam step counter
and learning rate live in a device state block, and the whole sequence is captured ONCE as a CUDA
graph and replayed (``graph=True``), which removes the ~25 Python/ctypes launches per step that
otherwise dominate a 1024-ray step.  With ``graph=False`` the same sequence is enqueued eagerly
(used by the tests to show graph replay == eager, and == the autograd path of ops.RenderPassFn).

Data-parallel (world_size > 1): the step is two graphs with ONE flat NCCL all-reduce between them
(the only exchange step of the path).
"""
from __future__ import annotations

import math

import torch

from . import _lib, ops
from . import dist as dist_mod
from ._lib import check, dll, ptr, stream_ptr

_ST_LR, _ST_B1, _ST_B2, _ST_EPS, _ST_SCALE, _ST_STEP = range(6)
_ST_LOSS, _ST_PSNR, _ST_GNORM = 10, 11, 12


class TrainStep:
    """``step = TrainStep(renderer, optimizer, n_rays); loss = step(rays_o, rays_d, target)``.

    * ``renderer``  -- nerf_mlp_b200.NeRFRenderer (its knobs are read at construction/capture time;
      call :meth:`recapture` after changing them)
    * ``optimizer`` -- nerf_mlp_b200.FlatAdam on the same model; LR schedulers attached to it keep
      working: the current ``param_groups[0]['lr']`` is pushed to the device whenever it changes
    * ``n_rays``    -- fixed batch size (the graph is shape-static, like the reference's DataLoader
      with a fixed batch_size)

    Returns the loss as a 0-d device tensor (no sync).  ``metrics`` -> device tensor
    ``[loss, psnr, grad_norm]`` (float64), ``read_metrics()`` -> the same three as Python floats
    (one D2H copy + sync, for the periodic logging the reference does every step).
    """

    def __init__(self, renderer, optimizer, n_rays, *, graph=True, warmup=2, stage_events=False, split_graphs=None,
                 graph_allreduce=True, peer_allreduce=True):
        self.renderer, self.opt, self.model = renderer, optimizer, renderer.model
        if torch.device(renderer.device).type != "cuda":
            raise RuntimeError("nerf_mlp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if optimizer.model is not self.model:
            raise ValueError("TrainStep: optimizer and renderer must share one NeRFMLP")
        if renderer.N_importance <= 0:
            raise NotImplementedError("TrainStep implements the coarse+fine step (N_importance > 0)")
        if renderer.coarse_grad is True:
            raise NotImplementedError("TrainStep implements the reference's loss (fine rgb_map only, "
                                      "scripts/train.py:374-376), whose coarse-pass gradient is identically zero; "
                                      "an explicit coarse_grad=True (a coarse loss term) needs the autograd path")
        self.R = int(n_rays)
        self.dev = renderer.device
        m = self.model
        m._ensure_flat()
        # world_size > 1: gradient exchange + Adam as ONE kernel over NVLink peer memory (nerf_adam_step_fused_peer): the
        # flat gradient buffer is placed in symmetric memory so that every rank can read every other rank's.  Falls
        # back to NCCL's all-reduce (captured in the step graph) when symmetric memory cannot be set up.
        self._peer = None
        self.peer_allreduce_error = None
        if peer_allreduce and self._world() > 1 and not split_graphs:
            self._setup_peer()
        m._bind_flat_grads()
        optimizer._ensure_moments()
        f32 = dict(device=self.dev, dtype=torch.float32)
        self._inputs = torch.zeros((3, self.R, 3), **f32)     # static batch buffers of the captured step (one block)
        self.rays_o, self.rays_d, self.target = self._inputs[0], self._inputs[1], self._inputs[2]
        self.rays_d[:, 2] = -1.0
        self._pipe = None                                     # copy stream / staging of submit()
        self.state = torch.zeros(_lib.TRAIN_STATE_DOUBLES, device=self.dev, dtype=torch.float64)
        self._loss = torch.zeros((), **f32)
        self._scratch_c = torch.zeros(int(dll().nerf_composite_train_scratch_bytes(self.R)) // 8, device=self.dev, dtype=torch.float64)
        self._scratch_a = torch.zeros(int(dll().nerf_adam_fused_scratch_bytes(m.flat_params.numel())) // 8, device=self.dev,
                                      dtype=torch.float64)
        self._lr_pushed = None
        self._step_pushed = None
        self.use_graph = bool(graph)
        self._graphs = None
        self._keep = None
        self._warmup = int(warmup)
        self._ptrs = None
        # optional device timeline: one (external, graph-capturable) CUDA event after every stage
        self.stage_events = bool(stage_events)
        # two graphs (forward+backward | metrics+Adam+repack) with the eager all-reduce between them: always
        # at world_size > 1; split_graphs=True forces the same structure on one GPU (tests)
        self.split_graphs = split_graphs
        # world_size > 1: capture the NCCL all-reduce INSIDE the one step graph (no host round trip between two
        # replays: the exchange costs its device time only).  Falls back to two graphs with an eager all-reduce in
        # between if the capture is refused by the installed NCCL / PyTorch.
        self.graph_allreduce = bool(graph_allreduce)
        self.allreduce_in_graph = False
        self._marks = []
        self._push_state(force=True)
        if self.use_graph:
            self.recapture()

    # ---- device state block ----------------------------------------------------------------------
    def _world(self):
        return self.opt._world_size()

    def _push_state(self, force=False):
        g = self.opt.param_groups[0]
        lr = float(g["lr"])
        if force or self._step_pushed != self.opt._step:
            vals = [lr, float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), 1.0 / self._world(),
                    float(self.opt._step)]
            self.state[:6].copy_(torch.tensor(vals, dtype=torch.float64))
            self._lr_pushed, self._step_pushed = lr, self.opt._step
        elif lr != self._lr_pushed:
            self.state[_ST_LR:_ST_LR + 1].copy_(torch.tensor([lr], dtype=torch.float64))
            self._lr_pushed = lr

    # ---- the launch sequence -----------------------------------------------------------------------
    def _mark(self, name):
        if self.stage_events:
            # inside a capture the record must be an "external" event node (and only there)
            e = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
            e.record(torch.cuda.current_stream(self.dev))
            self._marks.append((name, e))

    def stage_times(self):
        """{stage: ms} of the last step (device time between consecutive stage events; needs
        stage_events=True and a synchronize by the caller)."""
        out = {}
        for (_, e0), (name, e1) in zip(self._marks[:-1], self._marks[1:]):
            if name != "start":
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    def _fwd_bwd(self):
        """render (coarse no-save, resample, fine save) -> MSE -> analytic backward into the flat
        gradient buffer.  RNG draws in the reference's order (renderer.py:60, :136, :182, :136)."""
        r, m, R = self.renderer, self.model, self.R
        prec = r._prec()
        o, d = self.rays_o, self.rays_d
        white, cs = bool(r.white_bkgd), float(r.coord_scale)
        self._mark("start")
        t_rand = torch.rand((R, r.N_samples), device=self.dev) if r.perturb > 0 else None
        z = ops.stratified_z(r._linspace(r.N_samples), t_rand, R, r.near, r.far)
        noise0 = (torch.randn(z.shape, device=self.dev) * r.raw_noise_std).contiguous() if r.raw_noise_std > 0. else None
        self._mark("stratified_z")
        # the coarse pass only feeds the resampling: densities alone (NERF_FWD_DENSITY_ONLY) unless the
        # renderer asks for the whole network (coarse_density_only=False)
        dens = bool(r.coarse_density_only)
        raw0, _ = ops.mlp_fwd_rays(m, o, d, z, cs, prec, False, density_only=dens)
        self._mark("mlp_fwd_coarse")
        rgb0, depth0, acc0, w0 = ops.composite_fwd(raw0, z, d, noise0, white, True)
        self._mark("composite_fwd_coarse")
        u = r._linspace(r.N_importance) if r.perturb == 0. else torch.rand((R, r.N_importance), device=self.dev)
        z_fine = ops.resample_merge(z, w0, u)
        noise1 = (torch.randn(z_fine.shape, device=self.dev) * r.raw_noise_std).contiguous() if r.raw_noise_std > 0. else None
        self._mark("resample_merge")
        raw1, ws = ops.mlp_fwd_rays(m, o, d, z_fine, cs, prec, True)
        self._mark("mlp_fwd_fine_save")
        # fine compositing + MSE + its gradient + compositing backward + zero_grad: ONE launch (nerf_composite_train)
        f32 = dict(device=self.dev, dtype=torch.float32)
        rgb, depth, acc = torch.empty((R, 3), **f32), torch.empty((R,), **f32), torch.empty((R,), **f32)
        d_raw = torch.empty((R, z_fine.shape[1], 4), **f32)
        d_rgb = None
        check(dll().nerf_composite_train(ptr(raw1), ptr(z_fine), ptr(d), ptr(noise1), R, z_fine.shape[1], int(white),
                                         ptr(self.target), ptr(rgb), ptr(depth), ptr(acc), ptr(d_raw), ptr(self._loss),
                                         ptr(self._scratch_c), ptr(m._flat_grad), m._flat_grad.numel(),
                                         stream_ptr(self.dev)), "nerf_composite_train")
        self._mark("composite_train(fwd+mse+bwd+zero_grad)")
        if self.stage_events and not (_lib.bwd_fused() and prec == _lib.PREC_BF16):   # two-kernel backward: one mark in between
            ops.mlp_bwd(m, d_raw, ws, prec, m._flat_grad, z_fine.shape[1], _lib.BWD_DGRAD)
            self._mark("mlp_bwd_dgrad")
            ops.mlp_bwd(m, d_raw, ws, prec, m._flat_grad, z_fine.shape[1], _lib.BWD_WGRAD)
            self._mark("mlp_bwd_wgrad")
        else:
            ops.mlp_bwd(m, d_raw, ws, prec, m._flat_grad, z_fine.shape[1])       # fused dgrad + wgrad launch (+ heads)
            self._mark("mlp_bwd")
        self.outputs = {"rgb_map": rgb, "depth_map": depth, "acc_map": acc}
        if not dens:
            self.outputs.update({"rgb_map_coarse": rgb0, "depth_map_coarse": depth0, "acc_map_coarse": acc0})
        return (t_rand, z, noise0, raw0, w0, u, z_fine, noise1, raw1, ws, d_rgb, d_raw)

    def _update(self):
        """metrics + Adam (device-side scalars) + bf16 weight re-pack."""
        m, opt = self.model, self.opt
        n = m.flat_params.numel()
        st = stream_ptr(self.dev)
        if self._world() > 1 and self._peer is None:
            self._mark("grad_allreduce")
        if self._peer is not None:
            # gradient exchange (one-shot all-reduce over peer memory, summed in rank order) + metrics + Adam: ONE launch
            pr = self._peer
            check(dll().nerf_adam_step_fused_peer(ptr(m.flat_params), pr["grads"], pr["red"], pr["flags"], pr["rank"], pr["world"],
                                                  ptr(opt._m), ptr(opt._v), n, ptr(self.state), ptr(self._loss),
                                                  ptr(self._scratch_a), st), "nerf_adam_step_fused_peer")
        else:
            # metrics (loss, PSNR, grad norm) + step counter + Adam: ONE launch (nerf_adam_step_fused)
            check(dll().nerf_adam_step_fused(ptr(m.flat_params), ptr(m._flat_grad), ptr(opt._m), ptr(opt._v), n,
                                             ptr(self.state), ptr(self._loss), ptr(self._scratch_a), st), "nerf_adam_step_fused")
        if r_is_bf16(self.renderer):
            check(dll().nerf_pack_weights(ptr(m.flat_params), ptr(m._packed), st), "nerf_pack_weights")
        self._mark("metrics+adam+repack")

    def _setup_peer(self):
        """Place the flat gradient buffer in symmetric memory and exchange the peers' pointers.  On any failure the
        step keeps NCCL's all-reduce (the reason is kept in ``peer_allreduce_error``)."""
        import ctypes
        m, dist = self.model, torch.distributed
        try:
            import torch.distributed._symmetric_memory as symm
            world = self._world()
            if not (dist.is_available() and dist.is_initialized()) or world > _lib.PEER_MAX:
                raise RuntimeError("no initialised process group" if world <= _lib.PEER_MAX else f"world {world} > {_lib.PEER_MAX}")
            group = self.opt.process_group if self.opt.process_group is not None else dist.group.WORLD
            n = m.flat_params.numel()
            lay = dist_mod.peer_exchange_layout(n, world, peer_max=_lib.PEER_MAX)
            two_shot = lay["two_shot"]
            buf = symm.empty(lay["floats"], dtype=torch.float32, device=self.dev)
            buf.zero_()
            hdl = symm.rendezvous(buf, group)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            if len(ptrs) != world or ptrs[hdl.rank] != buf.data_ptr():
                raise RuntimeError("symmetric memory handle does not describe this buffer")
            torch.cuda.synchronize(self.dev)
            dist.barrier(group)                                   # every rank's flags are zero before anybody's first step
            arr = ctypes.c_void_p * world
            self._peer = {"buf": buf, "hdl": hdl, "rank": int(hdl.rank), "world": world, "n": n, "two_shot": two_shot,
                          "grads": arr(*ptrs), "red": arr(*[q + 4 * lay["red_off"] for q in ptrs]) if two_shot else None,
                          "flags": arr(*[q + 4 * lay["flag_off"] for q in ptrs])}
            for prm in m._param_list:                             # re-bind the parameters' .grad views to the new buffer
                prm.grad = None
            m._flat_grad = buf[:n]
        except Exception as exc:                                  # noqa: BLE001 -- any failure means "use NCCL"
            self._peer = None
            self.peer_allreduce_error = f"{type(exc).__name__}: {exc}"
        # the choice must be the same on every rank (a rank on NCCL and a rank in the peer kernel would wait forever)
        if dist.is_available() and dist.is_initialized():
            ok = torch.tensor([1 if self._peer is not None else 0], device=self.dev, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.opt.process_group)
            if int(ok.item()) == 0 and self._peer is not None:
                self._peer = None
                self.peer_allreduce_error = "another rank could not set up symmetric memory"
                for prm in m._param_list:
                    prm.grad = None
                m._flat_grad = None

    def _allreduce(self):
        if self._peer is not None:
            return                                                # done inside nerf_adam_step_fused_peer
        if self._world() > 1:
            torch.distributed.all_reduce(self.model._flat_grad, op=torch.distributed.ReduceOp.SUM,
                                         group=self.opt.process_group)

    def _eager(self):
        self._marks = []
        keep = self._fwd_bwd()
        self._allreduce()
        self._update()
        return keep

    # ---- capture -------------------------------------------------------------------------------------
    def _live_ptrs(self):
        m, opt = self.model, self.opt
        m._ensure_flat()
        m._bind_flat_grads()
        opt._ensure_moments()                                   # a checkpoint saved before the first step loads as None
        return (m.flat_params.data_ptr(), m._flat_grad.data_ptr(), opt._m.data_ptr(), opt._v.data_ptr(),
                m._packed.data_ptr() if m._packed is not None else 0)

    def recapture(self):
        """(Re)capture the step.  Warm-up steps run eagerly on a side stream first (lazy library
        initialisation, allocator warm-up); parameters, Adam moments and the RNG state are restored
        afterwards, so constructing a TrainStep does not train."""
        m, opt = self.model, self.opt
        if self._peer is not None and (m._flat_grad is None or m._flat_grad.data_ptr() != self._peer["buf"].data_ptr()):
            raise RuntimeError("TrainStep: the model's flat gradient buffer was re-created after the peer-memory "
                               "gradient exchange was set up; construct a new TrainStep (or pass peer_allreduce=False)")
        if r_is_bf16(self.renderer):
            m.packed_weights()                                        # clean packed image before capture
        opt._ensure_moments()
        snap = (m.flat_params.clone(), opt._m.clone(), opt._v.clone(), self.state.clone())
        rng = torch.cuda.get_rng_state(self.dev)
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(self._warmup):
                self._eager()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._graphs, self._keep, self._marks = [], [], []
        pool = None
        launches0 = dll().nerf_launch_count()
        with torch.no_grad():
            multi = self._world() > 1
            split = multi if self.split_graphs is None else (self.split_graphs or multi)
            self.allreduce_in_graph = False
            if multi and (self.graph_allreduce or self._peer is not None) and not self.split_graphs:
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._keep.append(self._fwd_bwd())
                        self._allreduce()
                        self._keep.append(self._update())
                    self._graphs.append(g)
                    self.allreduce_in_graph = True
                except Exception:                                     # capture of the collective refused: two graphs
                    torch.cuda.synchronize(self.dev)
                    self._graphs, self._keep, self._marks = [], [], []
            if not self._graphs:
                parts = [(self._fwd_bwd,), (self._update,)] if split else [(self._fwd_bwd, self._update)]
                for fns in parts:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        for fn in fns:
                            self._keep.append(fn())
                    pool = g.pool()
                    self._graphs.append(g)
        launches_after = dll().nerf_launch_count()
        torch.cuda.synchronize(self.dev)
        with torch.no_grad():
            m.flat_params.copy_(snap[0]); opt._m.copy_(snap[1]); opt._v.copy_(snap[2]); self.state.copy_(snap[3])
            if r_is_bf16(self.renderer):
                m.mark_dirty()
                m.packed_weights()
        torch.cuda.set_rng_state(rng, self.dev)
        self._ptrs = self._live_ptrs()
        self.launches_per_step = int(launches_after - launches0)      # libnerf_b200 kernels in one replay

    # ---- public ----------------------------------------------------------------------------------------
    def load_batch(self, rays_o, rays_d, target):
        """Copy one batch into the step's static buffers (device or pinned-host sources, async)."""
        for dst, src in ((self.rays_o, rays_o), (self.rays_d, rays_d), (self.target, target)):
            if src is not dst:
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"TrainStep: batch shape {tuple(src.shape)} != captured shape {tuple(dst.shape)}")
                dst.copy_(src, non_blocking=True)

    def __call__(self, rays_o=None, rays_d=None, target=None):
        if rays_o is not None:
            self.load_batch(rays_o, rays_d, target)
        self._push_state()
        m = self.model
        if self.use_graph and self._ptrs != self._live_ptrs():
            self.recapture()                                      # model moved / buffers re-created since capture
        if r_is_bf16(self.renderer):
            m.packed_weights()                                    # host-side check; re-packs only after an external weight load
        with torch.no_grad():
            if self.use_graph:
                self._graphs[0].replay()
                if len(self._graphs) > 1:
                    self._allreduce()
                    self._graphs[1].replay()
            else:
                self._keep = self._eager()
        # host mirrors: FlatAdam's counter, the LR scheduler's "optimizer.step() was called" flag
        self.opt._step += 1
        self._step_pushed = self.opt._step
        self.opt._opt_called = True
        return self._loss

    # ---- pipelined interface: H2D of the next batch and the metric read-back overlap the running step ----
    def _pipeline(self):
        if self._pipe is None:
            n_slots = 4
            self._pipe = {
                "copy_stream": torch.cuda.Stream(self.dev),
                "stage": torch.zeros_like(self._inputs),
                "ev_h2d": torch.cuda.Event(), "ev_stage_free": torch.cuda.Event(),
                "host": [torch.zeros(3, dtype=torch.float64).pin_memory() for _ in range(n_slots)],
                "done": [torch.cuda.Event() for _ in range(n_slots)],
                "slot": 0,
            }
        return self._pipe

    def submit(self, rays_o, rays_d, target):
        """Enqueue one step on a batch held in (pinned) HOST memory without waiting for anything:
        the three H2D copies run on a copy stream while the previous step is still computing (they land in a
        staging block; one 36 KB device copy moves it into the step's static buffers), the step is replayed,
        and [loss, psnr, grad_norm] are copied to a pinned slot.  Returns a ticket for :meth:`result`.  Every
        step still moves its own inputs host->device and its own metrics device->host -- only the waiting is
        taken off the critical path (the reference blocks on .item() / .cpu() several times per step)."""
        if rays_o.is_cuda or rays_d.is_cuda or target.is_cuda:
            self(rays_o, rays_d, target)                      # device-resident batch: nothing to overlap
            pipe = self._pipeline()
        else:
            pipe = self._pipeline()
            cs, stage = pipe["copy_stream"], pipe["stage"]
            for src in (rays_o, rays_d, target):
                if tuple(src.shape) != (self.R, 3):
                    raise RuntimeError(f"TrainStep: batch shape {tuple(src.shape)} != captured shape {(self.R, 3)}")
            main = torch.cuda.current_stream(self.dev)
            cs.wait_event(pipe["ev_stage_free"])              # the previous step has taken its batch out of the stage
            with torch.cuda.stream(cs):
                stage[0].copy_(rays_o, non_blocking=True)
                stage[1].copy_(rays_d, non_blocking=True)
                stage[2].copy_(target, non_blocking=True)
                pipe["ev_h2d"].record(cs)
            main.wait_event(pipe["ev_h2d"])
            self._inputs.copy_(stage, non_blocking=True)
            pipe["ev_stage_free"].record(main)
            self()
        slot = pipe["slot"]
        pipe["slot"] = (slot + 1) % len(pipe["host"])
        pipe["host"][slot].copy_(self.state[_ST_LOSS:_ST_GNORM + 1], non_blocking=True)
        pipe["done"][slot].record(torch.cuda.current_stream(self.dev))
        return slot

    def result(self, ticket):
        """Metrics of the step submitted under `ticket` (blocks until that step has finished; at most
        len(slots)-1 newer steps may have been submitted since)."""
        pipe = self._pipeline()
        pipe["done"][ticket].synchronize()
        loss, psnr, gnorm = pipe["host"][ticket].tolist()
        return {"loss": loss, "psnr": psnr, "grad_norm": gnorm}

    @property
    def loss(self):
        return self._loss

    @property
    def metrics(self):
        """device tensor [loss, psnr, grad_norm] (float64) of the last step; no sync."""
        return self.state[_ST_LOSS:_ST_GNORM + 1]

    def read_metrics(self):
        loss, psnr, gnorm = self.metrics.tolist()
        return {"loss": loss, "psnr": psnr, "grad_norm": gnorm}


def r_is_bf16(renderer):
    return renderer._prec() == _lib.PREC_BF16


def psnr_from_mse(mse):
    """skimage.metrics.peak_signal_noise_ratio(data_range=1.0) given the mse (scripts/train.py:33-37)."""
    return 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")

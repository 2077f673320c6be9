"""TrainStep: the reference's loop body scripts/train.py:374-388 as a fixed
launch sequence / CUDA graph.  Parity bars:

  * eager TrainStep vs the autograd path (NeRFRenderer._render_rays + loss.backward() + FlatAdam,
    i.e. the reference's own loop on the drop-in classes): same kernels in the same order, so the
    loss of step 1 is bit-identical and the gradient of step 1 agrees to 1e-5 relative L2 (the
    wgrad kernel's fp32 atomics may reorder); losses of 3 steps agree to 2e-4 relative;
    parameters after 3 steps: >= 85 % within 2e-5 and all within 2.1e-3 (= the bars of
    test_optimizer_steps_match_reference: Adam's m/sqrt(v) turns a reordered-atomics difference
    in a near-zero gradient into an O(lr) difference in that weight);
  * graph replay vs eager TrainStep: same bars, same RNG stream (perturb=1 draws included);
  * fp32 check mode vs the reference's golden training run (tests/golden/train_r32.npz): the same
    bars as test_optimizer_steps_match_reference;
  * metrics: loss == mean((rgb-target)^2), psnr == 10 log10(1/loss) (skimage, data_range 1),
    grad_norm == sqrt(sum_p |p.grad|^2) (scripts/train.py:60-67) to 1e-6 relative;
  * the learning-rate schedule (StepLR, scripts/train.py:260) advances on the device state block.
"""
import math

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, torch.float32)


@pytest.fixture(scope="module")
def nb():
    import nerf_mlp_b200
    return nerf_mlp_b200


def make(nb, seed, precision, perturb, lr=5e-4):
    m = nb.NeRFMLP(precision=precision)
    p = O.init_params(seed)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    m = m.to(DEV)
    r = nb.NeRFRenderer(m, DEV, perturb=perturb)
    return m, r, nb.FlatAdam(m, lr=lr)


def batch(R, seed):
    o, d = O.random_rays(R, seed)
    tgt = np.random.default_rng(seed + 1).uniform(0, 1, (R, 3)).astype(np.float32)
    return T(o), T(d), T(tgt)


def autograd_steps(nb, m, r, opt, o, d, tgt, n, sched=None, grads=None):
    losses = []
    for _ in range(n):
        loss = nb.ops.mse_loss(r._render_rays(o, d)["rgb_map"], tgt)
        opt.zero_grad()
        loss.backward()
        if grads is not None and not grads:
            grads.append(m.flat_grad.detach().cpu().numpy().copy())
        opt.step()
        if sched is not None:
            sched.step()
        losses.append(float(loss.detach()))
    return losses


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("perturb", [0.0, 1.0])
def test_eager_and_graph_match_autograd_path(nb, precision, perturb):
    R, n = 200, 3
    o, d, tgt = batch(R, 7)
    runs = {}
    for kind in ("autograd", "eager", "graph"):
        m, r, opt = make(nb, 11, precision, perturb)
        g1 = []
        if kind == "autograd":
            torch.manual_seed(123)
            losses = autograd_steps(nb, m, r, opt, o, d, tgt, n, grads=g1)
        else:
            step = nb.TrainStep(r, opt, R, graph=(kind == "graph"))
            torch.manual_seed(123)
            losses = []
            for i in range(n):
                losses.append(float(step(o, d, tgt)))
                if i == 0:
                    g1.append(m.flat_grad.detach().cpu().numpy().copy())
            assert opt._step == n
        runs[kind] = (losses, m.flat_params.detach().cpu().numpy().copy(), g1[0])
    ref_l, ref_p, ref_g = runs["autograd"]
    assert ref_l[-1] < ref_l[0]
    for kind in ("eager", "graph"):
        l, p, g = runs[kind]
        assert l[0] == ref_l[0], (kind, l, ref_l)                    # forward path is deterministic
        assert np.linalg.norm(g - ref_g) <= 1e-5 * np.linalg.norm(ref_g), kind
        np.testing.assert_allclose(l, ref_l, rtol=2e-4, err_msg=kind)
        np.testing.assert_allclose(p, ref_p, atol=2.1e-3, err_msg=kind)
        assert np.mean(np.abs(p - ref_p) <= 2e-5) > 0.85, kind


def test_renderer_knobs_inside_the_graph(nb):
    """raw_noise_std > 0 (torch.randn draws captured in the graph, reference order renderer.py:60,136,182,136),
    black background, coord_scale != 1, non-default near/far and sample counts: graph replay == autograd path."""
    R, n = 64, 2
    o, d, tgt = batch(R, 17)
    res = {}
    for kind in ("autograd", "graph"):
        m, _, opt = make(nb, 11, "bf16", 1.0)
        r = nb.NeRFRenderer(m, DEV, N_samples=32, N_importance=48, near=1.5, far=5.0, white_bkgd=False, perturb=1.0,
                            raw_noise_std=0.5, coord_scale=0.7)
        torch.manual_seed(9)
        if kind == "autograd":
            losses = autograd_steps(nb, m, r, opt, o, d, tgt, n)
        else:
            step = nb.TrainStep(r, opt, R)
            torch.manual_seed(9)
            losses = [float(step(o, d, tgt)) for _ in range(n)]
        res[kind] = losses
    assert res["graph"][0] == res["autograd"][0]
    np.testing.assert_allclose(res["graph"], res["autograd"], rtol=2e-4)


def test_construction_does_not_train(nb):
    m, r, opt = make(nb, 3, "bf16", 1.0)
    before = m.flat_params.clone()
    torch.manual_seed(5)
    a = torch.rand(4, device=DEV)
    torch.manual_seed(5)
    step = nb.TrainStep(r, opt, 64)
    b = torch.rand(4, device=DEV)
    assert torch.equal(before, m.flat_params) and torch.equal(a, b)
    assert opt._step == 0 and float(step.state[5]) == 0.0
    assert float(opt._m.abs().max()) == 0.0


def test_fp32_step_matches_reference_golden(nb):
    """Two graph-replayed steps reproduce the reference's own training run (same bars as
    test_optimizer_steps_match_reference in test_gpu_mlp_render.py)."""
    g = load_golden("train_r32")
    m, r, opt = make(nb, int(g["seed"]), "fp32", 0.0)
    R = g["rays_o"].shape[0]
    step = nb.TrainStep(r, opt, R)
    o, d, tgt = T(g["rays_o"]), T(g["rays_d"]), T(g["target"])
    for i in (1, 2):
        loss = float(step(o, d, tgt))
        assert abs(loss - float(g[f"loss_step{i - 1}"])) < 1e-4      # golden steps are 0-based
        flat = m.flat_params.detach().cpu().numpy()
        ref = g[f"params_after_step{i}_sub"]
        np.testing.assert_allclose(flat[::101], ref, atol=2.1e-3)
        assert np.mean(np.abs(flat[::101] - ref) < 2e-5) > 0.97


def test_metrics_and_lr_schedule(nb):
    R = 128
    o, d, tgt = batch(R, 21)
    m, r, opt = make(nb, 11, "bf16", 0.0, lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.1)
    step = nb.TrainStep(r, opt, R)
    m2, r2, opt2 = make(nb, 11, "bf16", 0.0, lr=1e-3)
    sched2 = torch.optim.lr_scheduler.StepLR(opt2, step_size=2, gamma=0.1)
    for i in range(5):
        loss = step(o, d, tgt)
        sched.step()
        got = step.read_metrics()
        rgb = step.outputs["rgb_map"]
        mse = float(torch.mean((rgb.double() - tgt.double()) ** 2))
        assert abs(got["loss"] - float(loss)) == 0.0
        assert abs(got["loss"] - mse) <= 1e-6 * mse
        assert abs(got["psnr"] - 10.0 * math.log10(1.0 / mse)) < 1e-4
        gn = math.sqrt(sum(float(p.grad.norm(2)) ** 2 for p in m.parameters()))      # scripts/train.py:60-67
        assert abs(got["grad_norm"] - gn) <= 1e-5 * gn
        assert float(step.state[5]) == i + 1
        autograd_steps(nb, m2, r2, opt2, o, d, tgt, 1, sched2)
        assert float(step.state[0]) == pytest.approx(1e-3 * 0.1 ** (i // 2), rel=1e-12)   # lr used by step i
        pa, pb = m.flat_params.cpu().numpy(), m2.flat_params.cpu().numpy()
        np.testing.assert_allclose(pa, pb, atol=4.2e-3)
        assert np.mean(np.abs(pa - pb) <= 2e-5) > 0.85
    assert opt.param_groups[0]["lr"] == pytest.approx(opt2.param_groups[0]["lr"])


def test_external_weight_load_is_seen_by_the_graph(nb):
    R = 64
    o, d, tgt = batch(R, 31)
    m, r, opt = make(nb, 11, "bf16", 0.0)
    step = nb.TrainStep(r, opt, R)
    l0 = float(step(o, d, tgt))
    p = O.init_params(12)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    m3, r3, opt3 = make(nb, 12, "bf16", 0.0)
    ref = float(nb.ops.mse_loss(r3._render_rays(o, d)["rgb_map"], tgt))
    l1 = float(step(o, d, tgt))
    assert l1 == ref and l1 != l0


def test_data_parallel_graph_structure(nb):
    """The world_size > 1 structure (two graphs with the gradient exchange between them, 1/world folded
    into the Adam kernel) on one GPU: identical to the single-graph step (bit-identical loss, same bars
    on the parameters); with world_size forced to 2 and the all-reduce stubbed by a doubling (two ranks
    with the same batch), Adam sees the same averaged gradient."""
    R = 96
    o, d, tgt = batch(R, 41)
    out = {}
    for kind in ("single", "split", "world2", "world2_one_graph"):
        m, r, opt = make(nb, 11, "bf16", 0.0)
        cls = nb.TrainStep
        if kind.startswith("world2"):
            opt._world = 2

            class TwoIdenticalRanks(nb.TrainStep):
                def _allreduce(self):                               # sum over two ranks holding the same batch
                    self.model._flat_grad.mul_(2.0)
            cls = TwoIdenticalRanks
        # world2_one_graph: the default at world_size > 1 -- the exchange is captured inside the single step graph
        step = cls(r, opt, R, split_graphs=(None if kind == "world2_one_graph" else kind != "single"))
        if kind.startswith("world2"):
            assert float(step.state[4]) == 0.5
        losses = [float(step(o, d, tgt)) for _ in range(3)]
        assert len(step._graphs) == (1 if kind in ("single", "world2_one_graph") else 2)
        assert step.allreduce_in_graph == (kind == "world2_one_graph")
        out[kind] = (losses, m.flat_params.detach().cpu().numpy().copy(), step.read_metrics()["grad_norm"])
    for kind in ("split", "world2", "world2_one_graph"):
        assert out[kind][0][0] == out["single"][0][0]
        np.testing.assert_allclose(out[kind][0], out["single"][0], rtol=2e-4)
        np.testing.assert_allclose(out[kind][1], out["single"][1], atol=2.1e-3)
        assert np.mean(np.abs(out[kind][1] - out["single"][1]) <= 2e-5) > 0.85
        assert abs(out[kind][2] - out["single"][2]) <= 1e-3 * out["single"][2]


def test_shape_and_config_errors(nb):
    m, r, opt = make(nb, 3, "bf16", 1.0)
    step = nb.TrainStep(r, opt, 32, graph=False)
    o, d, tgt = batch(16, 1)
    with pytest.raises(RuntimeError):
        step(o, d, tgt)
    r.coarse_grad = True
    with pytest.raises(NotImplementedError):
        nb.TrainStep(r, opt, 32)


def test_pipelined_submit_result(nb):
    """submit()/result(): batches in pinned host memory, H2D on a copy stream, metrics read one step late --
    the same training trajectory as the blocking call on the same batches."""
    R, n = 96, 6
    batches = []
    for k in range(n):
        o, d = O.random_rays(R, 50 + k)
        t = np.random.default_rng(60 + k).uniform(0, 1, (R, 3)).astype(np.float32)
        batches.append(tuple(torch.from_numpy(a).pin_memory() for a in (o, d, t)))
    res = {}
    for kind in ("blocking", "pipelined"):
        m, r, opt = make(nb, 11, "bf16", 0.0)
        step = nb.TrainStep(r, opt, R)
        losses = []
        if kind == "blocking":
            for b in batches:
                step(*b)
                losses.append(step.read_metrics()["loss"])
        else:
            ticket = None
            for b in batches:
                t_new = step.submit(*b)
                if ticket is not None:
                    losses.append(step.result(ticket)["loss"])
                ticket = t_new
            last = step.result(ticket)
            losses.append(last["loss"])
            assert np.isfinite(last["psnr"]) and last["grad_norm"] > 0
        assert opt._step == n
        res[kind] = (losses, m.flat_params.detach().cpu().numpy().copy())
    assert res["pipelined"][0][0] == res["blocking"][0][0]
    np.testing.assert_allclose(res["pipelined"][0], res["blocking"][0], rtol=2e-4)
    np.testing.assert_allclose(res["pipelined"][1], res["blocking"][1], atol=4.2e-3)
    # (six Adam steps on six different batches: the reordered-atomics noise of near-zero gradients has had more
    # steps to spread than in the 3-step tests above)
    assert np.mean(np.abs(res["pipelined"][1] - res["blocking"][1]) <= 2e-5) > 0.6
    # a device-resident batch goes through the same interface
    m, r, opt = make(nb, 11, "bf16", 0.0)
    step = nb.TrainStep(r, opt, R)
    tk = step.submit(*(b.to(DEV) for b in batches[0]))
    assert step.result(tk)["loss"] == res["blocking"][0][0]


def test_peer_gradient_exchange_kernel_two_ranks_on_one_gpu(nb):
    """nerf_adam_step_fused_peer (gradient exchange over peer memory + Adam in one kernel): two emulated ranks on ONE
    GPU -- their kernels run concurrently on two streams and hand-shake through the flag words exactly as two
    processes do over NVLink -- must both produce, bit for bit, what nerf_adam_step_fused produces on the summed
    gradient with grad_scale = 1/2; three steps, so that the epoch counter and the "read done" hand-shake are
    exercised.  world = 1 degenerates to nerf_adam_step_fused itself.  Both exchange schemes: one-shot (every rank reads
    all gradients) and two-shot (reduce-scatter + all-gather through the `red` buffers; also three ranks, whose slices
    do not divide the buffer evenly).  (Real ranks: tests/test_gpu_multi.py.)"""
    import ctypes
    from nerf_mlp_b200 import _lib
    dll, ptr = _lib.dll(), _lib.ptr
    n = 10_003                                                   # not a multiple of 4: the scalar tail is exercised
    n_pad = (n + 127) // 128 * 128
    n_flags = 2 * _lib.PEER_MAX + 32
    g = torch.Generator(DEV).manual_seed(3)
    f32 = dict(device=DEV, dtype=torch.float32)

    def state(world):
        s = torch.zeros(_lib.TRAIN_STATE_DOUBLES, device=DEV, dtype=torch.float64)
        s[:6] = torch.tensor([5e-4, 0.9, 0.999, 1e-8, 1.0 / world, 0.0], dtype=torch.float64)
        return s

    def scratch():
        return torch.zeros(int(dll.nerf_adam_fused_scratch_bytes(n)) // 8, device=DEV, dtype=torch.float64)

    for world, two_shot in ((1, False), (2, False), (1, True), (2, True), (3, True)):
        p_init = torch.randn(n, generator=g, **f32)
        bufs = [torch.zeros(2 * n_pad + n_flags, **f32) for _ in range(world)]     # gradient | reduced gradient | flags
        arr = ctypes.c_void_p * world
        grads_arr = arr(*[b.data_ptr() for b in bufs])
        red_arr = arr(*[b.data_ptr() + 4 * n_pad for b in bufs]) if two_shot else None
        flags_arr = arr(*[b.data_ptr() + 8 * n_pad for b in bufs])
        ranks = [dict(p=p_init.clone(), m=torch.zeros(n, **f32), v=torch.zeros(n, **f32), st=state(world), sc=scratch(),
                      stream=torch.cuda.Stream()) for _ in range(world)]
        ref = dict(p=p_init.clone(), m=torch.zeros(n, **f32), v=torch.zeros(n, **f32), st=state(world), sc=scratch())
        loss = torch.tensor(0.25, **f32)
        for step in range(3):
            local = [torch.randn(n, generator=g, **f32) * 1e-2 for _ in range(world)]
            for b, lg in zip(bufs, local):
                b[:n].copy_(lg)
            total = local[0].clone()
            for lg in local[1:]:
                total = total + lg                                # rank order, fp32: what the kernel computes
            torch.cuda.synchronize()
            for r, rk in enumerate(ranks):
                with torch.cuda.stream(rk["stream"]):
                    _lib.check(dll.nerf_adam_step_fused_peer(ptr(rk["p"]), grads_arr, red_arr, flags_arr, r, world, ptr(rk["m"]), ptr(rk["v"]),
                                                             n, ptr(rk["st"]), ptr(loss), ptr(rk["sc"]),
                                                             ctypes.c_void_p(rk["stream"].cuda_stream)), "peer")
            _lib.check(dll.nerf_adam_step_fused(ptr(ref["p"]), ptr(total), ptr(ref["m"]), ptr(ref["v"]), n, ptr(ref["st"]),
                                                ptr(loss), ptr(ref["sc"]), _lib.stream_ptr(DEV)), "fused")
            torch.cuda.synchronize()
            for rk in ranks:
                for k in ("p", "m", "v"):
                    assert torch.equal(rk[k], ref[k]), (world, two_shot, step, k)
                assert torch.equal(rk["st"][:12], ref["st"][:12]), (world, two_shot, step)     # scalars, step counter, loss, psnr
                # grad norm: fp64 sum of squares, partial sums grouped per float4 here and per element there
                assert abs(float(rk["st"][12]) - float(ref["st"][12])) <= 1e-12 * float(ref["st"][12])
            for b, lg in zip(bufs, local):
                assert torch.equal(b[:n], lg)                     # the gradient buffers keep the rank-local gradients
                flags = b[2 * n_pad:].view(torch.int32)
                assert int(flags[2 * _lib.PEER_MAX]) == step + 1  # epoch

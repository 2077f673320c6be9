"""The two stated relaxations of the parity gates, pinned with data.  Needs a B200: `pytest -m gpu`.

(a) fp32 check mode, END TO END through the inverse cdf.  north_star's 1e-4 gate holds stage by stage (the oracle
    continues from the kernels' own z_fine, tests/test_gpu_mlp_render.py); end to end the rule is ">= 70 % of the rays
    within 1e-4, none above 5e-3", justified by the conditioning of renderer.py:190-198 (pdf bins ~1e-4 wide at random
    init; `denom < 1e-5 -> 1` at :195).  Here that justification is MEASURED: the reference's own arithmetic
    (oracle/nerf_oracle_torch.py, bit-identical to the reference on CPU) run as eager fp32 PyTorch on this GPU differs
    from the reference's CPU image by a spread S; the kernels' end-to-end deviation from eager CUDA must stay within that
    spread (same quantiles, factor 2), and both are printed.
(b) bf16 gradients.  The tight bar (2e-2) is against the backward on the bf16 forward's own ReLU pattern; against the
    all-fp32 gradient the whole-gradient bar was 0.25 relative L2.  Added: per-tensor cosine similarity >= 0.99 (a wrong
    layer cannot pass), and a 120-step loss trajectory (perturb = 1, 1024 rays, Adam 5e-4) against the reference's loop
    on eager CUDA (torch autograd + torch.optim.Adam) with the same random stream.
(c) the coarse maps are differentiable by default, as in the reference: a coarse loss term produces a
    coarse-pass gradient through the lazily recomputed pass (ops.RenderPassFn save="lazy").
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import nerf_oracle_torch as TP
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def nb():
    import nerf_mlp_b200
    assert not torch.backends.cuda.matmul.allow_tf32
    return nerf_mlp_b200


def _model(nb, seed, precision):
    p = O.init_params(seed)
    m = nb.NeRFMLP(precision=precision)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    return p, m.to(DEV)


def _q(x):
    return dict(max=float(x.max()), p99=float(np.quantile(x, 0.99)), p90=float(np.quantile(x, 0.90)),
                frac_le_1e4=float((x <= 1e-4).mean()))


def test_fp32_e2e_gate_on_the_reference_view(nb):
    """BASELINE configs[0] (100x100 pinhole view, 64+128, perturb 0): END TO END -- through the inverse cdf -- the fp32
    check mode meets north_star's 1e-4 gate on EVERY pixel against the reference's own CPU image, and sits inside the
    spread between the reference's arithmetic on CPU and on this GPU (eager PyTorch)."""
    g = load_golden("render_pinhole_100x100")
    ref_cpu = g["image"].reshape(-1, 3)                                  # the reference's own CPU render()
    o, d, focal = O.pinhole_rays(100, 100)
    p, m32 = _model(nb, int(g["seed"]), "fp32")
    to, td_ = torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV)
    with torch.no_grad():
        eager = TP.render_rays(TP.params_from_numpy(p, device=DEV), to, td_, perturb=0.0)["rgb_map"].cpu().numpy()
        ours = nb.NeRFRenderer(m32, DEV, perturb=0.0).render(to, td_, 100, 100, focal).reshape(-1, 3).cpu().numpy()
    spread = np.abs(eager - ref_cpu).max(-1)                             # same arithmetic, CPU vs CUDA kernels of torch
    dev_k = np.abs(ours - eager).max(-1)                                 # our kernels vs the same arithmetic on the same GPU
    dev_cpu = np.abs(ours - ref_cpu).max(-1)
    print("reference CPU vs reference-arithmetic eager CUDA (the spread):", _q(spread))
    print("kernels (fp32 check mode) vs eager CUDA:                      ", _q(dev_k))
    print("kernels (fp32 check mode) vs reference CPU image:             ", _q(dev_cpu))
    assert dev_cpu.max() <= 1e-4 and dev_k.max() <= 1e-4                 # the strict gate, end to end
    assert dev_cpu.max() <= 4.0 * max(spread.max(), 2.5e-7)              # and no further from the reference than torch's own two backends


def test_fp32_e2e_relaxed_rule_is_the_references_own_spread_on_random_rays(nb):
    """The golden cases on i.i.d. rays (O.random_rays: origins ~ N((0,0,4), 0.1^2), non-unit directions) are where the
    end-to-end rule had to be relaxed to ">= 70 % within 1e-4, none above 5e-3" (tests/test_gpu_mlp_render.py): many
    coarse weights are exactly 0 there, pdf bins are ~1e-4 wide and a 1-ulp cdf difference moves a sample by ~1e-3 in z.
    Measured here on 4096 such rays: the spread between the reference's arithmetic on CPU and on this GPU, and the
    kernels' deviation from the GPU run -- the kernels must not be further from either than the two backends of the
    reference are from each other (same quantiles, factor 2)."""
    R = 4096
    p, m32 = _model(nb, 31, "fp32")
    o, d = O.random_rays(R, 32)
    to, td_ = torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV)
    with torch.no_grad():
        cpu = TP.render_rays(TP.params_from_numpy(p), torch.from_numpy(o), torch.from_numpy(d), perturb=0.0)
        eager = TP.render_rays(TP.params_from_numpy(p, device=DEV), to, td_, perturb=0.0)
        ours = nb.NeRFRenderer(m32, DEV, perturb=0.0)._render_rays(to, td_)
    for k, span in (("rgb_map", 1.0), ("acc_map", 1.0), ("depth_map", 4.0)):
        c, e, u = (x[k].detach().cpu().numpy().reshape(R, -1) for x in (cpu, eager, ours))
        spread = np.abs(e - c).max(-1) / span
        dev_k = np.abs(u - e).max(-1) / span
        dev_c = np.abs(u - c).max(-1) / span
        print(f"{k}: spread(ref CPU vs ref eager CUDA) {_q(spread)}")
        print(f"{k}: kernels vs eager CUDA              {_q(dev_k)}")
        print(f"{k}: kernels vs ref CPU                 {_q(dev_c)}")
        for dev in (dev_k, dev_c):
            for name, qf in (("max", np.max), ("p99", lambda x: np.quantile(x, 0.99)), ("p90", lambda x: np.quantile(x, 0.90))):
                assert qf(dev) <= 2.0 * qf(spread) + 2e-5, (k, name, float(qf(dev)), float(qf(spread)))
            assert (dev <= 1e-4).mean() >= (spread <= 1e-4).mean() - 0.05, k


def test_bf16_gradients_per_tensor_cosine(nb):
    """1024-ray training batch (configs[1]): every one of the 24 gradient tensors of the bf16 kernels points the same
    way as the reference's fp32 autograd gradient (cosine >= 0.99), and the whole gradient agrees to <= 0.05 rel L2
    (measured: worst cosine 0.9996, 0.0056)."""
    R = 1024
    p, m = _model(nb, 7, "bf16")
    o, d = O.random_rays(R, 8)
    tgt = np.random.default_rng(9).uniform(0, 1, (R, 3)).astype(np.float32)
    to, td_, tt = (torch.from_numpy(a).to(DEV) for a in (o, d, tgt))
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    loss = torch.mean((r._render_rays(to, td_)["rgb_map"] - tt) ** 2)
    loss.backward()
    got = {k: v.grad.detach().double().flatten() for k, v in m.named_parameters()}
    pt = TP.params_from_numpy(p, requires_grad=True, device=DEV)
    lref = torch.mean((TP.render_rays(pt, to, td_, perturb=0.0)["rgb_map"] - tt) ** 2)
    lref.backward()
    worst = (2.0, "")
    num = den = 0.0
    for k in O.PARAM_NAMES:
        a, b = got[k], pt[k].grad.detach().double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
        worst = min(worst, (cos, k))
        num += float((a - b).norm() ** 2); den += float(b.norm() ** 2)
    rel = (num / den) ** 0.5
    print(f"bf16 vs fp32-autograd gradients: worst per-tensor cosine {worst[0]:.5f} ({worst[1]}), whole-gradient rel L2 {rel:.4f}, "
          f"loss {float(loss):.6f} vs {float(lref):.6f}")
    assert worst[0] >= 0.99, worst
    assert rel <= 0.05
    assert abs(float(loss) - float(lref)) <= 1e-3 * abs(float(lref))


def test_bf16_loss_trajectory_120_steps(nb):
    """120 optimisation steps, perturb = 1, 1024 rays, Adam(5e-4), a learnable target (the loss falls ~100x).  The
    run is chaotic in the ordinary sense -- the reference's own loop run with TF32 matmuls instead of fp32 departs from
    its fp32 self by tens of percent step by step (a loss spike lands a step earlier or later) -- so the bar is that
    spread, measured here with the same random draws (eager launches on both sides consume the same torch.rand
    stream): the bf16 kernels' trajectory may not stray further from the fp32 reference than 1.5x the reference's
    TF32 run does (rms log-ratio), it coincides with the reference over the first steps, and it converges to the
    same loss level."""
    R, steps = 1024, 120
    p, m = _model(nb, 11, "bf16")
    o, d = O.random_rays(R, 12)
    dn = d / np.linalg.norm(d, axis=-1, keepdims=True)
    tgt = (0.5 + 0.5 * np.stack([np.sin(3 * dn[:, 0]), np.cos(2 * dn[:, 1]), np.sin(dn[:, 0] + dn[:, 1])], -1)).astype(np.float32)
    to, td_, tt = (torch.from_numpy(a).to(DEV) for a in (o, d, tgt))
    step = nb.TrainStep(nb.NeRFRenderer(m, DEV, perturb=1.0), nb.FlatAdam(m, lr=5e-4), R, graph=False)
    torch.manual_seed(1234)
    ours = []
    for _ in range(steps):
        step(to, td_, tt)
        ours.append(step.read_metrics()["loss"])
    runs = {}
    for tf32 in (False, True):
        tr = TP.Trainer(p, device=DEV, perturb=1.0)
        torch.manual_seed(1234)
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            runs[tf32] = np.array([float(tr.step(to, td_, tt)) for _ in range(steps)])
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
    ours, ref, ref_tf32 = np.array(ours), runs[False], runs[True]
    rms = lambda a_, b_: float(np.sqrt(np.mean(np.log(a_ / b_) ** 2)))
    d_ours, d_tf32 = rms(ours, ref), rms(ref_tf32, ref)
    gm = lambda x: float(np.exp(np.mean(np.log(x[-20:]))))
    print(f"loss first {ours[0]:.5f} / {ref[0]:.5f}; last-20 geometric mean: bf16 kernels {gm(ours):.5f}, reference fp32 {gm(ref):.5f}, "
          f"reference TF32 {gm(ref_tf32):.5f}; rms log-ratio vs the fp32 reference: bf16 kernels {d_ours:.3f}, reference TF32 {d_tf32:.3f}")
    assert gm(ref) < 0.05 * ref[0], "the reference run should make progress on this target"
    assert np.abs(ours[:4] - ref[:4]).max() <= 2e-3 * ref[0]          # same draws, same weights: the first steps coincide
    assert d_ours <= 1.5 * d_tf32 + 0.05
    assert 1 / 1.6 <= gm(ours) / gm(ref) <= 1.6


def test_coarse_loss_term_gets_its_gradient_by_default(nb):
    """renderer.py:79-80 returns differentiable coarse maps; a caller adding the usual coarse MSE term must get a coarse
    gradient without asking.  Default (lazy recompute) vs coarse_grad=True (eager save) vs False (detached)."""
    R = 256
    o, d = O.random_rays(R, 21)
    tgt = np.random.default_rng(22).uniform(0, 1, (R, 3)).astype(np.float32)
    to, td_, tt = (torch.from_numpy(a).to(DEV) for a in (o, d, tgt))
    grads = {}
    for mode in (None, True, False):
        _, m = _model(nb, 20, "fp32")
        r = nb.NeRFRenderer(m, DEV, perturb=0.0, coarse_grad=mode)
        out = r._render_rays(to, td_)
        assert out["rgb_map_coarse"].requires_grad == (mode is not False)
        loss = torch.mean((out["rgb_map"] - tt) ** 2) + torch.mean((out["rgb_map_coarse"] - tt) ** 2)
        loss.backward()
        grads[mode] = m.flat_grad.detach().double().clone()
    rel = float((grads[None] - grads[True]).norm() / grads[True].norm())
    print("coarse+fine loss: lazy vs eager-save gradient rel L2", rel, "| fine-only share",
          float(grads[False].norm() / grads[True].norm()))
    assert rel <= 1e-3                                                   # the recomputed pass is the same pass
    assert float((grads[True] - grads[False]).norm() / grads[True].norm()) > 0.1   # and the coarse term is really there

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


def test_fp32_e2e_error_is_within_the_references_own_cpu_vs_cuda_spread(nb):
    """BASELINE configs[0] (100x100 view, 64+128, perturb 0): |kernels - eager CUDA| vs |eager CUDA - reference CPU|."""
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
    # the spread itself shows the end-to-end 1e-4 gate is not attainable by ANY re-implementation, the reference's own included
    assert spread.max() > 1e-4, "the reference itself would pass 1e-4 end to end: tighten the kernels' rule"
    for name, qf in (("max", np.max), ("p99", lambda x: np.quantile(x, 0.99)), ("p90", lambda x: np.quantile(x, 0.90))):
        assert qf(dev_k) <= 2.0 * qf(spread) + 1e-5, (name, float(qf(dev_k)), float(qf(spread)))
    assert (dev_k <= 1e-4).mean() >= (spread <= 1e-4).mean() - 0.05
    assert (dev_cpu <= 1e-4).mean() >= 0.70 and dev_cpu.max() <= 5e-3      # the stated rule still holds


def test_bf16_gradients_per_tensor_cosine(nb):
    """1024-ray training batch (configs[1]): every one of the 24 gradient tensors of the bf16 kernels points the same
    way as the reference's fp32 autograd gradient (cosine >= 0.99), and the whole gradient agrees to <= 0.15 rel L2."""
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
    assert rel <= 0.15
    assert abs(float(loss) - float(lref)) <= 1e-3 * abs(float(lref))


def test_bf16_loss_trajectory_120_steps(nb):
    """120 optimisation steps, perturb = 1, 1024 rays, Adam(5e-4): the replayed TrainStep (bf16 kernels) follows the
    reference's loop (eager CUDA fp32 autograd + torch.optim.Adam) fed the same random draws."""
    R, steps = 1024, 120
    p, m = _model(nb, 11, "bf16")
    o, d = O.random_rays(R, 12)
    # a learnable target: a smooth function of the ray (so the loss actually falls, unlike i.i.d. noise targets)
    dn = d / np.linalg.norm(d, axis=-1, keepdims=True)
    tgt = (0.5 + 0.5 * np.stack([np.sin(3 * dn[:, 0]), np.cos(2 * dn[:, 1]), np.sin(dn[:, 0] + dn[:, 1])], -1)).astype(np.float32)
    to, td_, tt = (torch.from_numpy(a).to(DEV) for a in (o, d, tgt))
    r = nb.NeRFRenderer(m, DEV, perturb=1.0)
    step = nb.TrainStep(r, nb.FlatAdam(m, lr=5e-4), R)
    tr = TP.Trainer(p, device=DEV, perturb=1.0)
    ours, ref = [], []
    torch.manual_seed(1234)
    for _ in range(steps):
        step(to, td_, tt)
        ours.append(step.read_metrics()["loss"])
    torch.manual_seed(1234)
    for _ in range(steps):
        ref.append(float(tr.step(to, td_, tt)))
    ours, ref = np.array(ours), np.array(ref)
    rel = np.abs(ours - ref) / ref
    print(f"loss: first {ours[0]:.5f} / {ref[0]:.5f}, last {ours[-1]:.5f} / {ref[-1]:.5f}; max rel diff {rel.max():.4f}, "
          f"mean {rel.mean():.4f}")
    assert ref[-1] < 0.7 * ref[0], "the reference run should make progress on this target"
    assert rel[:10].max() <= 5e-3                    # same draws, same weights: the first steps coincide
    assert rel.max() <= 0.05 and rel.mean() <= 0.02  # and the trajectories stay together
    assert abs(np.mean(ours[-10:]) - np.mean(ref[-10:])) <= 0.03 * np.mean(ref[-10:])


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

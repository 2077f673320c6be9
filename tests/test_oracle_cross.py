"""The numpy checker (oracle/nerf_oracle.py) against the torch restatement (oracle/nerf_oracle_torch.py, which
reproduces the reference bit for bit on the golden vectors) on seeded random configurations the golden files do
not contain: ragged ray counts (1, 3, 130), small and odd sample counts, black / white background, density noise,
coord_scale, stratified jitter.  CPU only; widens the pin of the checker that the CUDA parity tests rely on.

Rules (as in test_oracle_golden.py): coarse maps and stage-isolated fine maps agree to fp32 rounding (<= 1e-5);
searchsorted indices are bit-exact given the same cdf; the analytic compositing backward matches autograd within
2e-4 + 2e-3 relative; parameter gradients (stage-isolated) within 2e-4 relative L2 per tensor (1e-3 for the two
sigma_linear tensors: they are plain sums of d_sigma over all samples, where autograd's divide-by-input cumprod
backward and the checker's suffix-sum form round differently and the sum cancels).
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import nerf_oracle_torch as T

CASES = [
    # R, S_c, N_imp, perturb, white, coord_scale, noise_std, seed
    (1, 64, 128, 0.0, True, 1.0, 0.0, 1),
    (3, 8, 16, 1.0, False, 1.0, 0.0, 2),
    (130, 16, 24, 1.0, True, 0.5, 0.0, 3),
    (17, 33, 7, 0.0, False, 2.0, 0.0, 4),
    (9, 64, 128, 1.0, True, 1.0, 0.7, 5),
    (5, 12, 0, 1.0, True, 1.0, 0.0, 6),
]


def _draws(R, S, N, perturb, noise_std, seed):
    g = np.random.default_rng(100 + seed)
    t_rand = g.uniform(0, 1, (R, S)).astype(np.float32) if perturb > 0 else None
    u = (g.uniform(0, 1, (R, N)).astype(np.float32) if perturb > 0 else np.linspace(0, 1, N, dtype=np.float32)) if N > 0 else None
    nc = (g.standard_normal((R, S)) * noise_std).astype(np.float32) if noise_std > 0 else None
    nf = (g.standard_normal((R, S + N)) * noise_std).astype(np.float32) if noise_std > 0 and N > 0 else None
    return t_rand, u, nc, nf


@pytest.mark.parametrize("case", CASES, ids=lambda c: "R%d_S%d+%d_p%g_w%d_cs%g_n%g" % c[:7])
def test_numpy_checker_tracks_torch_restatement(case):
    R, S, N, perturb, white, cs, noise_std, seed = case
    p = O.init_params(seed)
    o, d = O.random_rays(R, seed)
    t_rand, u, nc, nf = _draws(R, S, N, perturb, noise_std, seed)
    tt = lambda a: None if a is None else torch.from_numpy(a)
    with torch.no_grad():
        ref = T.render_rays(T.params_from_numpy(p), tt(o), tt(d), N_samples=S, N_importance=N, white_bkgd=white,
                            perturb=perturb, raw_noise_std=noise_std, coord_scale=cs, t_rand=tt(t_rand), u=tt(u),
                            noise_coarse=tt(nc), noise_fine=tt(nf))
    cfg = O.RenderConfig(N_samples=S, N_importance=N, white_bkgd=white, coord_scale=cs)
    t_vals = torch.linspace(0., 1., S).numpy()
    kw = dict(t_rand=t_rand, noise_coarse=nc, noise_fine=nf)
    if N == 0:
        out = O.render_rays(p, o, d, cfg, t_vals, None, **kw)
        for k in ("rgb_map", "depth_map", "acc_map"):
            np.testing.assert_allclose(out[k], ref[k].numpy(), atol=1e-5, rtol=0)
        return
    # stage-isolated: the checker continues from the reference's own fine depths
    out = O.render_rays(p, o, d, cfg, t_vals, u, z_fine_override=ref["z_fine"].numpy(), **kw)
    for k in ("rgb_map_coarse", "depth_map_coarse", "acc_map_coarse", "rgb_map", "depth_map", "acc_map"):
        np.testing.assert_allclose(out[k], ref[k].numpy(), atol=1e-5, rtol=0, err_msg=k)
    np.testing.assert_allclose(out["weights_coarse"], ref["weights_coarse"].numpy(), atol=2e-6, rtol=1e-5)
    # the checker's own resampling: merged depths are sorted, contain the coarse depths, and sit inside the bins
    own = O.render_rays(p, o, d, cfg, t_vals, u, **kw)
    zf = own["z_fine"]
    assert zf.shape == (R, S + N) and np.all(zf[:, 1:] >= zf[:, :-1])
    assert np.all(zf >= own["z_vals"].min(-1, keepdims=True) - 1e-6) and np.all(zf <= own["z_vals"].max(-1, keepdims=True) + 1e-6)
    # indices: pure compare/count on the reference's cdf bits
    w = ref["weights_coarse"].numpy()[:, 1:-1]
    cdf_ref = torch.cat([torch.zeros(R, 1), torch.cumsum(torch.from_numpy(w + np.float32(1e-5)) /
                                                       torch.sum(torch.from_numpy(w + np.float32(1e-5)), -1, keepdim=True), -1)], -1)
    uu = np.broadcast_to(u, (R, N)) if u.ndim == 1 else u
    ref_inds = torch.searchsorted(cdf_ref, torch.from_numpy(np.array(uu, order="C")), right=True).numpy()
    assert np.array_equal(O.searchsorted_right(cdf_ref.numpy(), uu), ref_inds)


@pytest.mark.parametrize("white", [True, False])
def test_compositing_backward_matches_autograd(white):
    g = np.random.default_rng(7)
    R, S = 11, 37
    raw = g.standard_normal((R, S, 4)).astype(np.float32)
    raw[..., 3] *= 3.0
    z = np.sort(g.uniform(2, 6, (R, S)).astype(np.float32), -1)
    rd = g.standard_normal((R, 3)).astype(np.float32)
    g_rgb, g_depth, g_acc = (g.standard_normal(s).astype(np.float32) for s in ((R, 3), (R,), (R,)))
    g_w = g.standard_normal((R, S)).astype(np.float32)
    traw = torch.from_numpy(raw).requires_grad_(True)
    rgb, depth, acc, w = T.raw2outputs(traw, torch.from_numpy(z), torch.from_numpy(rd), white)
    (rgb * torch.from_numpy(g_rgb)).sum().add((depth * torch.from_numpy(g_depth)).sum()).add(
        (acc * torch.from_numpy(g_acc)).sum()).add((w * torch.from_numpy(g_w)).sum()).backward()
    ref = traw.grad.numpy()
    got = O.raw2outputs_backward(raw, z, rd, white, g_rgb, g_depth, g_acc, g_w)
    assert np.all(np.abs(got - ref) <= 2e-4 + 2e-3 * np.abs(ref)), float(np.abs(got - ref).max())


def test_train_gradients_match_autograd_on_a_ragged_batch():
    R, seed = 37, 8
    p = O.init_params(seed)
    o, d = O.random_rays(R, seed)
    g = np.random.default_rng(seed)
    t_rand = g.uniform(0, 1, (R, 24)).astype(np.float32)
    u = g.uniform(0, 1, (R, 40)).astype(np.float32)
    tgt = g.uniform(0, 1, (R, 3)).astype(np.float32)
    pt = T.params_from_numpy(p, requires_grad=True)
    ref = T.render_rays(pt, torch.from_numpy(o), torch.from_numpy(d), N_samples=24, N_importance=40, perturb=1.0,
                        t_rand=torch.from_numpy(t_rand), u=torch.from_numpy(u))
    loss_ref = torch.mean((ref["rgb_map"] - torch.from_numpy(tgt)) ** 2)
    loss_ref.backward()
    cfg = O.RenderConfig(N_samples=24, N_importance=40)
    loss, grads, _ = O.train_grads(p, o, d, tgt, cfg, torch.linspace(0., 1., 24).numpy(), u, t_rand,
                                   z_fine_override=ref["z_fine"].detach().numpy())
    assert abs(float(loss) - float(loss_ref.detach())) < 1e-6
    for k in O.PARAM_NAMES:
        a, b = grads[k].astype(np.float64), pt[k].grad.numpy().astype(np.float64)
        tol = 1e-3 if k.startswith("sigma_linear") else 2e-4
        assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b) + 1e-12, (k, np.linalg.norm(a - b) / np.linalg.norm(b))


# ---- property tests (hypothesis): resampling on degenerate pdfs ---------------------------------------------------
from hypothesis import given, settings, strategies as st_  # noqa: E402


def _torch_cdf(w):
    tw = torch.from_numpy(w) + 1e-5
    pdf = tw / torch.sum(tw, -1, keepdim=True)
    return torch.cat([torch.zeros(w.shape[0], 1), torch.cumsum(pdf, -1)], -1)


@settings(max_examples=40, deadline=None)
@given(seed=st_.integers(0, 10_000), nbins=st_.integers(3, 70), nimp=st_.integers(1, 96),
       kind=st_.sampled_from(["zeros", "spike", "uniform", "sparse", "huge"]), det=st_.booleans())
def test_sample_pdf_on_degenerate_weights(seed, nbins, nimp, kind, det):
    """All-zero weights, one spike, exactly uniform, mostly-zero and 1e6-scale weights; u on a grid that hits 0 and 1
    exactly or random: given the reference's cdf bits the checker's indices are bit-exact and its samples within 2e-6;
    with its own cdf the samples stay inside the bins and are monotone for sorted u."""
    g = np.random.default_rng(seed)
    R = 5
    bins = np.sort(g.uniform(2, 6, (R, nbins)).astype(np.float32), -1)
    w = {"zeros": np.zeros((R, nbins - 1)), "uniform": np.full((R, nbins - 1), 0.25),
         "spike": np.eye(nbins - 1)[g.integers(0, nbins - 1, R)] * 3.0,
         "sparse": g.uniform(0, 1, (R, nbins - 1)) * (g.uniform(0, 1, (R, nbins - 1)) < 0.15),
         "huge": g.uniform(0, 1e6, (R, nbins - 1))}[kind].astype(np.float32)
    u = np.linspace(0, 1, nimp, dtype=np.float32) if det else np.sort(g.uniform(0, 1, (R, nimp)).astype(np.float32), -1)
    ref = T.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), torch.from_numpy(u)).numpy()
    cdf_ref = _torch_cdf(w).numpy()
    uu = np.broadcast_to(u, (R, nimp)) if u.ndim == 1 else u
    inds_ref = torch.searchsorted(torch.from_numpy(cdf_ref), torch.from_numpy(np.array(uu, order="C")), right=True).numpy()
    assert np.array_equal(O.searchsorted_right(cdf_ref, uu), inds_ref)
    np.testing.assert_allclose(O.sample_pdf(bins, w, u, cdf=cdf_ref), ref, atol=2e-6, rtol=0)
    own = O.sample_pdf(bins, w, u)
    assert np.all(own >= bins[:, :1] - 1e-6) and np.all(own <= bins[:, -1:] + 1e-6)
    assert np.all(np.diff(own, axis=-1) >= -2e-6)
    assert np.all(np.isfinite(own)) and np.all(np.isfinite(ref))

"""MLP + end-to-end parity of the CUDA path (through the drop-in Python classes / C ABI) against
the oracle and the reference's golden vectors.  Needs a B200: `pytest -m gpu`.

Tolerance rules (written down before the first measurement):

  fp32 check mode
    * MLP raw outputs: <= 2e-5 abs vs the oracle.
    * coarse maps (no resampling): <= 1e-5 abs.
    * fine maps, stage-isolated (oracle continues from the kernel's own z_fine): <= 1e-4 abs, all rays
      -- this is north_star's 1e-4 gate.
    * fine maps end to end vs the reference's golden maps: >= 70 % of rays within 1e-4, none above
      5e-3 (the inverse cdf is ill-conditioned at random init; same rule as the oracle-vs-golden test).
    * parameter gradients, stage-isolated: per-tensor relative L2 <= 1e-3.

  bf16 tensor-core mode
    * MLP raw outputs: <= 2e-2 abs vs the fp32 oracle (256-wide bf16 layers, fp32 accumulate).
    * maps: rays whose last-sample sigma is within 4e-3 of 0 in the oracle are "flip-prone"
      (alpha_last = [sigma_last > 0], renderer.py:123: a step function) and are excluded and counted;
      of the rest >= 99 % must be within 1e-2 abs on rgb/acc and within 1e-2 * (far-near) on depth,
      and none may exceed 10x that.
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dtype)


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def nb():
    import nerf_mlp_b200
    return nerf_mlp_b200


def make_model(nb, seed, precision):
    p = O.init_params(seed)
    m = nb.NeRFMLP(precision=precision)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    return m.to(DEV), p


# ------------------------------------------------------------------------------------------------
# NeRFMLP.forward on encoded inputs
# ------------------------------------------------------------------------------------------------
def test_state_dict_surface(nb):
    m = nb.NeRFMLP().to(DEV)
    sd = m.state_dict()
    assert list(sd) == list(O.PARAM_NAMES)
    assert all(tuple(sd[k].shape) == O.init_params(0)[k].shape for k in sd)
    assert len(list(m.parameters())) == 24 and all(p.dtype == torch.float32 and p.is_cuda for p in m.parameters())
    # parameters are views of one flat buffer, in state_dict order
    flat = m.flat_params
    assert flat.numel() == 595844 and next(m.parameters()).data_ptr() == flat.data_ptr()
    p = O.init_params(3)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    assert np.array_equal(N(m.flat_params), O.flatten_params(p))


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_mlp_forward_encoded(nb, precision, tol):
    st = load_golden("stages")
    m, p = make_model(nb, 7, precision)
    with torch.no_grad():
        out = N(m(T(st["pe10"][:200]), T(st["pe4"][:200])))
    assert out.shape == (200, 4)
    np.testing.assert_allclose(out, st["mlp_out"], atol=tol, rtol=0)        # vs the reference
    # ragged sizes around the 128/256-row tiles, incl. empty
    rng = np.random.default_rng(0)
    for M in (0, 1, 127, 128, 129, 255, 257, 1000):
        x = O.positional_encoding(rng.uniform(-4, 4, (M, 3)).astype(np.float32), 10)
        d = rng.standard_normal((M, 3)).astype(np.float32)
        d = O.positional_encoding(d / np.linalg.norm(d, axis=-1, keepdims=True), 4)
        with torch.no_grad():
            out = N(m(T(x), T(d)))
        assert out.shape == (M, 4)
        if M:
            np.testing.assert_allclose(out, O.mlp_forward(p, x, d), atol=tol, rtol=0, err_msg=f"M={M}")


def test_load_from_numpy(nb):
    p = O.init_params(5)
    arrs = []
    for i in range(8):
        arrs += [p[f"pts_linears.{i}.weight"].T.copy(), p[f"pts_linears.{i}.bias"]]
    for n in ("bottleneck_linear", "view_linear", "rgb_linear", "sigma_linear"):     # model.py:99-122 order
        arrs += [p[f"{n}.weight"].T.copy(), p[f"{n}.bias"]]
    m = nb.NeRFMLP(precision="fp32").to(DEV)
    m.load_from_numpy(arrs)
    assert np.array_equal(N(m.flat_params), O.flatten_params(p))


# ------------------------------------------------------------------------------------------------
# _render_rays end to end
# ------------------------------------------------------------------------------------------------
def run_case(nb, g, precision, check_mode=True):
    ns, ni, perturb, wb, cs, noise_std, seed = g["cfg"]
    m, p = make_model(nb, int(seed), precision)
    r = nb.NeRFRenderer(m, DEV, N_samples=int(ns), N_importance=int(ni), near=2.0, far=6.0, white_bkgd=bool(wb),
                        perturb=float(perturb), raw_noise_std=float(noise_std), coord_scale=float(cs))
    return m, p, r


def feed_random(monkeypatch, g):
    """Make the renderer's torch.rand / torch.randn draws return the golden run's numbers, in the
    reference's order (renderer.py:60 -> :136 -> :182 -> :136)."""
    rands = [g[k] for k in ("t_rand", "u_rand") if k in g]
    randns = [g[k] for k in ("noise_coarse", "noise_fine") if k in g]
    noise_std = float(g["cfg"][5])

    def fake_rand(*a, **k):
        return T(rands.pop(0))

    def fake_randn(*a, **k):
        return T(randns.pop(0)) / noise_std

    monkeypatch.setattr(torch, "rand", fake_rand)
    monkeypatch.setattr(torch, "randn", fake_randn)


def oracle_render(p, g, z_fine=None):
    ns, ni, perturb, wb, cs, noise_std, seed = g["cfg"]
    cfg = O.RenderConfig(N_samples=int(ns), N_importance=int(ni), white_bkgd=bool(wb), coord_scale=float(cs))
    u = None if ni == 0 else (g["u_rand"] if perturb > 0 else g["u_det"])
    return O.render_rays(p, g["rays_o"], g["rays_d"], cfg, g["t_vals"], u, t_rand=g.get("t_rand"),
                         noise_coarse=g.get("noise_coarse"), noise_fine=g.get("noise_fine"), z_fine_override=z_fine)


CASES = ["render_det_r96", "render_pinhole_12x12", "render_perturb_r48", "render_noise_blackbg_r32",
         "render_nofine_r32", "render_s128_256_r16"]


@pytest.mark.parametrize("name", CASES)
def test_render_rays_fp32(nb, name, monkeypatch):
    g = load_golden(name)
    m, p, r = run_case(nb, g, "fp32")
    feed_random(monkeypatch, g)
    # the renderer's own linspace must be the one the reference used
    assert np.array_equal(N(r._linspace(int(g["cfg"][0]))), g["t_vals"])
    with torch.no_grad():
        out = {k: N(v) for k, v in r._render_rays(T(g["rays_o"]), T(g["rays_d"])).items()}
    keys = [k[4:] for k in g if k.startswith("out_")]
    assert set(out) == set(keys)
    fine = int(g["cfg"][1]) > 0
    for k in keys:
        err = np.abs(out[k] - g["out_" + k]).reshape(out[k].shape[0], -1).max(-1)
        if k.endswith("_coarse") or not fine:
            assert err.max() <= 1e-5, (k, err.max())
        else:
            assert (err <= 1e-4).mean() >= 0.70 and err.max() <= 5e-3, (k, err.max(), (err <= 1e-4).mean())


@pytest.mark.parametrize("name", ["render_det_r96", "render_perturb_r48", "render_s128_256_r16"])
def test_render_rays_fp32_stage_isolated(nb, name, monkeypatch):
    """north_star's 1e-4 fp32 gate with the ill-conditioned inverse cdf taken out: the oracle's fine
    pass continues from the kernel's own z_fine (exported in check mode)."""
    g = load_golden(name)
    m, p, r = run_case(nb, g, "fp32")
    feed_random(monkeypatch, g)
    R = g["rays_o"].shape[0]
    o, d = T(g["rays_o"]), T(g["rays_d"])
    with torch.no_grad():
        t_rand = torch.rand((R, r.N_samples)) if r.perturb > 0 else None
        z = nb.ops.stratified_z(r._linspace(r.N_samples), t_rand, R, r.near, r.far)
        rgb0, depth0, acc0, w = r._pass(o, d, z, False)
        u = r._linspace(r.N_importance) if r.perturb == 0 else torch.rand((R, r.N_importance))
        z_fine, zs, inds, cdf = nb.ops.resample_merge(z, w, u, check_mode=True)
        rgb, depth, acc, _ = r._pass(o, d, z_fine, False)
    ref = oracle_render(p, g, z_fine=N(z_fine))
    assert np.array_equal(N(z), ref["z_vals"])
    np.testing.assert_allclose(N(w), ref["weights_coarse"], atol=5e-6)
    # searchsorted indices bit-exact given the kernel's cdf; cdf = the oracle's bits on the same
    # weights (the end-to-end cdf inherits the ~1e-6 relative differences of the coarse weights)
    uu = np.broadcast_to(N(u), (R, r.N_importance))
    assert np.array_equal(N(inds), O.searchsorted_right(N(cdf), uu))
    assert np.sum(N(cdf) != O.pdf_to_cdf(N(w)[:, 1:-1])) <= 2
    np.testing.assert_allclose(N(cdf), ref["cdf"], atol=1e-4)
    np.testing.assert_allclose(N(rgb), ref["rgb_map"], atol=1e-4)
    np.testing.assert_allclose(N(depth), ref["depth_map"], atol=1e-4)
    np.testing.assert_allclose(N(acc), ref["acc_map"], atol=1e-4)


def bf16_check(out, ref, sigma_last, far_near=4.0):
    flip = np.abs(sigma_last) < 4e-3      # ~5x the measured bf16 error of raw sigma (7e-4)
    keep = ~flip
    res = {"flip_prone": int(flip.sum()), "rays": int(flip.size)}
    for k, scale in (("rgb_map", 1.0), ("acc_map", 1.0), ("depth_map", far_near)):
        err = np.abs(out[k] - ref[k]).reshape(flip.size, -1).max(-1)[keep]
        tol = 1e-2 * scale
        res[k] = (float(err.max()), float(np.quantile(err, 0.99)))
        assert (err <= tol).mean() >= 0.99, (k, res)
        assert err.max() <= 10 * tol, (k, res)
    return res


@pytest.mark.parametrize("name", ["render_det_r96", "render_pinhole_12x12", "render_perturb_r48", "render_s128_256_r16"])
def test_render_rays_bf16(nb, name, monkeypatch):
    g = load_golden(name)
    m, p, r = run_case(nb, g, "bf16")
    feed_random(monkeypatch, g)
    with torch.no_grad():
        out = {k: N(v) for k, v in r._render_rays(T(g["rays_o"]), T(g["rays_d"])).items()}
    ref = oracle_render(p, g)
    golden = {k[4:]: v for k, v in g.items() if k.startswith("out_")}
    res = bf16_check(out, golden, ref["raw_fine"][:, -1, 3])
    print(name, res)


def test_render_bf16_larger(nb):
    """1024 random rays, 64+128, bf16 vs the fp32 oracle: the statistical rule on a larger sample,
    with the flip count reported."""
    R = 1024
    m, p = make_model(nb, 21, "bf16")
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    o, d = O.random_rays(R, 5)
    with torch.no_grad():
        out = {k: N(v) for k, v in r._render_rays(T(o), T(d)).items()}
    ref = O.render_rays(p, o, d, O.RenderConfig(), N(r._linspace(64)), N(r._linspace(128)))
    print("bf16 1024 rays:", bf16_check(out, ref, ref["raw_fine"][:, -1, 3]))


def test_render_image_entry(nb):
    """render(): chunk loop, (H,W,3) output, requires N == H*W (renderer.py:23-45)."""
    m, p = make_model(nb, 2, "fp32")
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    o, d, focal = O.pinhole_rays(12, 12)
    img = r.render(T(o), T(d), 12, 12, focal, chunk=50)
    g = load_golden("render_pinhole_12x12")
    assert img.shape == (12, 12, 3) and not img.requires_grad
    err = np.abs(N(img).reshape(-1, 3) - g["out_rgb_map"]).max(-1)
    assert (err <= 1e-4).mean() >= 0.70 and err.max() <= 5e-3
    with pytest.raises(RuntimeError):
        r.render(T(o), T(d), 5, 5, focal)


# ------------------------------------------------------------------------------------------------
# training step: gradients, optimiser
# ------------------------------------------------------------------------------------------------
def rel_l2(a, b):
    a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def grads_via_api(nb, m, r, g):
    out = r._render_rays(T(g["rays_o"]), T(g["rays_d"]))
    loss = torch.mean((out["rgb_map"] - T(g["target"])) ** 2)          # scripts/train.py:376
    m.zero_grad()
    loss.backward()
    return float(loss), {k: N(prm.grad) for k, prm in m.named_parameters()}


def test_train_grads_fp32(nb):
    g = load_golden("train_r32")
    m, p = make_model(nb, int(g["seed"]), "fp32")
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    loss, grads = grads_via_api(nb, m, r, g)
    assert abs(loss - float(g["loss"])) < 1e-5
    # (1) stage-isolated: oracle gradient from the kernel's own z_fine
    with torch.no_grad():
        z = nb.ops.stratified_z(r._linspace(64), None, 32, 2.0, 6.0)
        w = r._pass(T(g["rays_o"]), T(g["rays_d"]), z, False)[3]
        z_fine = N(nb.ops.resample_merge(z, w, r._linspace(128)))
    _, og, _ = O.train_grads(p, g["rays_o"], g["rays_d"], g["target"], O.RenderConfig(), g["t_vals"], g["u_det"],
                             z_fine_override=z_fine)
    worst = max((rel_l2(grads[k], og[k]), k) for k in O.PARAM_NAMES)
    print("fp32 grad rel-L2 (stage-isolated), worst:", worst)
    assert worst[0] <= 1e-3, worst
    # (2) vs the reference's own autograd (golden subsamples), end to end: <= 1e-2 (see oracle test)
    for k in O.PARAM_NAMES:
        gr = grads[k].reshape(-1)
        sub = gr[::97] if gr.size > 4096 else gr
        assert rel_l2(sub, g["gsub_" + k]) < 1e-2, k
    # flat gradient buffer: p.grad are views of it
    assert m.flat_grad is not None and m.flat_grad.data_ptr() == next(m.parameters()).grad.data_ptr()


def test_grad_accumulation_and_zero_grad(nb):
    g = load_golden("train_r32")
    m, p = make_model(nb, int(g["seed"]), "fp32")
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    _, g1 = grads_via_api(nb, m, r, g)
    # second backward without zero_grad accumulates (autograd semantics)
    out = r._render_rays(T(g["rays_o"]), T(g["rays_d"]))
    torch.mean((out["rgb_map"] - T(g["target"])) ** 2).backward()
    g2 = {k: N(prm.grad) for k, prm in m.named_parameters()}
    for k in ("pts_linears.3.weight", "rgb_linear.bias"):
        assert rel_l2(g2[k], 2 * g1[k]) < 1e-4
    m.zero_grad(set_to_none=True)
    assert all(prm.grad is None for prm in m.parameters())
    _, g3 = grads_via_api(nb, m, r, g)
    assert rel_l2(g3["pts_linears.3.weight"], g1["pts_linears.3.weight"]) < 1e-4


def test_optimizer_steps_match_reference(nb):
    """Two steps of torch.optim.Adam(model.parameters(), lr=5e-4) on the drop-in model (the
    reference's training loop, scripts/train.py:381-388) and of FlatAdam reproduce the golden
    parameters of the reference run."""
    g = load_golden("train_r32")
    for kind in ("torch", "flat"):
        m, p = make_model(nb, int(g["seed"]), "fp32")
        r = nb.NeRFRenderer(m, DEV, perturb=0.0)
        opt = torch.optim.Adam(m.parameters(), lr=5e-4) if kind == "torch" else nb.FlatAdam(m, lr=5e-4)
        for step in (1, 2):
            out = r._render_rays(T(g["rays_o"]), T(g["rays_d"]))
            loss = torch.mean((out["rgb_map"] - T(g["target"])) ** 2)
            opt.zero_grad()
            loss.backward()
            opt.step()
            flat = N(m.flat_params)
            ref = g[f"params_after_step{step}_sub"]
            np.testing.assert_allclose(flat[::101], ref, atol=2.1e-3)
            assert np.mean(np.abs(flat[::101] - ref) < 2e-5) > 0.97, kind
        assert abs(float(loss) - float(g["loss_step1"])) < 1e-4


def test_train_grads_bf16(nb):
    """bf16 tensor-core backward.  Two stated tolerances (per-tensor relative L2):
      (1) <= 2e-2 against the oracle's fp32 backward evaluated on the bf16 forward's OWN saved
          tensors (same ReLU pattern): this isolates the backward kernels (their only error is the
          bf16 rounding of the d(pre-activation) tensors);
      (2) <= 0.25 against the all-fp32 oracle gradient.  The gap between (1) and (2) is a property
          of the bf16 *forward*: ~0.1 % of the hidden units sit close enough to 0 for their ReLU
          to switch state under bf16 rounding, and each switch changes that unit's gradient by
          100 % (relative L2 ~ sqrt(fraction) per layer, compounding down the chain)."""
    g = load_golden("train_r32")
    m, p = make_model(nb, int(g["seed"]), "bf16")
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    loss, grads = grads_via_api(nb, m, r, g)
    assert abs(loss - float(g["loss"])) < 2e-3
    R, S = 32, 192
    with torch.no_grad():
        z = nb.ops.stratified_z(r._linspace(64), None, R, 2.0, 6.0)
        w = r._pass(T(g["rays_o"]), T(g["rays_d"]), z, False)[3]
        z_fine = nb.ops.resample_merge(z, w, r._linspace(128))
    # (1) same-pattern reference: re-run the fine pass with save, build the oracle's `saved` from it
    o, d = T(g["rays_o"]), T(g["rays_d"])
    raw, ws = nb.ops.mlp_fwd_rays(m, o, d, z_fine, 1.0, nb._lib.PREC_BF16, True)
    rgb, depth, acc, _ = nb.ops.composite_fwd(raw, z_fine, d, None, True, True)
    d_rgb = (2.0 * (rgb - T(g["target"])) / rgb.numel()).contiguous()
    d_raw = nb.ops.composite_bwd(raw, z_fine, d, None, True, d_rgb)
    flat = torch.zeros_like(m.flat_params)
    nb.ops.mlp_bwd(m, d_raw, ws, nb._lib.PREC_BF16, flat, S)
    v = nb.ops.bf16_workspace_views(ws, R * S)
    ar = torch.arange(32, device=DEV, dtype=torch.int32)
    bits = [N(((v["mask"][l].unsqueeze(-1) >> ar) & 1).reshape(R * S, 256)).astype(np.float32) for l in range(8)]
    hvbits = N(((v["hvmask"].unsqueeze(-1) >> ar) & 1).reshape(R * S, 128)).astype(np.float32)
    act = N(v["act"].float())
    xenc = N(v["xenc"].float())[:, :63]
    de = np.repeat(N(v["de"])[:R, :27], S, axis=0)
    assert np.abs(N(v["de16"].float())[:, :27] - de).max() < 8e-3       # per-sample bf16 copy used by the wgrad
    ins = [xenc] + [act[l - 1] for l in range(1, 5)] + [np.concatenate([xenc, act[4]], 1)] + [act[5], act[6]]
    saved = {"in": ins, "out": bits, "h7": act[7], "hv_in": np.concatenate([act[8], de], 1),
             "hv": N(v["hv"].float()) * hvbits}
    saved["hv"] = np.where(hvbits > 0, np.maximum(saved["hv"], 1e-30), 0.0).astype(np.float32)   # keep the pattern exact
    ref = O.mlp_backward(p, saved, N(d_raw).reshape(-1, 4))
    got = dict(zip(O.PARAM_NAMES, [N(t) for t in m._views_of(flat)]))
    worst = max((rel_l2(got[k], ref[k]), k) for k in O.PARAM_NAMES)
    print("bf16 backward kernels vs same-pattern fp32 reference, worst rel-L2:", worst)
    assert worst[0] <= 2e-2, worst
    # the autograd path produced the same numbers as the direct kernel calls
    assert max(rel_l2(grads[k], got[k]) for k in O.PARAM_NAMES) < 1e-3
    # (2) end to end vs the all-fp32 oracle gradient (stage-isolated on the kernel's z_fine)
    _, og, _ = O.train_grads(p, g["rays_o"], g["rays_d"], g["target"], O.RenderConfig(), g["t_vals"], g["u_det"],
                             z_fine_override=N(z_fine))
    worst = max((rel_l2(grads[k], og[k]), k) for k in O.PARAM_NAMES)
    print("bf16 grads vs all-fp32 oracle, worst rel-L2:", worst)
    assert worst[0] <= 0.25, worst


def test_bf16_training_reduces_loss(nb):
    """A few Adam steps in bf16 mode on a fixed batch reduce the loss like the fp32 path does."""
    g = load_golden("train_r32")
    losses = {}
    for prec in ("fp32", "bf16"):
        m, p = make_model(nb, int(g["seed"]), prec)
        r = nb.NeRFRenderer(m, DEV, perturb=0.0)
        opt = nb.FlatAdam(m, lr=5e-4)
        ls = []
        for _ in range(12):
            out = r._render_rays(T(g["rays_o"]), T(g["rays_d"]))
            loss = torch.mean((out["rgb_map"] - T(g["target"])) ** 2)
            opt.zero_grad()
            loss.backward()
            opt.step()
            ls.append(float(loss))
        losses[prec] = ls
    print("loss curves:", {k: [round(x, 5) for x in v[::3]] for k, v in losses.items()})
    assert losses["bf16"][-1] < losses["bf16"][0] * 0.97
    assert losses["fp32"][-1] < losses["fp32"][0] * 0.97
    assert abs(losses["bf16"][-1] - losses["fp32"][-1]) < 0.05 * losses["fp32"][0]


def test_density_only_coarse_pass(nb):
    """NERF_FWD_DENSITY_ONLY (include/nerf_b200.h): sigma is bit-identical to the full forward's, the colour
    channels are zero, and render() -- whose coarse pass runs in this mode -- returns bit-identical pixels to
    the full _render_rays path (same weights -> same fine samples -> same fine maps)."""
    m, p = make_model(nb, 3, "bf16")
    R, S = 777, 64
    o, d = O.random_rays(R, 5)
    z = nb.ops.stratified_z(torch.linspace(0., 1., S, device=DEV), None, R, 2.0, 6.0)
    full, _ = nb.ops.mlp_fwd_rays(m, T(o), T(d), z, 1.0, nb._lib.PREC_BF16, False)
    dens, _ = nb.ops.mlp_fwd_rays(m, T(o), T(d), z, 1.0, nb._lib.PREC_BF16, False, density_only=True)
    assert torch.equal(dens[..., 3], full[..., 3])
    assert float(dens[..., :3].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        nb.ops.mlp_fwd_rays(m, T(o), T(d), z, 1.0, nb._lib.PREC_BF16, True, density_only=True)
    for prec in ("bf16", "fp32"):
        m2, _ = make_model(nb, 3, prec)
        r = nb.NeRFRenderer(m2, DEV, perturb=0.0)
        img = r.render(T(o), T(d), R, 1, 1.0, chunk=300)
        with torch.no_grad():
            ref = torch.cat([r._render_rays(T(o)[i:i + 300], T(d)[i:i + 300])["rgb_map"] for i in range(0, R, 300)])
        assert torch.equal(img.view(R, 3), ref), prec


def test_render_100x100_against_reference_image(nb):
    """BASELINE.json configs[0]: the whole 100x100 view through `NeRFRenderer.render` (default chunking) against the
    image the reference's own CPU `renderer.render` produced (tests/golden/render_pinhole_100x100.npz).
      fp32 check mode: >= 70 % of the pixels within 1e-4, none above 5e-3 (the end-to-end rule through the
                       ill-conditioned inverse cdf, see the module docstring);
      bf16 mode:       >= 99 % within 1e-2 excluding the flip-prone pixels (|sigma_last| < 4e-3 in the fp32 run), none
                       of the others above 1e-1."""
    g = load_golden("render_pinhole_100x100")
    ref = g["image"].reshape(-1, 3)
    o, d, focal = O.pinhole_rays(100, 100)
    m32, _ = make_model(nb, int(g["seed"]), "fp32")
    r32 = nb.NeRFRenderer(m32, DEV, perturb=0.0)
    img32 = N(r32.render(T(o), T(d), 100, 100, focal)).reshape(-1, 3)
    err = np.abs(img32 - ref).max(-1)
    assert (err <= 1e-4).mean() >= 0.70 and err.max() <= 5e-3, (err.max(), (err <= 1e-4).mean())
    # flip-prone pixels from the fp32 run's last-sample density
    with torch.no_grad():
        z = nb.ops.stratified_z(r32._linspace(64), None, 10000, 2.0, 6.0)
        w = r32._pass(T(o), T(d), z, False)[3]
        z_fine = nb.ops.resample_merge(z, w, r32._linspace(128))
        raw, _ = nb.ops.mlp_fwd_rays(m32, T(o), T(d), z_fine, 1.0, nb._lib.PREC_FP32, False)
    flip = np.abs(N(raw[:, -1, 3])) < 4e-3
    m16, _ = make_model(nb, int(g["seed"]), "bf16")
    img16 = N(nb.NeRFRenderer(m16, DEV, perturb=0.0).render(T(o), T(d), 100, 100, focal)).reshape(-1, 3)
    e16 = np.abs(img16 - ref).max(-1)[~flip]
    assert (e16 <= 1e-2).mean() >= 0.99 and e16.max() <= 1e-1, (e16.max(), (e16 <= 1e-2).mean(), int(flip.sum()))

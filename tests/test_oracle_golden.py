"""Pin oracle/nerf_oracle.py against the golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only.

Tolerances: the oracle and the reference are both fp32 but use different GEMM kernels / summation
orders, so stage outputs agree to a few ulp (1e-5 abs on O(1) quantities); integer work
(searchsorted indices given the same cdf, sort-merge) is bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import nerf_oracle as O
from tests.conftest import load_golden


@pytest.fixture(scope="module")
def st():
    return load_golden("stages")


def test_param_count_and_names():
    p = O.init_params(0)
    assert len(p) == 24 and O.flatten_params(p).size == 595844
    assert p["pts_linears.5.weight"].shape == (256, 319)
    assert p["view_linear.weight"].shape == (128, 283)
    q = O.unflatten_params(O.flatten_params(p))
    assert all(np.array_equal(p[k], q[k]) for k in p)


def test_positional_encoding(st):
    # sin/cos of arguments up to 2^9*6 ~ 3e3 rad: libm vs torch's vectorised sleef differ by ~1ulp
    np.testing.assert_allclose(O.positional_encoding(st["pe_x"], 10), st["pe10"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(O.positional_encoding(st["pe_x"] / 6, 4), st["pe4"], atol=1e-6, rtol=0)
    # channel order: identity first, then per frequency sin(3) then cos(3)      (model.py:22-26)
    x = st["pe_x"][:1]
    e = O.positional_encoding(x, 10)
    assert np.array_equal(e[:, :3], x)
    np.testing.assert_allclose(e[:, 3:6], np.sin(x), atol=1e-7)
    np.testing.assert_allclose(e[:, 6:9], np.cos(x), atol=1e-7)
    assert e.shape[-1] == 63 and O.positional_encoding(x, 4).shape[-1] == 27


def test_mlp_forward(st):
    p = O.init_params(7)
    out = O.mlp_forward(p, st["pe10"][:200], st["pe4"][:200])
    np.testing.assert_allclose(out, st["mlp_out"], atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("wb", [True, False])
def test_raw2outputs(st, wb):
    rgb, depth, acc, w = O.raw2outputs(st["r2o_raw"], st["r2o_z"], st["r2o_d"], white_bkgd=wb)
    np.testing.assert_allclose(w, st[f"r2o_weights_wb{int(wb)}"], atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(rgb, st[f"r2o_rgb_wb{int(wb)}"], atol=2e-6)
    np.testing.assert_allclose(depth, st[f"r2o_depth_wb{int(wb)}"], atol=1e-5)
    np.testing.assert_allclose(acc, st[f"r2o_acc_wb{int(wb)}"], atol=2e-6)
    # empty rays composite to pure background / zero; opaque rays accumulate to 1
    assert np.all(acc[:4] == 0) and np.all(np.abs(acc[4:8] - 1) < 1e-6)


def test_raw2outputs_backward(st):
    d = O.raw2outputs_backward(st["r2o_raw"], st["r2o_z"], st["r2o_d"], True, st["r2o_g_rgb"],
                               st["r2o_g_depth"], st["r2o_g_acc"], st["r2o_g_w"])
    ref = st["r2o_d_raw"]
    # autograd's cumprod backward divides by the inputs; near-opaque samples amplify rounding
    err = np.abs(d - ref)
    assert np.all(err <= 2e-4 + 2e-3 * np.abs(ref)), err.max()
    d = O.raw2outputs_backward(st["r2o_raw"], st["r2o_z"], st["r2o_d"], True, st["r2o_g_rgb"])
    ref = st["r2o_d_raw_rgbonly"]
    assert np.all(np.abs(d - ref) <= 2e-4 + 2e-3 * np.abs(ref))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sample_pdf(st, tag):
    bins, w = st[f"pdf_{tag}_bins"], st[f"pdf_{tag}_w"]
    cdf_ref = st[f"pdf_{tag}_cdf"]
    # (b) cdf within 1e-6 of the reference's
    np.testing.assert_allclose(O.pdf_to_cdf(w), cdf_ref, atol=1e-6, rtol=0)
    for mode in ("det", "rnd"):
        u = st[f"pdf_{tag}_u_{mode}"]
        uu = np.broadcast_to(u, (cdf_ref.shape[0], u.shape[-1]))
        # (a) indices bit-exact given the same cdf (pure compare/count)
        assert np.array_equal(O.searchsorted_right(cdf_ref, uu), st[f"pdf_{tag}_inds_{mode}"])
        ref = st[f"pdf_{tag}_{mode}"]
        # index + lerp stage on the reference's own cdf bits: tight everywhere
        np.testing.assert_allclose(O.sample_pdf(bins, w, u, cdf=cdf_ref), ref, atol=2e-6, rtol=0)
        # (c) end to end with the oracle's own cdf (which differs from the reference's by <=2 ulp
        # because `sum` orders differ).  Conditioning: t = (u-cdf_b)/denom, so a cdf error e moves
        # the sample by <= 2e/denom * bin_width; and where u sits within a few ulp of a cdf knot the
        # index itself may flip across a ~zero-pdf bin.  Flips are counted, not hidden.
        s, cdf, inds = O.sample_pdf(bins, w, u, return_aux=True)
        iref = st[f"pdf_{tag}_inds_{mode}"]
        lo, hi = np.maximum(iref - 1, 0), np.minimum(iref, cdf_ref.shape[-1] - 1)
        denom = np.take_along_axis(cdf_ref, hi, -1) - np.take_along_axis(cdf_ref, lo, -1)
        denom = np.where(denom < 1e-5, 1.0, denom)
        width = np.take_along_axis(bins, hi, -1) - np.take_along_axis(bins, lo, -1)
        tol = 2e-5 + 2 * 6e-7 / denom * width
        bad = np.abs(s - ref) > tol
        near_knot = (np.abs(uu[:, :, None] - cdf_ref[:, None, :]) <= 6e-7).any(-1)
        assert not np.any(bad & ~near_knot), int(np.sum(bad & ~near_knot))
        assert np.array_equal(inds[~near_knot], iref[~near_knot])
        assert bad.mean() < 5e-3
        assert np.all(s >= bins[:, :1] - 1e-6) and np.all(s <= bins[:, -1:] + 1e-6)


def test_sort_merge(st):
    out = np.sort(np.concatenate([st["merge_zc"], st["merge_zs"]], -1), -1)
    assert np.array_equal(out, st["merge_out"])


def _run_render(g):
    ns, ni, perturb, wb, cs, noise_std, seed = g["cfg"]
    cfg = O.RenderConfig(N_samples=int(ns), N_importance=int(ni), white_bkgd=bool(wb), coord_scale=float(cs))
    p = O.init_params(int(seed))
    u = None
    if ni > 0:
        u = g["u_rand"] if perturb > 0 else g["u_det"]
    return O.render_rays(p, g["rays_o"], g["rays_d"], cfg, g["t_vals"], u,
                         t_rand=g.get("t_rand"), noise_coarse=g.get("noise_coarse"),
                         noise_fine=g.get("noise_fine"))


@pytest.mark.parametrize("name", ["render_det_r96", "render_pinhole_12x12", "render_perturb_r48",
                                  "render_noise_blackbg_r32", "render_nofine_r32",
                                  "render_s128_256_r16"])
def test_render_rays_end_to_end(name):
    g = load_golden(name)
    out = _run_render(g)
    keys = [k[4:] for k in g if k.startswith("out_")]
    assert set(keys) <= set(out)
    for k in keys:
        err = np.abs(out[k] - g["out_" + k]).reshape(out[k].shape[0], -1).max(-1)
        if k.endswith("_coarse") or int(g["cfg"][1]) == 0:
            # no resampling involved: straight fp32 agreement
            assert err.max() <= 1e-5, (k, err.max())
        else:
            # Fine maps go through the inverse cdf, which is ill-conditioned at random init (many
            # coarse weights are exactly 0, so pdf bins are ~1e-4 wide and a 1-ulp cdf change
            # moves a sample by up to ~1e-3 in z).  Rule (fixed before measuring the CUDA path):
            # the 1e-4 fp32 gate must hold for >= 70 % of rays and no ray may exceed 5e-3.
            assert (err <= 1e-4).mean() >= 0.70 and err.max() <= 5e-3, (k, err.max(), (err <= 1e-4).mean())


def _check_grads(grads, g, tol):
    for k in O.PARAM_NAMES:
        gr = grads[k].reshape(-1)
        ref_norm = float(g["gnorm_" + k])
        sub = gr[::97] if gr.size > 4096 else gr
        ref = g["gsub_" + k]
        rel = np.linalg.norm(sub.astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-30)
        assert rel < tol, (k, rel)
        assert abs(np.linalg.norm(gr.astype(np.float64)) - ref_norm) <= tol * ref_norm + 1e-12, k


def test_train_grads_and_adam():
    g = load_golden("train_r32")
    p = O.init_params(int(g["seed"]))
    cfg = O.RenderConfig()
    # (1) given the reference's own z_fine: tight per-tensor relative L2
    loss, grads, out = O.train_grads(p, g["rays_o"], g["rays_d"], g["target"], cfg, g["t_vals"], g["u_det"],
                                     z_fine_override=g["z_fine"])
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    np.testing.assert_allclose(out["rgb_map"], g["rgb_map"], atol=1e-5)
    _check_grads(grads, g, 2e-4)
    # (2) end to end through the oracle's own resampling: z_fine moves by up to ~1e-3 on a few
    # samples (ill-conditioned inverse cdf), which the 2^9 positional-encoding band amplifies
    loss, grads, out = O.train_grads(p, g["rays_o"], g["rays_d"], g["target"], cfg, g["t_vals"], g["u_det"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    _check_grads(grads, g, 1e-2)
    # two Adam steps (torch.optim.Adam(lr=5e-4), scripts/train.py:258) reproduce the parameters
    flat = O.flatten_params(p)
    m = np.zeros_like(flat)
    v = np.zeros_like(flat)
    for step in (1, 2):
        _, grads, _ = O.train_grads(O.unflatten_params(flat), g["rays_o"], g["rays_d"], g["target"], cfg,
                                    g["t_vals"], g["u_det"])
        flat, m, v = O.adam_step(flat, O.flatten_params(grads), m, v, step)
        # Adam's first steps move every weight by ~lr*sign(g) regardless of |g|; where g ~ 0 the
        # sign is rounding noise, so a handful of weights may differ by up to 2*lr per step
        np.testing.assert_allclose(flat[::101], g[f"params_after_step{step}_sub"], atol=2.1e-3)
        frac_exact = np.mean(np.abs(flat[::101] - g[f"params_after_step{step}_sub"]) < 2e-5)
        assert frac_exact > 0.97, frac_exact


def test_adam_kat():
    g = load_golden("adam")
    p, m, v = g["p0"], np.zeros(1000, np.float32), np.zeros(1000, np.float32)
    for s in range(3):
        p, m, v = O.adam_step(p, g[f"g{s}"], m, v, s + 1)
        np.testing.assert_allclose(p, g[f"p{s + 1}"], atol=1e-7, rtol=1e-6)


def test_render_100x100_rows_against_reference_image():
    """BASELINE.json configs[0] (100x100 view, 64+128, fp32): the oracle on three image rows (300 rays) against
    the image the reference's own `renderer.render` produced; same rule as test_render_rays_end_to_end."""
    g = load_golden("render_pinhole_100x100")
    p = O.init_params(int(g["seed"]))
    o, d, _ = O.pinhole_rays(100, 100)
    rows = np.r_[0:100, 5000:5100, 9900:10000]
    cfg = O.RenderConfig()
    out = O.render_rays(p, o[rows], d[rows], cfg, np.linspace(0, 1, 64, dtype=np.float32), np.linspace(0, 1, 128, dtype=np.float32))
    err = np.abs(out["rgb_map"] - g["image"].reshape(-1, 3)[rows]).max(-1)
    assert (err <= 1e-4).mean() >= 0.70 and err.max() <= 5e-3, (err.max(), (err <= 1e-4).mean())


# ---- the torch-CPU restatement (oracle/nerf_oracle_torch.py): what the bench's CPU legs time ----------------
@pytest.mark.parametrize("name", ["render_det_r96", "render_pinhole_12x12", "render_perturb_r48",
                                  "render_noise_blackbg_r32", "render_nofine_r32",
                                  "render_s128_256_r16"])
def test_torch_port_render_rays(name):
    """Same torch CPU kernels in the same order as the reference => BIT-IDENTICAL maps on every golden case
    (checked with 1, 3 and 8 host threads: the partitioning of these GEMM shapes does not change the summation
    order).  This is what makes the port a valid stand-in for the reference in bench.py's CPU arm."""
    import torch
    from oracle import nerf_oracle_torch as T
    g = load_golden(name)
    ns, ni, perturb, wb, cs, noise_std, seed = g["cfg"]
    p = T.params_from_numpy(O.init_params(int(seed)))
    tt = lambda k: torch.from_numpy(g[k]) if k in g else None
    with torch.no_grad():
        out = T.render_rays(p, tt("rays_o"), tt("rays_d"), N_samples=int(ns), N_importance=int(ni),
                            white_bkgd=bool(wb), perturb=float(perturb), raw_noise_std=float(noise_std),
                            coord_scale=float(cs), t_rand=tt("t_rand"),
                            u=(tt("u_rand") if perturb > 0 else tt("u_det")) if ni > 0 else None,
                            noise_coarse=tt("noise_coarse"), noise_fine=tt("noise_fine"))
    for k in [k[4:] for k in g if k.startswith("out_")]:
        err = np.abs(out[k].numpy() - g["out_" + k])
        assert err.max() == 0.0, (name, k, float(err.max()))


def test_torch_port_train_steps():
    """Two steps of the reference's loop body (render, MSE, backward, torch.optim.Adam) reproduce the golden
    loss, gradients and parameters of the reference's own run."""
    import torch
    from oracle import nerf_oracle_torch as T
    g = load_golden("train_r32")
    tr = T.Trainer(O.init_params(int(g["seed"])), perturb=0.0)
    o, d, tgt = (torch.from_numpy(g[k]) for k in ("rays_o", "rays_d", "target"))
    for step in (1, 2):
        loss = float(tr.step(o, d, tgt, u=torch.from_numpy(g["u_det"])))
        if step == 1:
            assert abs(loss - float(g["loss"])) < 1e-6
            _check_grads(tr.grads(), g, 5e-3)
        flat = np.concatenate([tr.p[k].detach().numpy().reshape(-1) for k in O.PARAM_NAMES])
        np.testing.assert_allclose(flat[::101], g[f"params_after_step{step}_sub"], atol=2.1e-3)
        assert np.mean(np.abs(flat[::101] - g[f"params_after_step{step}_sub"]) < 2e-5) > 0.97


# ---- oracle/_ref: the unmodified reference, byte-compiled by oracle/build_ref.py (bench.py's CPU arm) ------------
def _ref_or_skip():
    from oracle import build_ref
    ok, why = build_ref.available()
    if not ok:
        pytest.skip(f"oracle/_ref not built: {why}")
    return build_ref.load()


@pytest.mark.parametrize("name", ["render_det_r96", "render_pinhole_12x12", "render_nofine_r32"])
def test_ref_build_matches_golden(name):
    """The compiled reference package (what `bench.py --impl reference` times) IS the reference: on the
    deterministic golden cases (perturb = 0, no noise: no RNG draws to replay) its maps equal the committed
    vectors bit for bit."""
    import torch
    ref = _ref_or_skip()
    g = load_golden(name)
    ns, ni, perturb, wb, cs, noise_std, seed = g["cfg"]
    assert perturb == 0 and noise_std == 0
    model = ref.NeRFMLP()
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in O.init_params(int(seed)).items()})
    r = ref.NeRFRenderer(model, "cpu", N_samples=int(ns), N_importance=int(ni), white_bkgd=bool(wb), perturb=0.0,
                         raw_noise_std=0.0, coord_scale=float(cs))
    with torch.no_grad():
        out = r._render_rays(torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]))
    for k in [k[4:] for k in g if k.startswith("out_")]:
        assert np.array_equal(out[k].numpy(), g["out_" + k]), (name, k)


def test_ref_build_is_bytecode_only():
    """oracle/_ref holds compiled modules only -- no reference source is copied into the tree -- and is git-ignored."""
    from oracle import build_ref
    _ref_or_skip()
    import zipfile
    with zipfile.ZipFile(build_ref.ARCHIVE) as z:
        files = sorted(z.namelist())
    assert files and all(f.endswith(".pyc") or f == "BUILD_INFO.txt" for f in files), files
    assert not any(f.endswith(".py") for f in os.listdir(build_ref.OUT))
    gi = open(os.path.join(os.path.dirname(build_ref.HERE), ".gitignore")).read()
    assert "oracle/_ref/" in gi

"""Size-independent properties of the CUDA path at BASELINE.json's FULL sizes (where the oracle
would take minutes): 800x800 render (640k rays, configs[2]), 1024-ray / 4096-ray training batches
(configs[1], [3]) and the 256+256 sample-count stress (configs[4]).  Needs a B200: `pytest -m gpu`.
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def nb():
    import nerf_mlp_b200
    return nerf_mlp_b200


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def model_from_seed(nb, seed, precision="bf16"):
    p = O.init_params(seed)
    m = nb.NeRFMLP(precision=precision)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    return m.to(DEV)


def test_full_800x800_render_properties(nb):
    """640k rays, 64+128 samples, bf16: chunking / sharding invariance (rays are independent, so any
    partition of the rays must give bit-identical pixels), determinism, value ranges."""
    m = model_from_seed(nb, 0)
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    o, d, focal = O.pinhole_rays(800, 800)
    to, td = T(o), T(d)
    img = r.render(to, td, 800, 800, focal)                          # reference default chunk 16384
    assert img.shape == (800, 800, 3) and img.dtype == torch.float32
    img2 = r.render(to, td, 800, 800, focal, chunk=40000)            # different chunking
    assert torch.equal(img, img2)
    # rank-sharded rendering (world sizes of the scaling run) == unsharded, bit for bit
    flat = img.view(-1, 3)
    fn = nb.dist.make_render_fn(r)
    for world in (2, 8):
        parts = []
        for rank in range(world):
            lo, hi = nb.dist.shard_range(640000, rank, world)
            parts.append(torch.cat([fn(to[i:min(i + 16384, hi)], td[i:min(i + 16384, hi)]) for i in range(lo, hi, 16384)]))
        assert torch.equal(torch.cat(parts), flat), world
    assert torch.isfinite(img).all()
    assert float(img.min()) >= -1e-5 and float(img.max()) <= 1.0 + 1e-4       # white background composite of sigmoids


def test_render_maps_invariants_large(nb):
    """65 536 rays with stratified jitter: z_fine sorted and inside [near, far], weights >= 0,
    acc = sum(weights) in [0, 1], depth inside [near*acc, far*acc], searchsorted indices == count."""
    R = 65536
    m = model_from_seed(nb, 1)
    r = nb.NeRFRenderer(m, DEV, perturb=1.0)
    o, d = O.random_rays(R, 3)
    to, td = T(o), T(d)
    torch.manual_seed(0)
    with torch.no_grad():
        t_rand = torch.rand((R, 64), device=DEV)
        z = nb.ops.stratified_z(r._linspace(64), t_rand, R, 2.0, 6.0)
        assert bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.min()) >= 2.0 and float(z.max()) <= 6.0
        rgb0, depth0, acc0, w = r._pass(to, td, z, False)
        u = torch.rand((R, 128), device=DEV)
        z_fine, zs, inds, cdf = nb.ops.resample_merge(z, w, u, check_mode=True)
        # integer work bit-exact at full size: index = #{k : cdf[k] <= u}
        cnt = (cdf.unsqueeze(1) <= u.unsqueeze(2)).sum(-1)
        assert torch.equal(inds, cnt)
        assert bool((z_fine[:, 1:] >= z_fine[:, :-1]).all())
        assert float(z_fine.min()) >= 2.0 and float(z_fine.max()) <= 6.0
        # the merge is a permutation of cat[z, z_samples]
        assert torch.equal(torch.sort(torch.cat([z, zs], -1), -1)[0], z_fine)
        assert bool((cdf[:, 1:] >= cdf[:, :-1]).all()) and float((cdf[:, -1] - 1).abs().max()) < 1e-5
        rgb, depth, acc, wf = r._pass(to, td, z_fine, False)
    for a, ww, dep in ((acc0, w, depth0), (acc, wf, depth)):
        assert float(ww.min()) >= 0.0
        assert float((ww.sum(-1) - a).abs().max()) < 1e-5
        assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0 + 1e-5
        assert bool((dep >= 2.0 * a - 1e-4).all()) and bool((dep <= 6.0 * a + 1e-4).all())


def test_gradient_is_sum_over_rays(nb):
    """Data-parallel correctness at configs[1]/[3] sizes: the gradient of a mean loss over a batch is
    the ray-count-weighted mean of the shard gradients -- what the flat all-reduce + 1/world scaling
    computes.  bf16, perturb=0, 4096 rays split in 4 shards of 1024."""
    R, W = 4096, 4
    m = model_from_seed(nb, 2)
    r = nb.NeRFRenderer(m, DEV, perturb=0.0)
    o, d = O.random_rays(R, 7)
    tgt = np.random.default_rng(8).uniform(0, 1, (R, 3)).astype(np.float32)
    to, td, tt = T(o), T(d), T(tgt)

    def grad(lo, hi):
        m.zero_grad(set_to_none=True)
        out = r._render_rays(to[lo:hi], td[lo:hi])
        nb.ops.mse_loss(out["rgb_map"], tt[lo:hi]).backward()
        return m.flat_grad.clone()

    full = grad(0, R)
    shards = sum(grad(i * R // W, (i + 1) * R // W) for i in range(W)) / W
    rel = float((full - shards).norm() / full.norm())
    # identical rows, identical arithmetic per row; only the fp32 split-K summation order differs
    assert rel < 2e-3, rel


def test_sample_count_stress_256_256(nb):
    """configs[4]: 256 coarse + 256 importance samples (512 fine samples per ray, cdf length 255):
    fp32 check mode against the oracle (stage-isolated) and bf16 against fp32 on 64 rays; plus a
    training step at that sampling."""
    R = 64
    p = O.init_params(4)
    o, d = O.random_rays(R, 9)
    outs = {}
    for prec in ("fp32", "bf16"):
        m = model_from_seed(nb, 4, prec)
        r = nb.NeRFRenderer(m, DEV, N_samples=256, N_importance=256, perturb=0.0)
        with torch.no_grad():
            z = nb.ops.stratified_z(r._linspace(256), None, R, 2.0, 6.0)
            w = r._pass(T(o), T(d), z, False)[3]
            z_fine = nb.ops.resample_merge(z, w, r._linspace(256))
            rgb, depth, acc, _ = r._pass(T(o), T(d), z_fine, False)
        outs[prec] = (rgb.cpu().numpy(), depth.cpu().numpy(), acc.cpu().numpy(), z_fine.cpu().numpy(),
                      r._linspace(256).cpu().numpy())
    rgb, depth, acc, z_fine, lin = outs["fp32"]
    cfg = O.RenderConfig(N_samples=256, N_importance=256)
    ref = O.render_rays(p, o, d, cfg, lin, lin, z_fine_override=z_fine)
    assert z_fine.shape == (R, 512)
    np.testing.assert_allclose(rgb, ref["rgb_map"], atol=1e-4)
    np.testing.assert_allclose(depth, ref["depth_map"], atol=1e-4)
    np.testing.assert_allclose(acc, ref["acc_map"], atol=1e-4)
    keep = np.abs(ref["raw_fine"][:, -1, 3]) >= 4e-3
    assert np.abs(outs["bf16"][0] - rgb)[keep].max() < 1e-2
    # one training step at 256+256
    m = model_from_seed(nb, 4, "bf16")
    r = nb.NeRFRenderer(m, DEV, N_samples=256, N_importance=256, perturb=1.0)
    opt = nb.FlatAdam(m, lr=5e-4)
    loss = nb.ops.mse_loss(r._render_rays(T(o), T(d))["rgb_map"], torch.rand(R, 3, device=DEV))
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert np.isfinite(float(loss)) and bool(torch.isfinite(m.flat_params).all())


def test_training_step_determinism_and_repack(nb):
    """The bf16 weight image follows the parameters: after optimizer.step() (torch.optim.Adam on the
    24 views, i.e. the reference's own training loop) the next forward uses the new weights; two
    identical runs give identical losses (no atomics on the forward path)."""
    losses = []
    for _ in range(2):
        m = model_from_seed(nb, 11)      # (seed 5 is a dead init: sigma <= 0 everywhere, zero gradient)
        r = nb.NeRFRenderer(m, DEV, perturb=0.0)
        opt = torch.optim.Adam(m.parameters(), lr=5e-4)
        o, d = O.random_rays(256, 111)
        tgt = torch.rand(256, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
        ls = []
        for _ in range(3):
            loss = torch.mean((r._render_rays(T(o), T(d))["rgb_map"] - tgt) ** 2)
            opt.zero_grad()
            loss.backward()
            opt.step()
            ls.append(float(loss))
        losses.append(ls)
    assert losses[0][1] < losses[0][0]             # the second forward saw the updated weights
    assert abs(losses[0][0] - losses[1][0]) < 1e-7  # forward is deterministic
    assert abs(losses[0][2] - losses[1][2]) < 1e-3  # backward uses fp32 atomics (order may differ)

#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (dgsmith7/nerf-mlp) on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so these vectors --
outputs of the reference's own nerfmlp.NeRFMLP / nerfmlp.NeRFRenderer on seeded inputs -- are
the parity pin for oracle/nerf_oracle.py and, through it, for the CUDA path.  Weights come from
oracle.init_params(seed) (numpy Generator, reproducible anywhere) and are loaded into the
reference model through load_state_dict, so no weight tensors need to be committed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from nerfmlp import NeRFMLP, NeRFRenderer            # noqa: E402  (the reference)
from nerfmlp.model import PositionalEncoding          # noqa: E402
from oracle import nerf_oracle as O                   # noqa: E402

torch.set_num_threads(os.cpu_count())
DEV = torch.device("cpu")


def ref_model(seed):
    p = O.init_params(seed)
    m = NeRFMLP()
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    return m, p


class Recorder:
    """Records torch.rand / torch.randn_like draws made inside the reference so that the oracle
    and the CUDA path can be fed the identical numbers."""

    def __init__(self):
        self.rand, self.randn = [], []

    def __enter__(self):
        self._rand, self._randn_like = torch.rand, torch.randn_like

        def rand(*a, **k):
            t = self._rand(*a, **k)
            self.rand.append(t.clone())
            return t

        def randn_like(*a, **k):
            t = self._randn_like(*a, **k)
            self.randn.append(t.clone())
            return t

        torch.rand, torch.randn_like = rand, randn_like
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn_like = self._rand, self._randn_like


def render_case(name, R, seed, *, N_samples=64, N_importance=128, perturb=0.0, white_bkgd=True,
                coord_scale=1.0, raw_noise_std=0.0, pinhole=None):
    model, _ = ref_model(seed)
    r = NeRFRenderer(model, DEV, N_samples=N_samples, N_importance=N_importance, near=2.0, far=6.0,
                     white_bkgd=white_bkgd, perturb=perturb, raw_noise_std=raw_noise_std,
                     coord_scale=coord_scale)
    if pinhole:
        o, d, _ = O.pinhole_rays(*pinhole)
    else:
        o, d = O.random_rays(R, seed + 100)
    to, td = torch.from_numpy(o), torch.from_numpy(d)
    torch.manual_seed(seed)
    with Recorder() as rec, torch.no_grad():
        out = r._render_rays(to, td)
    g = dict(rays_o=o, rays_d=d,
             t_vals=torch.linspace(0., 1., steps=N_samples).numpy(),
             cfg=np.array([N_samples, N_importance, perturb, white_bkgd, coord_scale, raw_noise_std,
                           seed], np.float64))
    if N_importance > 0:
        g["u_det"] = torch.linspace(0., 1., N_importance).numpy()
    it = iter(rec.rand)
    if perturb > 0:
        g["t_rand"] = next(it).numpy()
        if N_importance > 0:
            g["u_rand"] = next(it).numpy()
    if raw_noise_std > 0:
        g["noise_coarse"] = (rec.randn[0] * raw_noise_std).numpy()
        if N_importance > 0:
            g["noise_fine"] = (rec.randn[1] * raw_noise_std).numpy()
    for k, v in out.items():
        g["out_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
    print(name, {k: v.shape for k, v in g.items() if k.startswith("out_")})


def stage_case():
    """Stage-level vectors: PE, MLP forward, _raw2outputs, _sample_pdf (+cdf, inds), sort-merge."""
    seed = 7
    model, _ = ref_model(seed)
    rng = np.random.default_rng(seed)
    g = {}
    # positional encoding (model.py:20-26), xyz range of real scenes incl. |x|~6
    x = (rng.uniform(-6, 6, (257, 3))).astype(np.float32)
    g["pe_x"] = x
    g["pe10"] = PositionalEncoding(10)(torch.from_numpy(x)).numpy()
    g["pe4"] = PositionalEncoding(4)(torch.from_numpy(x / 6)).numpy()
    # MLP forward on encoded inputs (model.py:57-81)
    xe = g["pe10"][:200]
    de = g["pe4"][:200]
    with torch.no_grad():
        g["mlp_out"] = model(torch.from_numpy(xe), torch.from_numpy(de)).numpy()
    # _raw2outputs (renderer.py:114-163): wide raw range so alpha spans (0,1); non-unit dirs
    R, S = 48, 192
    raw = (rng.standard_normal((R, S, 4)) * np.array([2, 2, 2, 6])).astype(np.float32)
    raw[:4, :, 3] = -1.0            # empty rays: sigma <= 0 everywhere
    raw[4:8, :, 3] = 50.0           # opaque rays: alpha -> 1 at the first sample
    z = np.sort(rng.uniform(2, 6, (R, S)).astype(np.float32), -1)
    z[8] = np.linspace(2, 6, S, dtype=np.float32)
    z[9, 10:20] = z[9, 10]          # repeated depths -> zero dists
    d = rng.standard_normal((R, 3)).astype(np.float32) * 2
    g["r2o_raw"], g["r2o_z"], g["r2o_d"] = raw, z, d
    for wb in (True, False):
        r = NeRFRenderer(model, DEV, white_bkgd=wb)
        outs = r._raw2outputs(torch.from_numpy(raw), torch.from_numpy(z), torch.from_numpy(d))
        for nme, t in zip(("rgb", "depth", "acc", "weights"), outs):
            g[f"r2o_{nme}_wb{int(wb)}"] = t.numpy()
    # its autograd gradient w.r.t. raw for random upstream grads on all four outputs
    r = NeRFRenderer(model, DEV, white_bkgd=True)
    traw = torch.from_numpy(raw).clone().requires_grad_(True)
    rgb_m, dep_m, acc_m, w_m = r._raw2outputs(traw, torch.from_numpy(z), torch.from_numpy(d))
    gr = [rng.standard_normal(t.shape).astype(np.float32) for t in (rgb_m, dep_m, acc_m, w_m)]
    (rgb_m * torch.from_numpy(gr[0])).sum().add((dep_m * torch.from_numpy(gr[1])).sum()) \
        .add((acc_m * torch.from_numpy(gr[2])).sum()).add((w_m * torch.from_numpy(gr[3])).sum()).backward()
    g["r2o_g_rgb"], g["r2o_g_depth"], g["r2o_g_acc"], g["r2o_g_w"] = gr
    g["r2o_d_raw"] = traw.grad.numpy()
    # rgb-only upstream (the training case)
    traw = torch.from_numpy(raw).clone().requires_grad_(True)
    rgb_m = r._raw2outputs(traw, torch.from_numpy(z), torch.from_numpy(d))[0]
    (rgb_m * torch.from_numpy(gr[0])).sum().backward()
    g["r2o_d_raw_rgbonly"] = traw.grad.numpy()

    # _sample_pdf (renderer.py:165-199).  The reference computes cdf/inds internally; recompute
    # them with the same torch calls to export them.
    def ref_cdf_inds(w, u):
        w = w + 1e-5
        pdf = w / torch.sum(w, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        return cdf, torch.searchsorted(cdf, u.contiguous(), right=True)

    for tag, nb, nimp in (("a", 63, 128), ("b", 255, 256)):
        Rr = 40
        bins = np.sort(rng.uniform(2, 6, (Rr, nb)).astype(np.float32), -1)
        w = rng.uniform(0, 1, (Rr, nb - 1)).astype(np.float32) ** 4
        w[0] = 0.0                   # all-zero weights -> uniform pdf from the 1e-5 floor
        w[1] = 0.0
        w[1, 17] = 1.0               # single spike
        w[2, :] = 1.0                # exactly uniform
        w[3, : (nb - 1) // 2] = 0.0  # long run of (near-)zero pdf -> denom < 1e-5 branch
        g[f"pdf_{tag}_bins"], g[f"pdf_{tag}_w"] = bins, w
        tb, tw = torch.from_numpy(bins), torch.from_numpy(w)
        r = NeRFRenderer(model, DEV)
        g[f"pdf_{tag}_det"] = r._sample_pdf(tb, tw, nimp, det=True).numpy()
        u_det = torch.linspace(0., 1., nimp)
        g[f"pdf_{tag}_u_det"] = u_det.numpy()
        cdf, inds = ref_cdf_inds(tw, u_det.expand(Rr, nimp))
        g[f"pdf_{tag}_cdf"], g[f"pdf_{tag}_inds_det"] = cdf.numpy(), inds.numpy()
        torch.manual_seed(3)
        with Recorder() as rec:
            g[f"pdf_{tag}_rnd"] = r._sample_pdf(tb, tw, nimp, det=False).numpy()
        g[f"pdf_{tag}_u_rnd"] = rec.rand[0].numpy()
        g[f"pdf_{tag}_inds_rnd"] = ref_cdf_inds(tw, rec.rand[0])[1].numpy()
    # sort-merge (renderer.py:90)
    zc = np.sort(rng.uniform(2, 6, (16, 64)).astype(np.float32), -1)
    zs = np.sort(rng.uniform(2, 6, (16, 128)).astype(np.float32), -1)
    zs[0, :64] = zc[0]               # exact ties
    g["merge_zc"], g["merge_zs"] = zc, zs
    g["merge_out"] = torch.sort(torch.cat([torch.from_numpy(zc), torch.from_numpy(zs)], -1), -1)[0].numpy()
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **g)
    print("stages", len(g), "arrays")


def grad_case(name="train_r32", R=32, seed=11):
    """Training-step gradient (scripts/train.py:374-382) with perturb=0, plus 2 Adam steps."""
    model, p0 = ref_model(seed)
    r = NeRFRenderer(model, DEV, N_samples=64, N_importance=128, perturb=0.0)
    o, d = O.random_rays(R, seed + 100)
    rng = np.random.default_rng(seed + 1)
    tgt = rng.uniform(0, 1, (R, 3)).astype(np.float32)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    g = dict(rays_o=o, rays_d=d, target=tgt, t_vals=torch.linspace(0., 1., 64).numpy(),
             u_det=torch.linspace(0., 1., 128).numpy(), seed=np.array(seed))
    for step in range(2):
        out = r._render_rays(torch.from_numpy(o), torch.from_numpy(d))
        loss = torch.mean((out["rgb_map"] - torch.from_numpy(tgt)) ** 2)
        opt.zero_grad()
        loss.backward()
        if step == 0:
            with torch.no_grad():   # the reference's own z_fine, recomputed with its own calls
                to, td = torch.from_numpy(o), torch.from_numpy(d)
                t = torch.linspace(0., 1., steps=64)
                z = (2.0 * (1. - t) + 6.0 * t).expand([R, 64])
                pts = (to.unsqueeze(1) + td.unsqueeze(1) * z.unsqueeze(2)).reshape(-1, 3)
                vd = td / (td.norm(dim=-1, keepdim=True) + 1e-8)
                raw = model(r.pos_enc(pts), r.dir_enc(vd).unsqueeze(1).expand(-1, 64, -1).reshape(-1, 27))
                w = r._raw2outputs(raw.view(R, 64, 4), z, td)[3]
                zs = r._sample_pdf(0.5 * (z[..., 1:] + z[..., :-1]), w[..., 1:-1], 128, det=True)
                g["z_fine"] = torch.sort(torch.cat([z, zs], -1), -1)[0].numpy()
            g["loss"] = np.array(loss.item(), np.float32)
            g["rgb_map"] = out["rgb_map"].detach().numpy()
            for k, prm in model.named_parameters():
                gr = prm.grad.numpy().reshape(-1)
                g["gnorm_" + k] = np.array(np.linalg.norm(gr.astype(np.float64)))
                g["gsub_" + k] = gr[::97].copy() if gr.size > 4096 else gr.copy()
        opt.step()
        flat = np.concatenate([prm.detach().numpy().reshape(-1) for _, prm in model.named_parameters()])
        g[f"params_after_step{step + 1}_sub"] = flat[::101].copy()
        g[f"loss_step{step}"] = np.array(loss.item(), np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
    print(name, float(g["loss"]))


def adam_case():
    rng = np.random.default_rng(5)
    p = torch.nn.Parameter(torch.from_numpy(rng.standard_normal(1000).astype(np.float32)))
    opt = torch.optim.Adam([p], lr=5e-4)
    g = {"p0": p.detach().numpy().copy()}
    for s in range(3):
        gr = rng.standard_normal(1000).astype(np.float32) * (10.0 ** (s - 2))
        p.grad = torch.from_numpy(gr.copy())
        opt.step()
        g[f"g{s}"] = gr
        g[f"p{s + 1}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "adam.npz"), **g)
    print("adam ok")


def data_case(name="data_3x12x12"):
    """nerfmlp.NeRFDataset (data.py:24-104) on a tiny synthetic Blender-style scene written to a temp
    dir, and the output post-processing of scripts/render_example.py (:12-26, :256-271)."""
    import json
    import tempfile
    from PIL import Image
    from nerfmlp import NeRFDataset
    sys.path.insert(0, "/root/reference/scripts")
    import render_example as RE
    rng = np.random.default_rng(77)
    N, S = 3, 12
    angle = 0.6911112070083618
    rgba = rng.integers(0, 256, (N, S, S, 4), dtype=np.uint8)
    rgba[0, :2, :, 3] = 255          # some fully opaque / fully transparent pixels, some dark ones (linear branch)
    rgba[1, :2, :, 3] = 0
    rgba[2, :3, :, :3] = rng.integers(0, 12, (3, S, 3), dtype=np.uint8)
    poses = []
    for _ in range(N):
        q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        m = np.eye(4)
        m[:3, :3] = q
        m[:3, 3] = rng.standard_normal(3) * 2.0
        poses.append(m.astype(np.float32))
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "train"))
        frames = []
        for i in range(N):
            Image.fromarray(rgba[i], "RGBA").save(os.path.join(d, "train", f"r_{i}.png"))
            frames.append({"file_path": f"./train/r_{i}", "transform_matrix": poses[i].astype(np.float64).tolist()})
        json.dump({"camera_angle_x": angle, "frames": frames}, open(os.path.join(d, "transforms_train.json"), "w"))
        ds = NeRFDataset(d, split="train", img_wh=(S, S), white_bkgd=True)
    idx = rng.permutation(len(ds))[:64]
    items = [ds[int(i)] for i in idx]
    g = dict(rgba=rgba, poses=np.stack(poses), camera_angle_x=np.array(angle), focal=np.array(ds.focal, np.float64),
             all_rays_o=torch.from_numpy(np.ascontiguousarray(ds.all_rays_o)).float().numpy(),
             all_rays_d=torch.from_numpy(np.ascontiguousarray(ds.all_rays_d)).float().numpy(),
             all_rgbs=torch.from_numpy(np.ascontiguousarray(ds.all_rgbs)).float().numpy(),
             idx=idx.astype(np.int64),
             batch_ray_o=torch.stack([it["ray_o"] for it in items]).numpy(),
             batch_ray_d=torch.stack([it["ray_d"] for it in items]).numpy(),
             batch_rgb=torch.stack([it["rgb"] for it in items]).numpy())
    # post-processing, with the reference's own statements (render_example.py:256-271)
    lin = rng.uniform(0.0, 1.3, (40, 17, 3)).astype(np.float32)
    lin[0, :, :] = np.linspace(0, 0.006, 17, dtype=np.float32)[:, None]      # around the 0.0031308 branch point
    g["pp_in"] = lin
    for boost in (1.0, 1.5):
        for gamma in (False, True):
            rgb = lin.copy()
            if boost != 1.0:
                rgb = rgb * boost
            rgb_final = RE.linear_to_srgb(rgb) if gamma else rgb
            g[f"pp_out_b{boost}_g{int(gamma)}"] = (np.clip(rgb_final, 0, 1) * 255).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
    print(name, len(ds), g["all_rays_d"].dtype, g["all_rgbs"].dtype)


def render100_case(name="render_pinhole_100x100", seed=8):
    """BASELINE.json configs[0]: the reference's own `renderer.render` of a 100x100 view (10 000 rays, 64+128
    samples, fp32, CPU), through the public entry point with its default chunking.  Only the image is kept
    (float16-free: fp32, 120 KB)."""
    model, _ = ref_model(seed)
    r = NeRFRenderer(model, DEV, N_samples=64, N_importance=128, near=2.0, far=6.0, white_bkgd=True, perturb=0.0)
    o, d, focal = O.pinhole_rays(100, 100)
    img = r.render(torch.from_numpy(o), torch.from_numpy(d), 100, 100, focal)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=np.array(seed), image=img.numpy())
    print(name, img.shape, float(img.mean()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "data":
        data_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "render100":
        render100_case()
        sys.exit(0)
    stage_case()
    render_case("render_det_r96", 96, 1)                                   # config-1 shape, small R
    render_case("render_pinhole_12x12", 144, 2, pinhole=(12, 12))          # render_example.py rays
    render_case("render_perturb_r48", 48, 3, perturb=1.0)                  # training sampling
    render_case("render_noise_blackbg_r32", 32, 4, perturb=1.0, white_bkgd=False, raw_noise_std=1.0,
                coord_scale=0.5)
    render_case("render_nofine_r32", 32, 5, N_importance=0)
    render_case("render_s128_256_r16", 16, 6, N_samples=128, N_importance=256)   # config-5 style
    grad_case()
    adam_case()
    data_case()
    render100_case()

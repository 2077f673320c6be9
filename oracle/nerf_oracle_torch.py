"""Second CPU oracle: the hot path restated on torch CPU ops -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (dgsmith7/nerf-mlp) keeps all of its arithmetic inside PyTorch,
and PyTorch -- unlike the reference checkout -- exists on the GPU box.  This module restates the path
as plain functions over a dict of the 24 parameter tensors, calling the same torch CPU kernels in
the same order as ``nerfmlp/model.py`` / ``nerfmlp/renderer.py`` (citations ``file:line`` into
``/root/reference``), with autograd for the backward and ``torch.optim.Adam`` for the update exactly
as ``scripts/train.py:258,381-387`` does.  It is therefore the closest thing to "the reference's own
CPU implementation" that can travel: it is what ``bench.py --impl reference`` and the ``cpu_baseline``
leg time on the box's host cores (the numpy port in ``nerf_oracle.py`` stays the bit-level checker
of the CUDA kernels; it is ~3x slower than torch's CPU GEMMs and would flatter the GPU/CPU ratio).

Only ``tests/`` and ``bench.py``'s CPU legs import it; nothing under ``nerf_mlp_b200/`` does.
Parity pin: ``tests/test_oracle_golden.py::test_torch_port_*`` check it against the golden vectors
that ``tests/golden/make_golden.py`` generated from the unmodified reference (same ops, so the
agreement is to the last bit or ulp where the reference's own run-to-run thread partitioning allows).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .nerf_oracle import LAYER_SHAPES, PARAM_NAMES


def params_from_numpy(p: dict, requires_grad: bool = False, device="cpu") -> dict:
    """numpy parameter dict (nerf_oracle.init_params) -> fp32 tensors on `device`, same 24 names.
    (`device="cuda"` runs the same restatement as eager fp32 PyTorch on the GPU: the second oracle of
    the `-m gpu` parity tests at sizes the CPU cannot finish in seconds)"""
    return {k: torch.from_numpy(p[k].copy()).to(device).requires_grad_(requires_grad) for k in PARAM_NAMES}


def positional_encoding(x: torch.Tensor, num_freqs: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ...], no pi factor (model.py:14-16, 20-26)."""
    bands = 2.0 ** torch.linspace(0.0, num_freqs - 1, num_freqs)
    parts = [x]
    for f in bands:
        parts += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(parts, dim=-1)


def mlp_forward(p: dict, x: torch.Tensor, viewdirs: torch.Tensor) -> torch.Tensor:
    """8 x 256 ReLU trunk with the [x, h] concat before layer 5, sigma / bottleneck / view / rgb heads,
    output [rgb, sigma] (model.py:57-81)."""
    h = x
    for i in range(8):
        if i == 5:
            h = torch.cat([x, h], -1)
        h = F.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
    sigma = F.linear(h, p["sigma_linear.weight"], p["sigma_linear.bias"])
    b = F.linear(h, p["bottleneck_linear.weight"], p["bottleneck_linear.bias"])
    hv = F.relu(F.linear(torch.cat([b, viewdirs], -1), p["view_linear.weight"], p["view_linear.bias"]))
    rgb = F.linear(hv, p["rgb_linear.weight"], p["rgb_linear.bias"])
    return torch.cat([rgb, sigma], -1)


def raw2outputs(raw, z_vals, rays_d, white_bkgd=True, noise=None):
    """Volume compositing (renderer.py:114-163)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    dens = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-F.relu(dens) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights.unsqueeze(-1) * rgb, dim=-2)
    depth_map = torch.sum(weights * z_vals, dim=-1)
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map.unsqueeze(-1))
    return rgb_map, depth_map, acc_map, weights


def sample_pdf(bins, weights, u):
    """Inverse-cdf sampling (renderer.py:165-199); `u` is [N] (shared, the det case) or [R, N]."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u.dim() == 1:
        u = u.expand(list(cdf.shape[:-1]) + [u.shape[0]])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, 0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b)


def _eval_samples(p, rays_o, rays_d, z, coord_scale):
    """points -> encodings -> network for one pass (renderer.py:63-77 / :91-104)."""
    R, S = z.shape
    pts = (rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(2)).reshape(-1, 3)
    if coord_scale != 1.0:
        pts = pts * coord_scale
    x = positional_encoding(pts, 10)
    vd = rays_d / (rays_d.norm(dim=-1, keepdim=True) + 1e-8)
    de = positional_encoding(vd, 4)
    de = de.unsqueeze(1).expand(-1, S, -1).reshape(-1, de.shape[-1])
    return mlp_forward(p, x, de).view(R, S, 4)


def render_rays(p, rays_o, rays_d, N_samples=64, N_importance=128, near=2.0, far=6.0, white_bkgd=True,
                perturb=0.0, raw_noise_std=0.0, coord_scale=1.0, t_rand=None, u=None,
                noise_coarse=None, noise_fine=None, z_fine_override=None):
    """Coarse + fine pass of one ray batch (renderer.py:47-112).  Random draws may be supplied
    (t_rand [R,S_c], u [R,N_imp], noise_*), otherwise they are drawn in the reference's order.
    `z_fine_override` (tests): continue the fine pass from given depths -- isolates the stages after the
    ill-conditioned inverse cdf, as the numpy oracle's option of the same name does."""
    R = rays_o.shape[0]
    dev = rays_o.device
    t = torch.linspace(0.0, 1.0, steps=N_samples, device=dev)
    z = (near * (1.0 - t) + far * t).expand([R, N_samples])
    if perturb > 0:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(z.shape, device=dev)
        z = lower + (upper - lower) * t_rand
    raw = _eval_samples(p, rays_o, rays_d, z, coord_scale)
    if raw_noise_std > 0 and noise_coarse is None:
        noise_coarse = torch.randn_like(raw[..., 3]) * raw_noise_std
    rgb0, depth0, acc0, w = raw2outputs(raw, z, rays_d, white_bkgd, noise_coarse)
    if N_importance <= 0:
        return {"rgb_map": rgb0, "depth_map": depth0, "acc_map": acc0}
    z_mid = 0.5 * (z[..., 1:] + z[..., :-1])
    if u is None:
        u = torch.linspace(0.0, 1.0, N_importance, device=dev) if perturb == 0.0 else torch.rand([R, N_importance], device=dev)
    z_samples = sample_pdf(z_mid, w[..., 1:-1], u).detach()
    z_fine, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
    if z_fine_override is not None:
        z_fine = z_fine_override
    raw_f = _eval_samples(p, rays_o, rays_d, z_fine, coord_scale)
    if raw_noise_std > 0 and noise_fine is None:
        noise_fine = torch.randn_like(raw_f[..., 3]) * raw_noise_std
    rgb, depth, acc, w_f = raw2outputs(raw_f, z_fine, rays_d, white_bkgd, noise_fine)
    return {"rgb_map": rgb, "depth_map": depth, "acc_map": acc,
            "rgb_map_coarse": rgb0, "depth_map_coarse": depth0, "acc_map_coarse": acc0,
            "z_fine": z_fine, "raw_fine": raw_f.detach(), "weights_fine": w_f.detach(), "weights_coarse": w.detach()}


class Trainer:
    """The reference's loop body on CPU (scripts/train.py:258-260, 374-388): render the batch,
    MSE on the fine rgb_map, backward, Adam(lr=5e-4)."""

    def __init__(self, p_numpy: dict, lr: float = 5e-4, device="cpu", **render_kw):
        self.p = params_from_numpy(p_numpy, requires_grad=True, device=device)
        self.opt = torch.optim.Adam([self.p[k] for k in PARAM_NAMES], lr=lr)
        self.render_kw = render_kw

    def step(self, rays_o, rays_d, target, **draws):
        out = render_rays(self.p, rays_o, rays_d, **self.render_kw, **draws)
        loss = torch.mean((out["rgb_map"] - target) ** 2)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.detach()

    def grads(self) -> dict:
        return {k: self.p[k].grad.detach().cpu().numpy() for k in PARAM_NAMES}


assert len(PARAM_NAMES) == 2 * len(LAYER_SHAPES)

"""Drop-in for the reference's nerfmlp/renderer.py: NeRFRenderer.

Same constructor, public attributes and methods as the reference (renderer.py:5-199).  The host
code below only orchestrates; every tensor op of the reference is one of libnerf_b200's kernels:

    renderer.py:52-61    z_vals / stratified jitter      -> nerf_stratified_z
    renderer.py:63-77    points, PE, view dirs, MLP      -> nerf_mlp_fwd_rays   (encodings in-register)
    renderer.py:79-80    _raw2outputs                    -> nerf_composite_fwd / _bwd
    renderer.py:86-90    z_mid, _sample_pdf, sort-merge  -> nerf_resample_merge
    renderer.py:91-107   fine pass                       -> the same two kernels

torch is used for memory, streams, RNG draws (so a seeded reference run on the same device sees
the same random stream, in the same order: renderer.py:60 -> :136 coarse -> :182 -> :136 fine)
and autograd bookkeeping.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .model import NeRFMLP, PositionalEncoding, _PRECISIONS


class NeRFRenderer:
    def __init__(self, model, device,
                 pos_enc_L=10, dir_enc_L=4,
                 N_samples=64, N_importance=128,
                 near=2.0, far=6.0, white_bkgd=True, perturb=1.0, raw_noise_std=0.0, coord_scale=1.0,
                 *, precision=None, coarse_grad=None, coarse_density_only=True):
        if not isinstance(model, NeRFMLP):
            raise TypeError("nerf_mlp_b200.NeRFRenderer needs a nerf_mlp_b200.NeRFMLP (the fused kernels "
                            "own the network; there is no generic nn.Module path)")
        if (pos_enc_L, dir_enc_L) != (10, 4):
            raise NotImplementedError("the fused MLP kernels implement pos_enc_L=10 / dir_enc_L=4 only")
        self.model = model
        self.device = torch.device(device)
        self.N_samples = N_samples
        self.N_importance = N_importance
        self.near = near
        self.far = far
        self.white_bkgd = white_bkgd
        self.perturb = perturb
        self.raw_noise_std = raw_noise_std
        self.coord_scale = coord_scale
        self.pos_enc = PositionalEncoding(pos_enc_L).to(device)
        self.dir_enc = PositionalEncoding(dir_enc_L).to(device)
        # extras (keyword-only, defaults keep the reference's behaviour)
        self.precision = precision            # None -> follow model.precision
        # None (default) = the reference's behaviour at no cost: under autograd the coarse maps ARE differentiable
        # (renderer.py:79-80 builds them with grad, so a caller who adds the usual coarse MSE term gets its gradient),
        # but lazily -- the coarse forward runs the inference kernel and is re-run in save mode only if a gradient
        # actually reaches it (ops.RenderPassFn save="lazy").  True saves eagerly (and makes TrainStep refuse: its fused
        # step implements the reference's fine-only loss).  False is the explicit opt-out (*_coarse come back detached).
        self.coarse_grad = coarse_grad
        # render() / TrainStep: evaluate the coarse pass for its densities only (its colour maps are dropped by
        # render(), renderer.py:44, and unused by the loss); False = the whole network in both passes
        self.coarse_density_only = coarse_density_only
        self._lin = {}

    # ---- helpers -------------------------------------------------------------------------------
    def _prec(self):
        return _PRECISIONS[self.precision or self.model.precision]

    def _linspace(self, n):
        """torch.linspace(0,1,n) on the device, cached (never recomputed in-kernel)."""
        key = (n, self.device)
        t = self._lin.get(key)
        if t is None:
            t = torch.linspace(0., 1., steps=n, device=self.device)
            self._lin[key] = t
        return t

    def _pass(self, rays_o, rays_d, z_vals, want_grad, density_only=False):
        """want_grad: False | True (save activations now) | "lazy" (differentiable, recomputed on demand)."""
        noise = None
        if self.raw_noise_std > 0.:
            noise = (torch.randn(z_vals.shape, device=z_vals.device) * self.raw_noise_std).contiguous()  # :134-136
        m = self.model
        if want_grad:
            return ops.RenderPassFn.apply(m, rays_o, rays_d, z_vals, noise, bool(self.white_bkgd),
                                          float(self.coord_scale), self._prec(), want_grad, *m._param_list)
        with torch.no_grad():
            return ops.RenderPassFn.apply(m, rays_o, rays_d, z_vals, noise, bool(self.white_bkgd),
                                          float(self.coord_scale), self._prec(), "density" if density_only else False,
                                          *m._param_list)

    # ---- reference API -------------------------------------------------------------------------
    def render(self, rays_o, rays_d, H, W, focal, chunk=1024 * 16):
        """rays_o, rays_d: (N_rays, 3) -> (H, W, 3) image; chunked, no grad (renderer.py:23-45).
        `focal` is unused, as in the reference."""
        N_rays = rays_o.shape[0]
        results = []
        for i in range(0, N_rays, chunk):
            with torch.no_grad():
                # only the fine rgb_map leaves this function (renderer.py:44): the coarse pass is evaluated for
                # its densities alone (same weights, hence bit-identical fine samples and fine maps)
                results.append(self._render_rays(rays_o[i:i + chunk], rays_d[i:i + chunk],
                                                 _coarse_density_only=self.coarse_density_only)['rgb_map'])
        rgb_map = torch.cat(results, 0)
        return rgb_map.view(H, W, 3)

    def render_maps(self, rays_o, rays_d, H, W, focal, chunk=1024 * 16):
        """`render` that keeps everything `_render_rays` returns (the reference's `render` drops the
        depth / acc / coarse maps, renderer.py:44): dict of (H, W, 3) / (H, W) tensors, no grad."""
        parts = []
        for i in range(0, rays_o.shape[0], chunk):
            with torch.no_grad():
                parts.append(self._render_rays(rays_o[i:i + chunk], rays_d[i:i + chunk]))
        out = {}
        for k in parts[0]:
            v = torch.cat([p[k] for p in parts], 0)
            out[k] = v.view(H, W, 3) if v.dim() == 2 else v.view(H, W)
        return out

    def _render_rays(self, rays_o, rays_d, _coarse_density_only=False):
        """reference renderer.py:47-112.  `_coarse_density_only` (internal, used by render()): the coarse
        colour/depth/acc maps are not needed by the caller, so the coarse pass computes densities only and
        the returned dict has no *_coarse entries."""
        rays_o = _lib.f32c(rays_o)
        rays_d = _lib.f32c(rays_d)
        N_rays = rays_o.shape[0]
        self.model._ensure_flat()
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.model._param_list)

        # === coarse sampling (renderer.py:52-61) ===
        t_rand = None
        if self.perturb > 0:
            t_rand = torch.rand((N_rays, self.N_samples), device=self.device)                    # :60
        z_vals = ops.stratified_z(self._linspace(self.N_samples), t_rand, N_rays, self.near, self.far)
        fine = self.N_importance > 0
        density_only = bool(_coarse_density_only) and fine and not grad
        if not grad or (fine and self.coarse_grad is False):
            coarse_mode = False
        elif fine and self.coarse_grad is None:
            coarse_mode = "lazy"
        else:
            coarse_mode = True
        rgb0, depth0, acc0, weights = self._pass(rays_o, rays_d, z_vals, coarse_mode, density_only)  # :63-80
        if not fine:
            return {'rgb_map': rgb0, 'depth_map': depth0, 'acc_map': acc0}                       # :112

        # === hierarchical sampling (renderer.py:86-90); z_samples is detached in the reference ===
        if self.perturb == 0.:
            u = self._linspace(self.N_importance)                                                # :179
        else:
            u = torch.rand((N_rays, self.N_importance), device=self.device)                      # :182
        z_fine = ops.resample_merge(z_vals, weights, u)

        # === fine pass (renderer.py:91-107), same network ===
        rgb, depth, acc, _ = self._pass(rays_o, rays_d, z_fine, grad)
        if density_only:
            return {'rgb_map': rgb, 'depth_map': depth, 'acc_map': acc}
        return {'rgb_map': rgb, 'depth_map': depth, 'acc_map': acc,
                'rgb_map_coarse': rgb0, 'depth_map_coarse': depth0, 'acc_map_coarse': acc0}

    def _raw2outputs(self, raw, z_vals, rays_d):
        """renderer.py:114-163 -> (rgb_map, depth_map, acc_map, weights)."""
        raw, z_vals, rays_d = _lib.f32c(raw), _lib.f32c(z_vals), _lib.f32c(rays_d)
        noise = None
        if self.raw_noise_std > 0.:
            noise = (torch.randn_like(raw[..., 3]) * self.raw_noise_std).contiguous()
        if raw.requires_grad and torch.is_grad_enabled():
            return ops.CompositeFn.apply(raw, z_vals, rays_d, noise, bool(self.white_bkgd))
        return ops.composite_fwd(raw.detach(), z_vals, rays_d, noise, bool(self.white_bkgd), True)

    def _sample_pdf(self, bins, weights, N_samples, det=False):
        """renderer.py:165-199."""
        bins = bins if (bins.is_cuda and bins.dtype == torch.float32 and bins.stride(-1) == 1) else _lib.f32c(bins)
        weights = weights if (weights.is_cuda and weights.dtype == torch.float32 and weights.stride(-1) == 1) \
            else _lib.f32c(weights)
        if det:
            u = self._linspace(N_samples) if bins.device == self.device else \
                torch.linspace(0., 1., N_samples, device=bins.device)
        else:
            u = torch.rand(list(bins.shape[:-1]) + [N_samples], device=bins.device)
        return ops.sample_pdf(bins.detach(), weights.detach(), u)

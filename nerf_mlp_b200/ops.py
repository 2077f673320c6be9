"""Thin Python wrappers over the C ABI (one function per kernel entry point) and the
torch.autograd.Function glue that lets ``loss.backward()`` in the reference's training scripts
(scripts/train.py:381-382) reach the fused backward kernels unchanged."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import PREC_BF16, PREC_FP32, check, dll, ptr, stream_ptr


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


def positional_encoding(x, freqs, include_input=True):
    """model.py:20-26 materialised: x [..., d] -> [..., d*(include_input + 2L)]."""
    x = _lib.f32c(x)
    _lib.require_cuda(x, freqs)
    d = x.shape[-1]
    n = x.numel() // max(d, 1)
    L = freqs.numel()
    out = _empty(tuple(x.shape[:-1]) + (d * (int(include_input) + 2 * L),), x)
    check(dll().nerf_positional_encoding(ptr(x), n, d, ptr(freqs), L, int(bool(include_input)), ptr(out),
                                         stream_ptr(x.device)), "nerf_positional_encoding")
    return out


# ---------------------------------------------------------------------------------------------
# sampling
# ---------------------------------------------------------------------------------------------
def stratified_z(t_vals, t_rand, R, near, far):
    """renderer.py:52-61.  t_vals [S] (torch.linspace on device), t_rand [R,S] or None."""
    _lib.require_cuda(t_vals, t_rand)
    S = t_vals.numel()
    z = _empty((R, S), t_vals)
    check(dll().nerf_stratified_z(ptr(t_vals), ptr(t_rand), R, S, float(near), float(far), ptr(z),
                                  stream_ptr(z.device)), "nerf_stratified_z")
    return z


def sample_pdf(bins, weights, u, check_mode=False):
    """renderer.py:165-199.  bins [R,NB], weights [R,NB-1]; u [N_imp] (shared) or [R,N_imp]."""
    _lib.require_cuda(u)
    for t in (bins, weights):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise RuntimeError("sample_pdf: bins/weights must be 2-D float32 CUDA tensors with unit inner stride")
    R, NB = bins.shape
    if weights.shape != (R, NB - 1):
        raise RuntimeError(f"sample_pdf: weights must be [R, NB-1]; got {tuple(weights.shape)} for bins {tuple(bins.shape)}")
    n_imp = u.shape[-1]
    shared = 1 if u.dim() == 1 else 0
    if not shared and u.shape != (R, n_imp):
        raise RuntimeError("sample_pdf: u must be [N] or [R,N]")
    samples = _empty((R, n_imp), bins)
    inds = _empty((R, n_imp), bins, torch.int64) if check_mode else None
    cdf = _empty((R, NB), bins) if check_mode else None
    check(dll().nerf_sample_pdf(ptr(bins), bins.stride(0), ptr(weights), weights.stride(0), ptr(u), shared,
                                R, NB, n_imp, ptr(samples), ptr(inds), ptr(cdf), stream_ptr(bins.device)),
          "nerf_sample_pdf")
    return (samples, inds, cdf) if check_mode else samples


def resample_merge(z_coarse, weights, u, check_mode=False):
    """renderer.py:86-90 fused: z_mid, _sample_pdf(weights[:,1:-1]), sort(cat[z, z_samples])."""
    _lib.require_cuda(z_coarse, weights, u)
    R, S = z_coarse.shape
    n_imp = u.shape[-1]
    shared = 1 if u.dim() == 1 else 0
    z_fine = _empty((R, S + n_imp), z_coarse)
    zs = _empty((R, n_imp), z_coarse) if check_mode else None
    inds = _empty((R, n_imp), z_coarse, torch.int64) if check_mode else None
    cdf = _empty((R, S - 1), z_coarse) if check_mode else None
    check(dll().nerf_resample_merge(ptr(z_coarse), ptr(weights), ptr(u), shared, R, S, n_imp, ptr(z_fine),
                                    ptr(zs), ptr(inds), ptr(cdf), stream_ptr(z_coarse.device)),
          "nerf_resample_merge")
    return (z_fine, zs, inds, cdf) if check_mode else z_fine


# ---------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------
def composite_fwd(raw, z_vals, rays_d, noise, white_bkgd, need_weights=True):
    _lib.require_cuda(raw, z_vals, rays_d, noise)
    R, S = z_vals.shape
    rgb, depth, acc = _empty((R, 3), raw), _empty((R,), raw), _empty((R,), raw)
    w = _empty((R, S), raw) if need_weights else None
    check(dll().nerf_composite_fwd(ptr(raw), ptr(z_vals), ptr(rays_d), ptr(noise), R, S, int(bool(white_bkgd)),
                                   ptr(rgb), ptr(depth), ptr(acc), ptr(w), stream_ptr(raw.device)),
          "nerf_composite_fwd")
    return rgb, depth, acc, w


def composite_bwd(raw, z_vals, rays_d, noise, white_bkgd, d_rgb, d_depth=None, d_acc=None, d_weights=None):
    _lib.require_cuda(raw, z_vals, rays_d, noise, d_rgb, d_depth, d_acc, d_weights)
    R, S = z_vals.shape
    d_raw = _empty((R, S, 4), raw)
    check(dll().nerf_composite_bwd(ptr(raw), ptr(z_vals), ptr(rays_d), ptr(noise), R, S, int(bool(white_bkgd)),
                                   ptr(d_rgb), ptr(d_depth), ptr(d_acc), ptr(d_weights), ptr(d_raw),
                                   stream_ptr(raw.device)), "nerf_composite_bwd")
    return d_raw


def _cg(t):
    """contiguous float32 gradient or None."""
    return None if t is None else t.contiguous().float()


class CompositeFn(torch.autograd.Function):
    """_raw2outputs with its analytic backward w.r.t. raw (used when the renderer's
    _raw2outputs is called on a tensor that requires grad)."""

    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise, white_bkgd):
        rgb, depth, acc, w = composite_fwd(raw, z_vals, rays_d, noise, white_bkgd, True)
        ctx.save_for_backward(raw, z_vals, rays_d, noise if noise is not None else raw.new_empty(0))
        ctx.white = white_bkgd
        ctx.has_noise = noise is not None
        return rgb, depth, acc, w

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_acc, d_w):
        raw, z, d, noise = ctx.saved_tensors
        d_raw = composite_bwd(raw, z, d, noise if ctx.has_noise else None, ctx.white,
                              _cg(d_rgb) if d_rgb is not None else torch.zeros_like(d), _cg(d_depth), _cg(d_acc), _cg(d_w))
        return d_raw, None, None, None, None


# ---------------------------------------------------------------------------------------------
# MLP
# ---------------------------------------------------------------------------------------------
def mlp_workspace(M, precision, save, device):
    n = dll().nerf_mlp_workspace_bytes(int(M), int(precision), int(bool(save)))
    return torch.empty(max(int(n), 16), device=device, dtype=torch.uint8)


def mlp_fwd_rays(model, rays_o, rays_d, z_vals, coord_scale, precision, save, density_only=False):
    """density_only (inference only): raw[..., 3] alone is computed (include/nerf_b200.h NERF_FWD_DENSITY_ONLY)."""
    _lib.require_cuda(rays_o, rays_d, z_vals)
    if density_only and save:
        raise RuntimeError("mlp_fwd_rays: density_only is an inference mode")
    R, S = z_vals.shape
    raw = _empty((R, S, 4), z_vals)
    mode = _lib.FWD_DENSITY_ONLY if density_only else int(bool(save))
    ws = mlp_workspace(R * S, precision, mode, z_vals.device)
    packed = model.packed_weights() if precision == PREC_BF16 else None
    check(dll().nerf_mlp_fwd_rays(ptr(rays_o), ptr(rays_d), ptr(z_vals), R, S, float(coord_scale),
                                  ptr(model.flat_params), ptr(packed), ptr(raw), ptr(ws), ws.numel(),
                                  int(precision), mode, stream_ptr(z_vals.device)), "nerf_mlp_fwd_rays")
    return raw, (ws if save else None)


def mlp_fwd_encoded(model, x_enc, d_enc, precision, save):
    _lib.require_cuda(x_enc, d_enc)
    M = x_enc.shape[0]
    out = _empty((M, 4), x_enc)
    ws = mlp_workspace(M, precision, save, x_enc.device)
    packed = model.packed_weights() if precision == PREC_BF16 else None
    check(dll().nerf_mlp_fwd_encoded(ptr(x_enc), ptr(d_enc), M, ptr(model.flat_params), ptr(packed), ptr(out),
                                     ptr(ws), ws.numel(), int(precision), int(bool(save)),
                                     stream_ptr(x_enc.device)), "nerf_mlp_fwd_encoded")
    return out, (ws if save else None)


def mlp_bwd(model, d_raw, ws, precision, flat_grads, rows_per_dir, stage=_lib.BWD_ALL):
    """stage: BWD_ALL (default) | BWD_DGRAD | BWD_WGRAD -- see include/nerf_b200.h nerf_mlp_bwd_stage."""
    _lib.require_cuda(d_raw, flat_grads)
    M = d_raw.numel() // 4
    packed = model.packed_weights() if precision == PREC_BF16 else None
    check(dll().nerf_mlp_bwd_stage(ptr(d_raw), M, int(rows_per_dir), ptr(model.flat_params), ptr(packed), ptr(flat_grads),
                                   ptr(ws), ws.numel(), int(precision), int(stage), stream_ptr(d_raw.device)),
          "nerf_mlp_bwd")


def _untile(img, rows, feats):
    """Decode a shared-memory tile image (csrc/tc_common.cuh) into a row-major [rows, feats] tensor:
    per 128-row tile and 64-feature block a [128][8 chunks][8] block whose 16-byte chunks are
    XOR-swizzled by (row & 7)."""
    nt, nfb = rows // 128, feats // 64
    t = img.view(nt, nfb, 128, 8, 8)
    r = torch.arange(128, device=img.device).view(1, 1, 128, 1, 1)
    c = torch.arange(8, device=img.device).view(1, 1, 1, 8, 1)
    idx = (c ^ (r & 7)).expand(nt, nfb, 128, 8, 8)
    return torch.gather(t, 3, idx).permute(0, 2, 1, 3, 4).reshape(rows, feats)


def bf16_workspace_views(ws, M):
    """Decoded views of the bf16-mode MLP workspace written by a forward with save=1 and by the
    backward (layout: csrc/tc_common.cuh ws_layout).  For tests and debugging."""
    def al(n):
        return (n + 1023) & ~1023
    Mp = (M + 511) // 512 * 512
    out, off = {}, 0
    spec = (("vb", M * 128 * 4, torch.float32, (M, 128), None),
            ("de", M * 32 * 4, torch.float32, (M, 32), None),
            ("act", 9 * Mp * 256 * 2, torch.bfloat16, None, (9, 256)),
            ("hv", Mp * 128 * 2, torch.bfloat16, None, (1, 128)),
            ("xenc", Mp * 64 * 2, torch.bfloat16, None, (1, 64)),
            ("de16", Mp * 64 * 2, torch.bfloat16, None, (1, 64)),
            ("mask", 8 * M * 8 * 4, torch.int32, (8, M, 8), None),
            ("hvmask", M * 4 * 4, torch.int32, (M, 4), None),
            ("dpre", 9 * Mp * 256 * 2, torch.bfloat16, None, (9, 256)),
            ("dhv", Mp * 128 * 2, torch.bfloat16, None, (1, 128)),
            ("flags", (10 * (Mp // 128) + 160) * 4, torch.int32, (10 * (Mp // 128) + 160,), None))
    for name, nbytes, dt, shape, img in spec:
        flat = ws[off:off + nbytes].view(dt)
        if img is None:
            out[name] = flat.view(shape)
        else:
            n, feats = img
            dec = torch.stack([_untile(flat.view(n, -1)[i], Mp, feats)[:M] for i in range(n)])
            out[name] = dec if n > 1 else dec[0]
        off += al(nbytes)
    if off != ws.numel():
        raise RuntimeError(f"workspace size {ws.numel()} does not match the save layout for M={M} ({off})")
    return out


def _param_grads(model, run_bwd):
    """Route parameter gradients.  Fast path: accumulate straight into the model's flat gradient
    buffer (the 24 ``p.grad`` are views of it) and return None to autograd -- one kernel-side
    accumulation instead of 24 AccumulateGrad nodes.  If the user has bound foreign ``.grad``
    tensors, fall back to handing autograd 24 views of a fresh flat buffer."""
    if model._grads_bound_or_bindable():
        flat = model._bind_flat_grads()
        run_bwd(flat)
        return [None] * len(model._param_list)
    flat = torch.zeros_like(model.flat_params)
    run_bwd(flat)
    return list(model._views_of(flat))


class RenderPassFn(torch.autograd.Function):
    """One pass of _render_rays: points + encoding + MLP + compositing, fused on the device
    (reference renderer.py:63-80 for coarse, :91-107 for fine).  Differentiable w.r.t. the model
    parameters through rgb_map / depth_map / acc_map; `weights` is non-differentiable, exactly as
    in the reference where it only feeds the detached _sample_pdf (renderer.py:87-88)."""

    @staticmethod
    def forward(ctx, model, rays_o, rays_d, z_vals, noise, white_bkgd, coord_scale, precision, save, *params):
        # `save` is decided by the caller: grad mode is always off inside Function.forward
        # save == "density": weights-only coarse pass (rgb/depth/acc maps of this pass are not meaningful)
        # save == "lazy":    differentiable WITHOUT paying for it up front: the forward runs the inference kernel (nothing
        #                    saved); if a gradient ever reaches this pass, backward() re-runs the forward in save mode on
        #                    the same samples / noise first (activation checkpointing of the whole pass).  Used for the
        #                    coarse pass, whose maps the reference returns differentiable but never puts in its loss.
        density = isinstance(save, str) and save == "density"
        lazy = isinstance(save, str) and save == "lazy"
        raw, ws = mlp_fwd_rays(model, rays_o, rays_d, z_vals, coord_scale, precision, False if (density or lazy) else save,
                               density_only=density)
        rgb, depth, acc, w = composite_fwd(raw, z_vals, rays_d, noise, white_bkgd, True)
        ctx.model, ctx.white, ctx.precision = model, white_bkgd, precision
        ctx.ws, ctx.noise = ws, noise
        ctx.lazy = (rays_o, float(coord_scale)) if lazy else None
        ctx.save_for_backward(raw, z_vals, rays_d)
        ctx.mark_non_differentiable(w)
        return rgb, depth, acc, w

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_acc, _d_w):
        raw, z, d = ctx.saved_tensors
        if ctx.ws is None and ctx.lazy is not None:
            rays_o, coord_scale = ctx.lazy                    # recompute the pass in save mode (same z, same noise)
            raw, ctx.ws = mlp_fwd_rays(ctx.model, rays_o, d, z, coord_scale, ctx.precision, True)
        if ctx.ws is None:
            raise RuntimeError("RenderPassFn.backward: forward ran without saving activations")
        d_rgb = _cg(d_rgb) if d_rgb is not None else torch.zeros_like(d)
        d_raw = composite_bwd(raw, z, d, ctx.noise, ctx.white, d_rgb, _cg(d_depth), _cg(d_acc), None)
        model, ws, prec = ctx.model, ctx.ws, ctx.precision
        S = z.shape[1]
        grads = _param_grads(model, lambda flat: mlp_bwd(model, d_raw, ws, prec, flat, S))
        ctx.ws = None
        return (None,) * 9 + tuple(grads)


class MLPEncodedFn(torch.autograd.Function):
    """Drop-in NeRFMLP.forward on pre-encoded inputs (reference model.py:57-81)."""

    @staticmethod
    def forward(ctx, model, x_enc, d_enc, precision, save, *params):
        out, ws = mlp_fwd_encoded(model, x_enc, d_enc, precision, save)
        ctx.model, ctx.ws, ctx.precision = model, ws, precision
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.ws is None:
            raise RuntimeError("MLPEncodedFn.backward: forward ran without saving activations")
        model, ws, prec = ctx.model, ctx.ws, ctx.precision
        d_out = _cg(d_out)
        grads = _param_grads(model, lambda flat: mlp_bwd(model, d_out, ws, prec, flat, 1))
        ctx.ws = None
        return (None,) * 5 + tuple(grads)


# ---------------------------------------------------------------------------------------------
# optimiser / loss
# ---------------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    _lib.require_cuda(params, grads, exp_avg, exp_avg_sq)
    check(dll().nerf_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(),
                               float(lr), float(betas[0]), float(betas[1]), float(eps), int(step),
                               float(grad_scale), stream_ptr(params.device)), "nerf_adam_step")


class MSELossFn(torch.autograd.Function):
    """mean((pred-target)**2) (scripts/train.py:376) in one launch, gradient produced in the same pass."""

    @staticmethod
    def forward(ctx, pred, target):
        pred, target = _lib.f32c(pred), _lib.f32c(target)
        loss = _empty((), pred)
        d_pred = torch.empty_like(pred)
        check(dll().nerf_mse_loss(ptr(pred), ptr(target), pred.numel(), ptr(loss), ptr(d_pred),
                                  stream_ptr(pred.device)), "nerf_mse_loss")
        ctx.save_for_backward(d_pred)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d_pred,) = ctx.saved_tensors
        return d_pred * g, None


def mse_loss(pred, target):
    return MSELossFn.apply(pred, target)

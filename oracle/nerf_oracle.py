"""CPU oracle for the NeRF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of the reference algorithm (dgsmith7/nerf-mlp,
``nerfmlp/model.py`` + ``nerfmlp/renderer.py``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the timed CPU
baseline.  Nothing under ``nerf_mlp_b200/`` imports it.

Parity pin: the reference ships no tests, golden vectors or fixtures (SURVEY.md
section 4), so this oracle is pinned against outputs of the *reference itself*,
imported and run in the build container by ``tests/golden/make_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.

All arithmetic is float32 unless noted; the operation order follows the
reference line by line (citations are ``file:line`` into ``/root/reference``).
Where torch's CPU kernels accumulate float32 scans in double (``cumsum``,
``cumprod``: at::acc_type<float, /*cuda=*/false> is double) the oracle does the
same so that it tracks the reference's CPU path to the last ulp where possible.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------------------------
# Parameter container: same 24 tensors / names / shapes as NeRFMLP.state_dict()  (model.py:39-53)
# ----------------------------------------------------------------------------------------------
LAYER_SHAPES = (
    # name, out, in
    ("pts_linears.0", 256, 63),
    ("pts_linears.1", 256, 256),
    ("pts_linears.2", 256, 256),
    ("pts_linears.3", 256, 256),
    ("pts_linears.4", 256, 256),
    ("pts_linears.5", 256, 319),
    ("pts_linears.6", 256, 256),
    ("pts_linears.7", 256, 256),
    ("sigma_linear", 1, 256),
    ("bottleneck_linear", 256, 256),
    ("view_linear", 128, 283),
    ("rgb_linear", 3, 128),
)
PARAM_NAMES = tuple(f"{n}.{k}" for n, _, _ in LAYER_SHAPES for k in ("weight", "bias"))
N_PARAMS = sum(o * i + o for _, o, i in LAYER_SHAPES)  # 595 844
assert N_PARAMS == 595844


def init_params(seed: int = 0) -> dict:
    """Random init with the distribution of torch's nn.Linear default (model.py:39-53 builds
    nn.Linear layers; their reset_parameters() draws weight and bias from
    U(-1/sqrt(fan_in), 1/sqrt(fan_in))).  Uses a numpy Generator so the *same bits* are
    reproducible on any box without torch's RNG."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, out_f, in_f in LAYER_SHAPES:
        bound = 1.0 / math.sqrt(in_f)
        p[f"{name}.weight"] = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(F32)
        p[f"{name}.bias"] = rng.uniform(-bound, bound, size=(out_f,)).astype(F32)
    return p


def flatten_params(p: dict) -> np.ndarray:
    return np.concatenate([p[n].reshape(-1) for n in PARAM_NAMES]).astype(F32)


def unflatten_params(flat: np.ndarray) -> dict:
    p, off = {}, 0
    for name, out_f, in_f in LAYER_SHAPES:
        p[f"{name}.weight"] = flat[off:off + out_f * in_f].reshape(out_f, in_f)
        off += out_f * in_f
        p[f"{name}.bias"] = flat[off:off + out_f]
        off += out_f
    return p


# ----------------------------------------------------------------------------------------------
# Positional encoding                                                        (model.py:5-26)
# ----------------------------------------------------------------------------------------------
def positional_encoding(x: np.ndarray, num_freqs: int) -> np.ndarray:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^{L-1} x), cos(2^{L-1} x)]; no pi (model.py:22-26).
    freq_bands = 2**linspace(0, L-1, L) are exact powers of two in fp32 (model.py:15)."""
    x = np.asarray(x, dtype=F32)
    out = [x]
    for k in range(num_freqs):
        f = F32(2.0 ** k)
        fx = (f * x).astype(F32)
        out.append(np.sin(fx, dtype=F32))
        out.append(np.cos(fx, dtype=F32))
    return np.concatenate(out, axis=-1)


# ----------------------------------------------------------------------------------------------
# MLP forward / backward                                                      (model.py:57-81)
# ----------------------------------------------------------------------------------------------
def _linear(h, w, b, relu=False):
    y = h @ w.T                      # float32 x float32 -> float32 (BLAS sgemm)
    y += b
    if relu:
        np.maximum(y, F32(0), out=y)
    return y


def mlp_forward(p: dict, x: np.ndarray, viewdirs: np.ndarray, save: bool = False):
    """NeRFMLP.forward (model.py:57-81): 8x256 ReLU trunk, [x,h] concat feeding layer 5
    (model.py:62-63), sigma head (:69), bottleneck (:70), [bottleneck,viewdirs] (:72),
    view layer + ReLU (:73-74), rgb (:75); output [rgb, sigma] (:77)."""
    x = np.asarray(x, F32)
    viewdirs = np.asarray(viewdirs, F32)
    saved = {"x": x, "d": viewdirs, "in": [], "out": []}
    h = x
    for i in range(8):
        if i == 5:
            h = np.concatenate([x, h], -1)
        if save:
            saved["in"].append(h)
        h = _linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"], relu=True)
        if save:
            saved["out"].append(h)
    sigma = _linear(h, p["sigma_linear.weight"], p["sigma_linear.bias"])
    bott = _linear(h, p["bottleneck_linear.weight"], p["bottleneck_linear.bias"])
    hv_in = np.concatenate([bott, viewdirs], -1)
    hv = _linear(hv_in, p["view_linear.weight"], p["view_linear.bias"], relu=True)
    rgb = _linear(hv, p["rgb_linear.weight"], p["rgb_linear.bias"])
    out = np.concatenate([rgb, sigma], -1)
    if save:
        saved.update(h7=h, hv_in=hv_in, hv=hv)
        return out, saved
    return out


def mlp_backward(p: dict, saved: dict, d_out: np.ndarray) -> dict:
    """Reverse-mode gradient of mlp_forward w.r.t. the 24 parameters (what autograd computes at
    scripts/train.py:382).  Inputs x / viewdirs receive no gradient (they do not depend on
    parameters)."""
    g = {}
    d_out = np.asarray(d_out, F32)
    d_rgb, d_sigma = d_out[:, :3], d_out[:, 3:4]
    # rgb_linear
    g["rgb_linear.weight"] = d_rgb.T @ saved["hv"]
    g["rgb_linear.bias"] = d_rgb.sum(0)
    d_hv = d_rgb @ p["rgb_linear.weight"]
    d_hv *= saved["hv"] > 0
    # view_linear
    g["view_linear.weight"] = d_hv.T @ saved["hv_in"]
    g["view_linear.bias"] = d_hv.sum(0)
    d_bott = (d_hv @ p["view_linear.weight"])[:, :256]
    # bottleneck + sigma
    g["bottleneck_linear.weight"] = d_bott.T @ saved["h7"]
    g["bottleneck_linear.bias"] = d_bott.sum(0)
    g["sigma_linear.weight"] = d_sigma.T @ saved["h7"]
    g["sigma_linear.bias"] = d_sigma.sum(0)
    d_h = d_bott @ p["bottleneck_linear.weight"] + d_sigma @ p["sigma_linear.weight"]
    for i in range(7, -1, -1):
        d_pre = d_h * (saved["out"][i] > 0)
        g[f"pts_linears.{i}.weight"] = d_pre.T @ saved["in"][i]
        g[f"pts_linears.{i}.bias"] = d_pre.sum(0)
        if i == 0:
            break
        d_in = d_pre @ p[f"pts_linears.{i}.weight"]
        d_h = d_in[:, 63:] if i == 5 else d_in
    return {k: np.asarray(v, F32) for k, v in g.items()}


# ----------------------------------------------------------------------------------------------
# Volume rendering integral                                               (renderer.py:114-163)
# ----------------------------------------------------------------------------------------------
def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x, dtype=F32))).astype(F32)


def raw2outputs(raw, z_vals, rays_d, white_bkgd=True, noise=None):
    raw = np.asarray(raw, F32)
    z_vals = np.asarray(z_vals, F32)
    rays_d = np.asarray(rays_d, F32)
    dists = z_vals[..., 1:] - z_vals[..., :-1]                                   # :120
    dists = np.concatenate([dists, np.full_like(dists[..., :1], 1e10)], -1)      # :123
    dists = (dists * np.linalg.norm(rays_d[..., None, :], axis=-1)).astype(F32)  # :127
    rgb = _sigmoid(raw[..., :3])                                                 # :130
    sig = raw[..., 3] if noise is None else (raw[..., 3] + np.asarray(noise, F32))  # :134-136
    with np.errstate(over="ignore"):
        alpha = (F32(1.0) - np.exp(-np.maximum(sig, F32(0)) * dists, dtype=F32)).astype(F32)  # :140
    one_m = ((F32(1.0) - alpha) + F32(1e-10)).astype(F32)                        # :147 association
    # exclusive cumprod; torch CPU accumulates float cumprod in double                    # :147
    t_incl = np.cumprod(one_m.astype(np.float64), axis=-1).astype(F32)
    trans = np.concatenate([np.ones_like(alpha[..., :1]), t_incl[..., :-1]], -1)
    weights = (alpha * trans).astype(F32)                                        # :148
    rgb_map = np.sum(weights[..., None] * rgb, axis=-2, dtype=F32)               # :151
    depth_map = np.sum(weights * z_vals, axis=-1, dtype=F32)                     # :154
    acc_map = np.sum(weights, axis=-1, dtype=F32)                                # :157
    if white_bkgd:
        rgb_map = rgb_map + (F32(1.0) - acc_map[..., None])                      # :160-161
    return rgb_map.astype(F32), depth_map.astype(F32), acc_map.astype(F32), weights


def raw2outputs_backward(raw, z_vals, rays_d, white_bkgd, d_rgb_map, d_depth=None, d_acc=None,
                         d_weights=None, noise=None):
    """Analytic reverse-mode gradient of raw2outputs w.r.t. raw:
    g_i = dC.rgb_i + dD z_i + dA - [white] sum_c dC_c (+ dW_i);
    dalpha_i = T_i g_i - (sum_{k>i} w_k g_k) / ((1-alpha_i)+1e-10);
    dsigma_i = dalpha_i dists_i exp(-sigma'_i dists_i) [raw_sigma+noise > 0];
    drgb_raw = w_i dC rgb (1-rgb).   Computed in float64 and rounded (it is a checker)."""
    raw = np.asarray(raw, np.float64)
    z = np.asarray(z_vals, np.float64)
    R, S = z.shape
    dn = np.linalg.norm(np.asarray(rays_d, F32), axis=-1).astype(np.float64)
    dists = np.concatenate([(np.asarray(z_vals, F32)[:, 1:] - np.asarray(z_vals, F32)[:, :-1]),
                            np.full((R, 1), 1e10, F32)], -1)
    dists = (dists * dn[:, None].astype(F32)).astype(np.float64)
    rgb = 1.0 / (1.0 + np.exp(-raw[..., :3]))
    sig = raw[..., 3] + (0.0 if noise is None else np.asarray(noise, np.float64))
    sp = np.maximum(sig, 0.0)
    e = np.exp(-sp * dists)
    alpha = 1.0 - e
    one_m = (1.0 - alpha) + 1e-10
    T = np.concatenate([np.ones((R, 1)), np.cumprod(one_m, -1)[:, :-1]], -1)
    w = alpha * T
    dC = np.asarray(d_rgb_map, np.float64)
    g = (dC[:, None, :] * rgb).sum(-1)
    if white_bkgd:
        g = g - dC.sum(-1, keepdims=True)
    if d_depth is not None:
        g = g + np.asarray(d_depth, np.float64)[:, None] * z
    if d_acc is not None:
        g = g + np.asarray(d_acc, np.float64)[:, None]
    if d_weights is not None:
        g = g + np.asarray(d_weights, np.float64)
    wg = w * g
    suffix = np.concatenate([np.cumsum(wg[:, ::-1], -1)[:, ::-1][:, 1:], np.zeros((R, 1))], -1)
    d_alpha = T * g - suffix / one_m
    d_sigma = d_alpha * dists * e * (sig > 0)
    d_raw = np.empty((R, S, 4), np.float64)
    d_raw[..., :3] = w[..., None] * dC[:, None, :] * rgb * (1.0 - rgb)
    d_raw[..., 3] = d_sigma
    return d_raw.astype(F32)


# ----------------------------------------------------------------------------------------------
# Hierarchical sampling                                                   (renderer.py:165-199)
# ----------------------------------------------------------------------------------------------
def searchsorted_right(cdf: np.ndarray, u: np.ndarray) -> np.ndarray:
    """torch.searchsorted(cdf, u, right=True) (renderer.py:185): per row, the number of cdf
    entries <= u.  Pure compare/count semantics, int64, bit-exact by construction."""
    cdf = np.asarray(cdf, F32)
    u = np.asarray(u, F32)
    return (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)


def pdf_to_cdf(weights: np.ndarray) -> np.ndarray:
    """weights -> cdf (renderer.py:172-175).  Summation order is pinned so that the CUDA kernel
    can reproduce the bits: the normaliser and every prefix are accumulated in float64 and
    rounded to float32 once (torch's CPU cumsum does exactly this for the prefixes; its `sum`
    uses a vectorised fp32 cascade whose last ulp is implementation-defined)."""
    w = (np.asarray(weights, F32) + F32(1e-5)).astype(F32)                       # :172
    total = np.sum(w.astype(np.float64), -1, keepdims=True).astype(F32)
    pdf = (w / total).astype(F32)                                                # :173
    cdf = np.cumsum(pdf.astype(np.float64), -1).astype(F32)                      # :174
    return np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)                # :175


def sample_pdf(bins, weights, u, return_aux=False, cdf=None):
    """renderer.py:165-199 with the uniform samples `u` passed in ([N_imp] shared when det,
    or [R, N_imp]); see SURVEY.md H4 for why linspace/rand are inputs.  `cdf` overrides the
    internally built cdf (used to check the index/lerp stage against another implementation's
    cdf bits: where the pdf is ~0 the inverse cdf is discontinuous, so a 1-ulp cdf difference can
    move a sample by a whole bin)."""
    bins = np.asarray(bins, F32)
    cdf = pdf_to_cdf(weights) if cdf is None else np.asarray(cdf, F32)
    u = np.asarray(u, F32)
    if u.ndim == 1:
        u = np.broadcast_to(u, (cdf.shape[0], u.shape[0]))                       # :180
    inds = searchsorted_right(cdf, u)                                            # :185
    below = np.maximum(inds - 1, 0)                                              # :186
    above = np.minimum(inds, cdf.shape[-1] - 1)                                  # :187
    cdf_b = np.take_along_axis(cdf, below, -1)
    cdf_a = np.take_along_axis(cdf, above, -1)
    bins_b = np.take_along_axis(bins, below, -1)
    bins_a = np.take_along_axis(bins, above, -1)
    denom = (cdf_a - cdf_b).astype(F32)                                          # :194
    denom = np.where(denom < F32(1e-5), F32(1.0), denom)                         # :195
    t = ((u - cdf_b) / denom).astype(F32)                                        # :196
    samples = (bins_b + t * (bins_a - bins_b)).astype(F32)                       # :197
    if return_aux:
        return samples, cdf, inds
    return samples


# ----------------------------------------------------------------------------------------------
# _render_rays                                                             (renderer.py:47-112)
# ----------------------------------------------------------------------------------------------
@dataclass
class RenderConfig:
    N_samples: int = 64
    N_importance: int = 128
    near: float = 2.0
    far: float = 6.0
    white_bkgd: bool = True
    coord_scale: float = 1.0
    pos_enc_L: int = 10
    dir_enc_L: int = 4


def stratified_z(t_vals, near, far, R, t_rand=None):
    t_vals = np.asarray(t_vals, F32)
    z = (F32(near) * (F32(1.0) - t_vals) + F32(far) * t_vals).astype(F32)       # :53
    z = np.broadcast_to(z, (R, z.shape[0]))                                      # :54
    if t_rand is not None:                                                       # :56-61
        mids = (F32(0.5) * (z[..., 1:] + z[..., :-1])).astype(F32)
        upper = np.concatenate([mids, z[..., -1:]], -1)
        lower = np.concatenate([z[..., :1], mids], -1)
        z = (lower + (upper - lower) * np.asarray(t_rand, F32)).astype(F32)
    return np.ascontiguousarray(z)


def encode_samples(rays_o, rays_d, z_vals, cfg: RenderConfig):
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]           # :63
    pts = pts.reshape(-1, 3).astype(F32)
    if cfg.coord_scale != 1.0:
        pts = (pts * F32(cfg.coord_scale)).astype(F32)                           # :67-68
    x_enc = positional_encoding(pts, cfg.pos_enc_L)                              # :70
    viewdirs = rays_d / (np.linalg.norm(rays_d, axis=-1, keepdims=True).astype(F32) + F32(1e-8))  # :72
    d_enc = positional_encoding(viewdirs.astype(F32), cfg.dir_enc_L)             # :73
    S = z_vals.shape[1]
    d_enc = np.repeat(d_enc[:, None, :], S, axis=1).reshape(-1, d_enc.shape[-1])  # :74
    return x_enc, d_enc


def render_rays(p, rays_o, rays_d, cfg: RenderConfig, t_vals, u, t_rand=None,
                noise_coarse=None, noise_fine=None, save=False, z_fine_override=None):
    """NeRFRenderer._render_rays (renderer.py:47-112).  `t_vals` = linspace(0,1,N_samples),
    `u` = linspace(0,1,N_importance) when perturb==0 else rand [R,N_imp]; `t_rand` given iff
    perturb>0.  Returns the six maps plus every intermediate (for stage-level parity)."""
    rays_o = np.asarray(rays_o, F32)
    rays_d = np.asarray(rays_d, F32)
    R = rays_o.shape[0]
    z = stratified_z(t_vals, cfg.near, cfg.far, R, t_rand)
    x_enc, d_enc = encode_samples(rays_o, rays_d, z, cfg)
    raw = mlp_forward(p, x_enc, d_enc).reshape(R, cfg.N_samples, 4)              # :76-77
    rgb0, depth0, acc0, w0 = raw2outputs(raw, z, rays_d, cfg.white_bkgd, noise_coarse)  # :79-80
    out = {"z_vals": z, "raw_coarse": raw, "weights_coarse": w0}
    if cfg.N_importance <= 0:                                                    # :112
        out.update(rgb_map=rgb0, depth_map=depth0, acc_map=acc0)
        return out
    z_mid = (F32(0.5) * (z[..., 1:] + z[..., :-1])).astype(F32)                  # :86
    z_samples, cdf, inds = sample_pdf(z_mid, w0[..., 1:-1], u, return_aux=True)  # :87
    z_fine = np.sort(np.concatenate([z, z_samples], -1), -1)                     # :90
    if z_fine_override is not None:
        # stage-isolation hook for tests: continue from another implementation's fine depths (the
        # inverse cdf is ill-conditioned at random init, see tests/test_oracle_golden.py)
        z_fine = np.asarray(z_fine_override, F32)
    xf, df = encode_samples(rays_o, rays_d, z_fine, cfg)                         # :91-101
    S_f = z_fine.shape[1]
    if save:
        raw_f, saved = mlp_forward(p, xf, df, save=True)
        out["saved"] = saved
    else:
        raw_f = mlp_forward(p, xf, df)
    raw_f = raw_f.reshape(R, S_f, 4)                                             # :103-104
    rgb, depth, acc, w = raw2outputs(raw_f, z_fine, rays_d, cfg.white_bkgd, noise_fine)  # :106-107
    out.update(rgb_map=rgb, depth_map=depth, acc_map=acc,
               rgb_map_coarse=rgb0, depth_map_coarse=depth0, acc_map_coarse=acc0,
               cdf=cdf, inds=inds, z_samples=z_samples, z_fine=z_fine, raw_fine=raw_f,
               weights_fine=w)
    return out


def train_grads(p, rays_o, rays_d, target, cfg: RenderConfig, t_vals, u, t_rand=None,
                z_fine_override=None):
    """loss = mean((rgb_map_fine - target)^2) (scripts/train.py:374-376) and its parameter
    gradients.  The coarse pass receives no gradient: the loss reads only the fine rgb_map and
    z_samples is detached (renderer.py:88)."""
    out = render_rays(p, rays_o, rays_d, cfg, t_vals, u, t_rand, save=True,
                      z_fine_override=z_fine_override)
    diff = (out["rgb_map"] - np.asarray(target, F32)).astype(F32)
    loss = F32(np.mean(diff.astype(np.float64) ** 2))
    d_rgb_map = (F32(2.0) * diff / F32(diff.size)).astype(F32)
    d_raw = raw2outputs_backward(out["raw_fine"], out["z_fine"], rays_d, cfg.white_bkgd, d_rgb_map)
    grads = mlp_backward(p, out["saved"], d_raw.reshape(-1, 4))
    del out["saved"]
    return loss, grads, out


# ----------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults; scripts/train.py:258)
# ----------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=5e-4, b1=0.9, b2=0.999, eps=1e-8):
    """One torch.optim.Adam step (no amsgrad / weight decay) on flat float32 arrays, in the
    operation order of torch's single-tensor implementation; `step` is 1-based."""
    p, g, m, v = (np.asarray(a, F32) for a in (p, g, m, v))
    m = (m + (g - m) * F32(1 - b1)).astype(F32)                # exp_avg.lerp_(grad, 1-beta1)
    v = (v * F32(b2) + (g * g) * F32(1 - b2)).astype(F32)      # mul_(beta2).addcmul_(g, g, 1-beta2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / F32(math.sqrt(bc2)) + F32(eps)).astype(F32)
    p = (p + (F32(-step_size) * m) / denom).astype(F32)      # addcdiv_: self + value*t1/t2
    return p, m, v


# ----------------------------------------------------------------------------------------------
# Synthetic rays
# ----------------------------------------------------------------------------------------------
def pinhole_rays(H: int, W: int, fov: float = 0.6911):
    """Rays of a pinhole camera at (0,0,4) looking down -z with identity rotation, as built by
    scripts/render_example.py:245-250."""
    focal = 0.5 * W / math.tan(0.5 * fov)
    i, j = np.meshgrid(np.arange(W, dtype=F32), np.arange(H, dtype=F32), indexing="xy")
    dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1)
    rays_d = dirs.reshape(-1, 3).astype(F32)
    rays_o = np.broadcast_to(np.array([0, 0, 4], F32), rays_d.shape).copy()
    return rays_o, rays_d, focal


def random_rays(R: int, seed: int = 0):
    """i.i.d. o ~ N((0,0,4), 0.1^2), non-unit d ~ N(0,I) with d_z <- -|d_z|-1."""
    rng = np.random.default_rng(seed)
    o = (np.array([0, 0, 4], F32) + 0.1 * rng.standard_normal((R, 3))).astype(F32)
    d = rng.standard_normal((R, 3)).astype(F32)
    d[:, 2] = -np.abs(d[:, 2]) - 1.0
    return o, d


# ----------------------------------------------------------------------------------------------
# Callers either side of the path
# ----------------------------------------------------------------------------------------------
def srgb_to_linear(img):
    """nerfmlp/data.py:8-22 (float32 where/power)."""
    img = img.astype(np.float32)
    return np.where(img <= 0.04045, img / 12.92, np.power((img + 0.055) / 1.055, 2.4))


def preprocess_rgba(rgba_u8, white_bkgd=True):
    """nerfmlp/data.py:46-62 without the file I/O: uint8 RGBA [..., 4] -> float32 linear RGB [..., 3]."""
    img = np.asarray(rgba_u8) / 255.0                                       # float64, data.py:47
    rgb, alpha = img[..., :3], img[..., 3:]
    if white_bkgd:
        rgb = rgb * alpha + (1 - alpha)                                     # data.py:55
    return srgb_to_linear(rgb)


def dataset_rays(poses, H, W, focal):
    """nerfmlp/data.py:76-94 followed by __getitem__'s .float() (:99-104): per-image ray tables,
    image-major.  Returns float32 (rays_o [N*H*W,3], rays_d [N*H*W,3])."""
    i, j = np.meshgrid(np.arange(W), np.arange(H), indexing='xy')
    dirs = np.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -np.ones_like(i)], -1)
    ro, rd = [], []
    for pose in poses:
        d = (dirs @ pose[:3, :3].T).reshape(-1, 3)
        ro.append(np.broadcast_to(pose[:3, 3], d.shape))
        rd.append(d)
    return np.concatenate(ro, 0).astype(np.float32), np.concatenate(rd, 0).astype(np.float32)


def linear_to_srgb(img):
    """scripts/render_example.py:12-26 (float32 where/power)."""
    img = img.astype(np.float32)
    with np.errstate(invalid="ignore"):
        return np.where(img <= 0.0031308, img * 12.92, 1.055 * np.power(img, 1 / 2.4) - 0.055)


def to_uint8(rgb, brightness=1.0, gamma_correction=False):
    """scripts/render_example.py:256-271: brightness boost, optional linear->sRGB, clip, 8-bit."""
    rgb = np.asarray(rgb, np.float32)
    if brightness != 1.0:
        rgb = rgb * brightness
    rgb_final = linear_to_srgb(rgb) if gamma_correction else rgb
    return (np.clip(rgb_final, 0, 1) * 255).astype(np.uint8)

"""Drop-in for the reference's nerfmlp/model.py: PositionalEncoding and NeRFMLP.

Same constructor signatures, attributes, state_dict keys/shapes and load_from_numpy semantics as
the reference (model.py:5-127); the arithmetic runs in libnerf_b200's sm_100a kernels.  The 24
fp32 nn.Parameters are views into ONE flat buffer (state_dict order), so that the kernels take a
single pointer, gradients form one contiguous all-reduce message, and Adam is one launch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import N_PARAMS, PREC_BF16, PREC_FP32

_PRECISIONS = {"bf16": PREC_BF16, "fp32": PREC_FP32}


class PositionalEncoding(nn.Module):
    """reference model.py:5-26 (same attributes; freq_bands is a plain attribute, not a buffer)."""

    def __init__(self, num_freqs, include_input=True, log_sampling=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.include_input = include_input
        self.log_sampling = log_sampling
        if log_sampling:
            self.freq_bands = 2.0 ** torch.linspace(0., num_freqs - 1, num_freqs)          # model.py:15
        else:
            self.freq_bands = torch.linspace(2. ** 0, 2. ** (num_freqs - 1), num_freqs)     # model.py:18
        self._dev_freqs = {}

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("nerf_mlp_b200.PositionalEncoding runs on CUDA only (no CPU fallback)")
        f = self._dev_freqs.get(x.device)
        if f is None:
            f = self.freq_bands.to(device=x.device, dtype=torch.float32).contiguous()
            self._dev_freqs[x.device] = f
        return ops.positional_encoding(x, f, self.include_input)


class NeRFMLP(nn.Module):
    """reference model.py:28-127.  Only the default 8x256 / view-dependent configuration (the one
    every caller in the reference constructs) is implemented by the fused kernels; any other
    configuration raises NotImplementedError -- there is no unfused fallback.

    Extra keyword-only argument: precision = 'bf16' (tcgen05 tensor-core kernels, default) or
    'fp32' (CUDA-core check mode, 1e-4 parity gate)."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=27, skips=[5],
                 use_viewdirs=True, output_ch=4, *, precision="bf16"):
        super().__init__()
        if (D, W, input_ch, input_ch_views, bool(use_viewdirs)) != (8, 256, 63, 27, True):
            raise NotImplementedError(
                "nerf_mlp_b200.NeRFMLP implements the reference's default configuration only "
                "(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True)")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.D = D
        self.W = W
        self.input_ch = input_ch
        self.input_ch_views = input_ch_views
        self.skips = skips              # stored but ignored, like the reference (skip hard-wired at i==5)
        self.use_viewdirs = use_viewdirs
        self.precision = precision
        # identical construction order to the reference so that torch.manual_seed(s); NeRFMLP()
        # draws the same initial weights (model.py:39-53)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) for _ in range(1, 5)] +
            [nn.Linear(W + input_ch, W)] +
            [nn.Linear(W, W) for _ in range(6, D)])
        self.sigma_linear = nn.Linear(W, 1)
        self.bottleneck_linear = nn.Linear(W, 256)
        self.view_linear = nn.Linear(256 + input_ch_views, W // 2)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self.flat_params = None
        self._flat_grad = None
        self._packed = None
        self._packed_key = None
        self._dirty = True
        self._flatten()

    # ---- flat storage ------------------------------------------------------------------------
    @property
    def _param_list(self):
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        for l in (self.sigma_linear, self.bottleneck_linear, self.view_linear, self.rgb_linear):
            ps += [l.weight, l.bias]
        return ps

    def _views_of(self, flat):
        off = 0
        for p in self._param_list:
            n = p.numel()
            yield flat[off:off + n].view(p.shape)
            off += n

    def _flatten(self):
        ps = self._param_list
        dev = ps[0].device
        if any(p.dtype != torch.float32 for p in ps):
            raise RuntimeError("NeRFMLP master weights must stay float32 (bf16 copies are derived)")
        flat = torch.empty(N_PARAMS, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, v in zip(ps, self._views_of(flat)):
                v.copy_(p.data)
                p.data = v
        self.flat_params = flat
        self._flat_grad = None
        self._dirty = True

    def _ensure_flat(self):
        ps = self._param_list
        f = self.flat_params
        if (f is None or ps[0].data_ptr() != f.data_ptr()
                or ps[-1].data_ptr() != f.data_ptr() + 4 * (N_PARAMS - ps[-1].numel())):
            self._flatten()

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._flatten()
        return self

    def mark_dirty(self):
        """Call after mutating parameters through a path autograd's version counters do not see
        (``p.data.copy_``, a raw kernel): the packed bf16 image is rebuilt on next use."""
        self._dirty = True

    def packed_weights(self):
        """bf16 packed/swizzled weight image for the tcgen05 kernels (derived, never saved)."""
        self._ensure_flat()
        key = (self.flat_params.data_ptr(), sum(p._version for p in self._param_list))
        if self._packed is None or self._packed.device != self.flat_params.device:
            self._packed = torch.empty(int(_lib.dll().nerf_packed_weight_bytes()), device=self.flat_params.device,
                                       dtype=torch.uint8)
            self._dirty = True
        if self._dirty or key != self._packed_key:
            _lib.check(_lib.dll().nerf_pack_weights(_lib.ptr(self.flat_params), _lib.ptr(self._packed),
                                                    _lib.stream_ptr(self.flat_params.device)), "nerf_pack_weights")
            self._packed_key, self._dirty = key, False
        return self._packed

    # ---- gradient routing (see ops._param_grads) -------------------------------------------------
    def _grads_bound_or_bindable(self):
        ps = self._param_list
        if all(p.grad is None for p in ps):
            return True
        g = self._flat_grad
        if g is None or any(p.grad is None for p in ps):
            return False
        off = 0
        for p in ps:
            if p.grad.data_ptr() != g.data_ptr() + 4 * off or not p.grad.is_contiguous():
                return False
            off += p.numel()
        return True

    def _bind_flat_grads(self):
        ps = self._param_list
        if self._flat_grad is None or self._flat_grad.device != self.flat_params.device:
            self._flat_grad = torch.zeros_like(self.flat_params)
            unbound = True
        else:
            unbound = all(p.grad is None for p in ps)
            if unbound:
                self._flat_grad.zero_()
        if unbound or any(p.grad is None for p in ps):
            for p, v in zip(ps, self._views_of(self._flat_grad)):
                p.grad = v
        return self._flat_grad

    @property
    def flat_grad(self):
        """The flat fp32 gradient buffer (one all-reduce message); None before the first backward."""
        return self._flat_grad if self._grads_bound_or_bindable() else None

    # ---- reference API ---------------------------------------------------------------------------
    def forward(self, x, viewdirs=None):
        """x: (..., 63) encoded points, viewdirs: (..., 27) encoded directions -> (..., 4) =
        [rgb, sigma] raw (reference model.py:57-81)."""
        if viewdirs is None:
            raise NotImplementedError("NeRFMLP.forward without viewdirs is not implemented "
                                      "(the reference has no output_linear in this configuration either)")
        if x.shape[-1] != self.input_ch or viewdirs.shape[-1] != self.input_ch_views:
            raise RuntimeError(f"NeRFMLP.forward: expected (...,{self.input_ch}) and (...,{self.input_ch_views}), "
                               f"got {tuple(x.shape)} and {tuple(viewdirs.shape)}")
        if x.requires_grad or viewdirs.requires_grad:
            raise NotImplementedError("gradients w.r.t. the encoded inputs are not implemented "
                                      "(no caller in the reference needs them)")
        self._ensure_flat()
        lead = x.shape[:-1]
        xe = _lib.f32c(x).reshape(-1, self.input_ch)
        de = _lib.f32c(viewdirs).reshape(-1, self.input_ch_views)
        if xe.shape[0] != de.shape[0]:
            raise RuntimeError("NeRFMLP.forward: x and viewdirs must have the same number of rows")
        save = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        out = ops.MLPEncodedFn.apply(self, xe, de, _PRECISIONS[self.precision], save, *self._param_list)
        return out.view(*lead, 4)

    def load_from_numpy(self, np_arrays):
        """reference model.py:83-127: official TF weight list, arrays are [in,out] (transposed on
        load); order 8 trunk pairs, bottleneck, view, rgb, sigma."""
        idx = 0
        order = list(self.pts_linears) + [self.bottleneck_linear, self.view_linear, self.rgb_linear, self.sigma_linear]
        names = [f"pts_linears[{i}]" for i in range(len(self.pts_linears))] + \
                ["bottleneck_linear", "view_linear", "rgb_linear", "sigma_linear"]
        with torch.no_grad():
            for name, l in zip(names, order):
                w, b = np.asarray(np_arrays[idx]), np.asarray(np_arrays[idx + 1])
                print(f"Loading {name}.weight with shape {tuple(l.weight.shape)} from np_arrays[{idx}].T {w.shape}")
                l.weight.data.copy_(torch.from_numpy(np.ascontiguousarray(w.T)))
                print(f"Loading {name}.bias with shape {tuple(l.bias.shape)} from np_arrays[{idx + 1}].shape {b.shape}")
                l.bias.data.copy_(torch.from_numpy(b))
                idx += 2
        self.mark_dirty()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.mark_dirty()
        return out

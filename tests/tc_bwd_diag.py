#!/usr/bin/env python
"""Diagnosis of the tensor-core backward (dgrad chain + wgrad) against the fp32 check mode and a
plain torch re-derivation on the same inputs (GPU only; a development tool, not a test)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops
from oracle import nerf_oracle as O


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main(R=300, S=64):
    dev = torch.device("cuda")
    p = O.init_params(0)
    o, d = O.random_rays(R, 1)
    to, td = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    z = ops.stratified_z(torch.linspace(0., 1., S, device=dev), None, R, 2.0, 6.0)
    M = R * S
    torch.manual_seed(0)
    d_raw = (torch.randn(M, 4, device=dev) * 1e-3).contiguous()
    models = {}
    for prec in ("fp32", "bf16"):
        m = nb.NeRFMLP(precision=prec)
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        models[prec] = m.to(dev)
    grads = {}
    wss = {}
    for prec, code in (("fp32", nb._lib.PREC_FP32), ("bf16", nb._lib.PREC_BF16)):
        raw, ws = ops.mlp_fwd_rays(models[prec], to, td, z, 1.0, code, True)
        g = torch.zeros(595844, device=dev)
        ops.mlp_bwd(models[prec], d_raw, ws, code, g, S)
        torch.cuda.synchronize()
        print(prec, "fwd+bwd done", flush=True)
        grads[prec], wss[prec] = g, ws
    # fp32 saved activations
    f = wss["fp32"].view(torch.float32)
    off = 0
    X = f[off:off + M * 319].view(M, 319); off += M * 319
    H = {}
    for i in (0, 1, 2, 3, 5, 6, 7):
        H[i] = f[off:off + M * 256].view(M, 256); off += M * 256
    H[4] = X[:, 63:]
    V = f[off:off + M * 283].view(M, 283); off += M * 283
    HV = f[off:off + M * 128].view(M, 128)
    W = {k: torch.from_numpy(v).to(dev) for k, v in p.items()}
    v = ops.bf16_workspace_views(wss["bf16"], M)
    dpre, dhv, xenc, mask = v["dpre"].float(), v["dhv"].float(), v["xenc"].float(), v["mask"]
    print("x_enc save vs fp32: max %.3e" % (xenc[:, :63] - X[:, :63]).abs().max().item())
    act16, hv16, hvm = v["act"].float(), v["hv"].float(), v["hvmask"]
    ar = torch.arange(32, device=dev, dtype=torch.int32)
    bits = {}
    for l in range(8):
        bits[l] = ((mask[l].unsqueeze(-1) >> ar) & 1).reshape(M, 256).bool()
        print(f"mask {l}: mismatches vs fp32 relu {int((bits[l] != (H[l] > 0)).sum())} / {M * 256}; vs own bf16 act>0 {int((bits[l] != (act16[l] > 0)).sum())}")
    hvbits = ((hvm.unsqueeze(-1) >> ar) & 1).reshape(M, 128).bool()
    # torch re-derivation of d(pre-activations) ON THE bf16 FORWARD'S OWN ReLU PATTERN (mask flips near
    # zero pre-activations are a property of the bf16 forward, not an error of the backward kernels)
    dref = {}
    d_hv = (d_raw[:, :3] @ W["rgb_linear.weight"]) * hvbits
    dref["hv"] = d_hv
    d_bott = d_hv @ W["view_linear.weight"][:, :256]
    dref[8] = d_bott
    dh = (d_bott @ W["bottleneck_linear.weight"] + d_raw[:, 3:4] @ W["sigma_linear.weight"]) * bits[7]
    dref[7] = dh
    for l in range(7, 0, -1):
        w = W[f"pts_linears.{l}.weight"]
        w = w[:, 63:] if l == 5 else w
        dh = (dref[l] @ w) * bits[l - 1]
        dref[l - 1] = dh
    # reference gradients from the same pattern and the bf16-saved activations (fp32 math)
    gref = {}
    ins = {0: xenc[:, :63], 5: torch.cat([xenc[:, :63], act16[4]], 1)}
    for l in range(8):
        xin = ins.get(l, act16[l - 1] if l > 0 else None)
        gref[f"pts_linears.{l}.weight"] = dref[l].T @ xin
        gref[f"pts_linears.{l}.bias"] = dref[l].sum(0)
    gref["sigma_linear.weight"] = d_raw[:, 3:4].T @ act16[7]
    gref["sigma_linear.bias"] = d_raw[:, 3].sum(0, keepdim=True)
    gref["bottleneck_linear.weight"] = dref[8].T @ act16[7]
    gref["bottleneck_linear.bias"] = dref[8].sum(0)
    de = v["de"][:R, :27]
    print("de16 save vs fp32 de: max %.3e" % (v["de16"].float()[:, :27] - de.repeat_interleave(S, 0)).abs().max().item())
    vin = torch.cat([act16[8], de.repeat_interleave(S, 0)], 1)
    gref["view_linear.weight"] = d_hv.T @ vin
    gref["view_linear.bias"] = d_hv.sum(0)
    gref["rgb_linear.weight"] = d_raw[:, :3].T @ hv16
    gref["rgb_linear.bias"] = d_raw[:, :3].sum(0)
    print("d_hv   rel %.3e" % rel(dhv, dref["hv"]))
    for l in (8, 7, 6, 5, 4, 3, 2, 1, 0):
        print(f"dpre {l}: rel {rel(dpre[l], dref[l]):.3e}   (ref norm {dref[l].norm().item():.3e})")
    # gradients per tensor
    off = 0
    for name, out_f, in_f in O.LAYER_SHAPES:
        for kind, n in (("weight", out_f * in_f), ("bias", out_f)):
            a, b = grads["bf16"][off:off + n], grads["fp32"][off:off + n]
            print(f"grad {name}.{kind}: vs fp32-mode rel {rel(a, b):.3e} | vs same-pattern reference rel "
                  f"{rel(a, gref[name + '.' + kind]):.3e}  norm {b.norm().item():.3e}")
            off += n
    w5 = slice(int(sum(o * i + o for _, o, i in O.LAYER_SHAPES[:5])), None)
    g5a = grads["bf16"][w5][:256 * 319].view(256, 319)
    g5b = grads["fp32"][w5][:256 * 319].view(256, 319)
    print("  L5 x-part rel %.3e  h-part rel %.3e" % (rel(g5a[:, :63], g5b[:, :63]), rel(g5a[:, 63:], g5b[:, 63:])))


if __name__ == "__main__":
    main()

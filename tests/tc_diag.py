#!/usr/bin/env python
"""Layer-by-layer diagnosis of the tcgen05 forward against the fp32 check mode on the same inputs
(GPU only; a development tool, not a test).  Prints max / mean abs error of every saved activation
so that a descriptor / swizzle / pipeline bug can be located from one GPU run."""
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops
from oracle import nerf_oracle as O


def main(R=300, S=64):
    dev = torch.device("cuda")
    p = O.init_params(0)
    o, d = O.random_rays(R, 1)
    to, td = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    z = ops.stratified_z(torch.linspace(0., 1., S, device=dev), None, R, 2.0, 6.0)
    M = R * S
    models = {}
    for prec in ("fp32", "bf16"):
        m = nb.NeRFMLP(precision=prec)
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        models[prec] = m.to(dev)
    raw32, ws32 = ops.mlp_fwd_rays(models["fp32"], to, td, z, 1.0, nb._lib.PREC_FP32, True)
    torch.cuda.synchronize()
    print("fp32 forward done", flush=True)
    raw16, ws16 = ops.mlp_fwd_rays(models["bf16"], to, td, z, 1.0, nb._lib.PREC_BF16, True)
    torch.cuda.synchronize()
    print("bf16 forward done", flush=True)
    f = ws32.view(torch.float32)
    off = 0
    X = f[off:off + M * 319].view(M, 319); off += M * 319
    H = {}
    for i in (0, 1, 2, 3, 5, 6, 7):
        H[i] = f[off:off + M * 256].view(M, 256); off += M * 256
    H[4] = X[:, 63:]
    V = f[off:off + M * 283].view(M, 283); off += M * 283
    HV = f[off:off + M * 128].view(M, 128)
    v = ops.bf16_workspace_views(ws16, M)
    sv, hv16, vb = v["act"].float(), v["hv"].float(), v["vb"][:R]
    # oracle cross-check of the fp32 path itself
    xe, de = O.encode_samples(o, d, z.cpu().numpy(), O.RenderConfig(N_samples=S))
    ref = O.mlp_forward(p, xe, de)
    print("fp32 raw vs oracle: max %.3e" % np.abs(raw32.cpu().numpy().reshape(-1, 4) - ref).max())
    vb_ref = (de.reshape(R, S, 27)[:, 0] @ p["view_linear.weight"][:, 256:].T + p["view_linear.bias"])
    print("view bias vs oracle: max %.3e" % np.abs(vb.cpu().numpy() - vb_ref).max())
    for g in range(9):
        refa = H[g] if g < 8 else V[:, :256]
        e = (sv[g] - refa).abs()
        print(f"layer {g}: max {e.max().item():.3e} mean {e.mean().item():.3e} | ref absmax {refa.abs().max().item():.3f} "
              f"got absmax {sv[g].abs().max().item():.3f} nan {int(torch.isnan(sv[g]).sum())}")
        if g == 0:
            rows = e.max(1).values
            cols = e.max(0).values
            print("   worst rows", rows.topk(4).indices.tolist(), "row err by 32-block",
                  [round(rows[i:i + 32].max().item(), 4) for i in range(0, 256, 32)])
            print("   col err by 32-block", [round(cols[i:i + 32].max().item(), 4) for i in range(0, 256, 32)])
    e = (hv16 - HV).abs()
    print(f"view  : max {e.max().item():.3e} mean {e.mean().item():.3e}")
    e = (raw16 - raw32).abs().view(-1, 4)
    print("raw   : max per channel", e.max(0).values.tolist(), "mean", e.mean(0).tolist())
    print("raw sample fp32", raw32.view(-1, 4)[:2].tolist(), "bf16", raw16.view(-1, 4)[:2].tolist())


if __name__ == "__main__":
    main()

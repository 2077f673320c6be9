#!/usr/bin/env python
"""Development check (not a test): the fused backward launch (NERF_BWD_ALL) against the two-kernel path
(NERF_BWD_DGRAD then NERF_BWD_WGRAD) on the same saved workspace -- same bf16 d(pre-activations), so the 24 gradient
tensors must agree up to fp32 summation order -- and its CUDA-event time.

    python tests/fused_check.py [R=1024] [S=192] [iters=10]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops
from nerf_mlp_b200._lib import BWD_ALL, BWD_DGRAD, BWD_WGRAD, PREC_BF16
from oracle import nerf_oracle as O


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = nb.NeRFMLP().to(dev)
    m._ensure_flat()
    o = torch.randn(R, 3, device=dev) * 0.1 + torch.tensor([0., 0., 4.], device=dev)
    d = torch.randn(R, 3, device=dev)
    z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0].contiguous()
    d_raw = (torch.randn(R, S, 4, device=dev) * 1e-3).contiguous()
    raw, ws = ops.mlp_fwd_rays(m, o, d, z, 1.0, PREC_BF16, True)
    g2 = torch.zeros_like(m.flat_params)
    ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g2, S, BWD_DGRAD)
    ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g2, S, BWD_WGRAD)
    torch.cuda.synchronize()
    v2 = ops.bf16_workspace_views(ws, R * S)
    dpre2, dhv2 = v2["dpre"].float().clone(), v2["dhv"].float().clone()
    # poison the published tensors so that stale data cannot pass
    g1 = torch.zeros_like(m.flat_params)
    ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g1, S, BWD_ALL)
    torch.cuda.synchronize()
    v1 = ops.bf16_workspace_views(ws, R * S)
    print(f"R={R} S={S}: d(pre-act) images identical: {bool(torch.equal(v1['dpre'].float(), dpre2))}, d_hv identical: "
          f"{bool(torch.equal(v1['dhv'].float(), dhv2))}; flags min/max {int(v1['flags'][:10 * (-(-R * S // 512) * 4)].min())}/{int(v1['flags'][:10 * (-(-R * S // 512) * 4)].max())}")
    worst = 0.0
    off = 0
    for name, out_f, in_f in O.LAYER_SHAPES:
        for kind, n in (("weight", out_f * in_f), ("bias", out_f)):
            a, b = g1[off:off + n].double(), g2[off:off + n].double()
            rel = float((a - b).norm() / (b.norm() + 1e-30))
            worst = max(worst, rel)
            if rel > 1e-4:
                print(f"  MISMATCH {name}.{kind}: rel {rel:.3e} (norms {float(a.norm()):.3e} vs {float(b.norm()):.3e})")
            off += n
    print(f"  worst relative L2 over the 24 tensors: {worst:.3e}  ->  {'OK' if worst <= 1e-4 else 'FAIL'}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for mode, stages in (("fused", (BWD_ALL,)), ("two-kernel", (BWD_DGRAD, BWD_WGRAD))):
        ts = []
        for it in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in stages:
                ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g1, S, st)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"  {mode}: {ms:.3f} ms = {R * S * 2302208 / (ms * 1e-3) / 1e12:.0f} TFLOP/s ({R * S * 2302208 / (ms * 1e-3) / 1e12 / 1691.8 * 100:.1f} % of burst peak)", flush=True)
    return 0 if worst <= 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())

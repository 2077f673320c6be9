"""Development stress: many TrainStep replays, every gradient checked for non-finite / absurd values (and, with `cmp`,
against the two-kernel backward on the same saved tensors)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops
from nerf_mlp_b200._lib import PREC_BF16, BWD_ALL, BWD_DGRAD, BWD_WGRAD
from oracle import nerf_oracle as O
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
mode = sys.argv[3] if len(sys.argv) > 3 else "step"
o, d = O.random_rays(R, 1)
to, td = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
tgt = torch.rand(R, 3, device=dev)
torch.manual_seed(0)
m = nb.NeRFMLP(precision="bf16").to(dev)
bad = 0
if mode == "step":
    r = nb.NeRFRenderer(m, dev, perturb=1.0)
    step = nb.TrainStep(r, nb.FlatAdam(m, lr=5e-4), R)
    for i in range(N):
        step(to, td, tgt)
        g = m._flat_grad
        n = float(g.double().norm())
        if not np.isfinite(n) or n > 1e3 or n == 0.0:
            bad += 1
            nz = (~torch.isfinite(g)).nonzero().flatten()
            print(i, "BAD grad norm", n, "nonfinite", int(nz.numel()), nz[:5].tolist(), step.read_metrics())
            break
else:
    m._ensure_flat()
    S = 192
    z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0].contiguous()
    d_raw = (torch.randn(R, S, 4, device=dev) * 1e-3).contiguous()
    raw, ws = ops.mlp_fwd_rays(m, to, td, z, 1.0, PREC_BF16, True)
    g2 = torch.zeros_like(m.flat_params)
    ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g2, S, BWD_DGRAD)
    ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g2, S, BWD_WGRAD)
    torch.cuda.synchronize()
    flush = torch.empty(200 << 20, dtype=torch.uint8, device=dev)
    for i in range(N):
        if i % 3 == 0:
            flush.zero_()
        g1 = torch.zeros_like(m.flat_params)
        ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g1, S, BWD_ALL)
        rel = float((g1.double() - g2.double()).norm() / g2.double().norm())
        if not np.isfinite(rel) or rel > 1e-3:
            bad += 1
            diff = (g1.double() - g2.double()).abs()
            idx = int(diff.argmax())
            print(i, "BAD rel", rel, "worst idx", idx, float(g1[idx]), float(g2[idx]), "nonfinite", int((~torch.isfinite(g1)).sum()))
            if bad > 5:
                break
print(f"R={R} N={N} mode={mode}: bad={bad}")

#!/usr/bin/env python
"""Development diagnostic: loss trajectories of TrainStep (bf16 / fp32 kernels, eager launches) vs the reference loop on eager CUDA."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from oracle import nerf_oracle as O, nerf_oracle_torch as TP
DEV = "cuda"
R, steps = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 120
perturb = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
p = O.init_params(11)
o, d = O.random_rays(R, 12)
dn = d / np.linalg.norm(d, axis=-1, keepdims=True)
tgt = (0.5 + 0.5 * np.stack([np.sin(3 * dn[:, 0]), np.cos(2 * dn[:, 1]), np.sin(dn[:, 0] + dn[:, 1])], -1)).astype(np.float32)
to, td_, tt = (torch.from_numpy(a).to(DEV) for a in (o, d, tgt))
curves = {}
for prec in ("bf16", "fp32"):
    m = nb.NeRFMLP(precision=prec); m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}); m = m.to(DEV)
    step = nb.TrainStep(nb.NeRFRenderer(m, DEV, perturb=perturb), nb.FlatAdam(m, lr=5e-4), R, graph=False)
    torch.manual_seed(1234)
    c = []
    for _ in range(steps):
        step(to, td_, tt); c.append(step.read_metrics()["loss"])
    curves[prec] = np.array(c)
tr = TP.Trainer(p, device=DEV, perturb=perturb)
torch.manual_seed(1234)
curves["ref"] = np.array([float(tr.step(to, td_, tt)) for _ in range(steps)])
tr2 = TP.Trainer(p, device=DEV, perturb=perturb)
torch.manual_seed(1234)
torch.backends.cuda.matmul.allow_tf32 = True
curves["ref_tf32"] = np.array([float(tr2.step(to, td_, tt)) for _ in range(steps)])
torch.backends.cuda.matmul.allow_tf32 = False
print("step   bf16      fp32      ref       ref_tf32")
for i in list(range(0, 12)) + list(range(12, steps, 6)):
    print(f"{i:4d}  {curves['bf16'][i]:.6f}  {curves['fp32'][i]:.6f}  {curves['ref'][i]:.6f}  {curves['ref_tf32'][i]:.6f}")
for a in ("bf16", "fp32", "ref_tf32"):
    rel = np.abs(curves[a] - curves["ref"]) / curves["ref"]
    print(a, "vs ref: max rel", rel.max(), "mean", rel.mean(), "log-ratio rms", float(np.sqrt(np.mean(np.log(curves[a] / curves['ref']) ** 2))))

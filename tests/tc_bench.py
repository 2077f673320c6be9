#!/usr/bin/env python
"""Micro-benchmark of the fused MLP forward kernel alone (development tool): CUDA-event time of
nerf_mlp_fwd_rays on R rays x S samples, reported as TFLOP/s of algorithmic work."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    save = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = nb.NeRFMLP().to(dev)
    o = torch.randn(R, 3, device=dev) * 0.1 + torch.tensor([0., 0., 4.], device=dev)
    d = torch.randn(R, 3, device=dev)
    z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0].contiguous()
    for _ in range(2):
        ops.mlp_fwd_rays(m, o, d, z, 1.0, nb._lib.PREC_BF16, bool(save))
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mlp_fwd_rays(m, o, d, z, 1.0, nb._lib.PREC_BF16, bool(save))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    tf = R * S * 1186816 / (ms * 1e-3) / 1e12
    print(f"{os.environ.get('NERF_B200_LIB', 'default')}: R={R} S={S} save={save} median {ms:.3f} ms  {tf:.1f} TFLOP/s  "
          f"({tf / 1691.8 * 100:.1f}% of measured bf16 peak)")


if __name__ == "__main__":
    main()

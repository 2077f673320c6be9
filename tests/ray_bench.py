#!/usr/bin/env python
"""Micro-benchmark of the warp-per-ray kernels (development tool): achieved HBM GB/s of compositing forward /
backward and of resample+merge at a launch size that fills the machine (R rays x S samples), against the
algorithmic bytes of SURVEY.md section 8d."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops


ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
WARM = int(sys.argv[2]) if len(sys.argv) > 2 else 3
SIZES = tuple(int(x) for x in sys.argv[3:]) or (16384, 262144)


def t(fn, iters=None):
    iters = iters or ITERS
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda")
    for R in SIZES:
        S_c, N_imp = 64, 128
        S = S_c + N_imp
        g = torch.Generator(dev).manual_seed(0)
        raw = torch.randn(R, S, 4, device=dev, generator=g)
        z = torch.sort(torch.rand(R, S, device=dev, generator=g) * 4 + 2, -1)[0].contiguous()
        d = torch.randn(R, 3, device=dev, generator=g)
        d_rgb = torch.randn(R, 3, device=dev, generator=g)
        zc = z[:, :S_c].contiguous()
        w = torch.rand(R, S_c, device=dev, generator=g)
        u = torch.rand(R, N_imp, device=dev, generator=g)
        ms = t(lambda: ops.composite_fwd(raw, z, d, None, True, True))
        b = R * (S * 24 + 32)
        print(f"R={R:7d} composite_fwd      {ms*1e3:8.1f} us  {b/ms/1e6:7.0f} GB/s (algorithmic {b/1e6:.0f} MB)")
        ms = t(lambda: ops.composite_bwd(raw, z, d, None, True, d_rgb))
        b = R * (S * 36 + 32)
        print(f"R={R:7d} composite_bwd      {ms*1e3:8.1f} us  {b/ms/1e6:7.0f} GB/s (algorithmic {b/1e6:.0f} MB)")
        ms = t(lambda: ops.resample_merge(zc, w, u))
        b = R * (S_c * 8 + N_imp * 4 + S * 4)
        print(f"R={R:7d} resample_merge     {ms*1e3:8.1f} us  {b/ms/1e6:7.0f} GB/s (algorithmic {b/1e6:.0f} MB)")


if __name__ == "__main__":
    main()

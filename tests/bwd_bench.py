#!/usr/bin/env python
"""Development micro-benchmark (not a test): CUDA-event times of the three training MLP kernels -- save-mode forward,
dgrad chain, wgrad -- on R rays x S samples, with and without an L2 flush between launches.  A small problem whose
tile images fit the 126 MB L2 shows what each kernel does when HBM is out of the picture.

    python tests/bwd_bench.py [R=1024] [S=192] [iters=10]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_mlp_b200 as nb
from nerf_mlp_b200 import ops
from nerf_mlp_b200._lib import BWD_DGRAD, BWD_WGRAD, PREC_BF16

FWD, BWD = 1186816, 2302208


def med(ts):
    return sorted(ts)[len(ts) // 2]


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = nb.NeRFMLP().to(dev)
    m._ensure_flat()
    g = m._bind_flat_grads()
    o = torch.randn(R, 3, device=dev) * 0.1 + torch.tensor([0., 0., 4.], device=dev)
    d = torch.randn(R, 3, device=dev)
    z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0].contiguous()
    d_raw = (torch.randn(R, S, 4, device=dev) * 1e-3).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    M = R * S
    for do_flush in (True, False):
        t = {"fwd_save": [], "dgrad": [], "wgrad": [], "bwd_all": []}
        for it in range(iters + 2):
            def timed(name, fn):
                if do_flush:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn()
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    t[name].append(e0.elapsed_time(e1))
                return out
            raw, ws = timed("fwd_save", lambda: ops.mlp_fwd_rays(m, o, d, z, 1.0, PREC_BF16, True))
            timed("dgrad", lambda: ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g, S, BWD_DGRAD))
            timed("wgrad", lambda: ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g, S, BWD_WGRAD))
            timed("bwd_all", lambda: ops.mlp_bwd(m, d_raw, ws, PREC_BF16, g, S))
        fl = {"fwd_save": FWD, "dgrad": BWD - FWD, "wgrad": FWD, "bwd_all": BWD}
        print(f"R={R} S={S} rows={M} ({M * 21.4e3 / 1e6:.0f} MB of tile images) flush={'yes' if do_flush else 'no '}: " +
              "  ".join(f"{k} {med(v):.3f} ms = {M * fl[k] / (med(v) * 1e-3) / 1e12:.0f} TF/s" for k, v in t.items()), flush=True)


if __name__ == "__main__":
    main()

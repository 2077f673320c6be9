#!/usr/bin/env python
"""One replayed training step out of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
the kernels between two consecutive pack_kernel launches near the end of the list, with per-kernel shares.
usage: timeline.py launches.csv > timeline.txt"""
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ev = [(r[k], float(r[v]) / 1e3) for r in rows[1:]]
    # a step = from one metrics+adam (adam_fused_kernel) to the next, taken from the tail (graph replays)
    ends = [i for i, (n, _) in enumerate(ev) if "adam_fused_kernel" in n]
    if len(ends) < 2:
        raise SystemExit("no two optimizer launches in the list")
    step = None
    for i in range(len(ends) - 1, 0, -1):          # last step whose coarse pass is the full network (<fwd, no save, not sigma-only>)
        cand = ev[ends[i - 1] + 1:ends[i] + 1]
        if any("mlp_tc_kernel<0, 0, 2, 0>" in n or "mlp_tc_kernel<0, 0, 1, 0>" in n for n, _ in cand):
            step = cand
            break
    if step is None:
        step = ev[ends[-2] + 1:ends[-1] + 1]
    step = [(n, t) for n, t in step if "FillFunctor<unsigned char>" not in n]      # bench.py's untimed L2 flush between steps
    tot = sum(t for _, t in step)
    print("# One replayed training step (1024 rays, 64+128, full coarse pass) from the ncu launch list of")
    print("# `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (gpu__time_duration per kernel node of the CUDA graph;")
    print("# serialised, cold-cache: use the SHARES, not the absolutes)")
    print("# (the 256 MB L2-flush fill that bench.py issues between steps, outside the timed events, is left out)")
    print(f"# kernels in the step: {len(step)}   sum of durations: {tot:.1f} us")
    for n, t in step:
        print(f"{t:9.1f} us  {100 * t / tot:5.1f}%  {n[:96]}")


if __name__ == "__main__":
    main(sys.argv[1])

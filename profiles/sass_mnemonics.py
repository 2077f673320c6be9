#!/usr/bin/env python
"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA use, per kernel, from the in-tree objects
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk async copies -> UBLKCP,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS).  Usage: python profiles/sass_mnemonics.py > profiles/rXX_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys

OBJS = ["nerf_mlp_b200/csrc/nerf_mlp_tc.o", "nerf_mlp_b200/csrc/nerf_mlp_wgrad.o"]
PAT = re.compile(r"\b(UTC\w*MMA|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|SYNCS|REDG?|RED)\b")

print("# kernel | mnemonic counts (cuobjdump -sass, sm_100a)")
for obj in OBJS:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn and "/*" in line:
            m = PAT.search(line)
            if m:
                counts[fn][m.group(1)] += 1
    for fn, c in counts.items():
        if c:
            print(f"{obj.split('/')[-1]} | {fn[:110]} | " + ", ".join(f"{k}={v}" for k, v in sorted(c.items())))

#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass): the record that the hot kernels are hand-written
sm_100a code -- UBLKCP (TMA bulk copies), SYNCS (mbarrier), CREDUX (warp min/max reductions, sm_100a only), REDG...F32x4
(vector reductions), LDS.128 -- and hold no library or tensor-core instructions.
  python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "celllistmap.jl_b200", "libclm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["cu++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
WATCH = ("UBLKCP", "SYNCS", "CREDUX", "REDG", "RED.", "LDS.128", "ATOMG", "MUFU", "HMMA", "UTCMMA", "R2UR", "LDGSTS")
print(f"# opcode histogram of {os.path.relpath(so, ROOT)} (cuobjdump -sass), one block per kernel: total instructions, the")
print("# Blackwell-specific / memory-path opcodes, then the ten most frequent opcodes")
for (k, c), name in zip(hist.items(), demangle):
    tot = sum(c.values())
    if tot < 150 and "k_sweep" not in name:
        continue
    watch = {op: n for op, n in c.items() if op.startswith(WATCH)}
    print(f"\n{name[:230]}\n  {tot} instructions; " + ", ".join(f"{op} {n}" for op, n in sorted(watch.items())))
    print("  top: " + ", ".join(f"{op} {n}" for op, n in c.most_common(10)))

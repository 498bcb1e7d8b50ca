#!/usr/bin/env python
"""Per-source-line instruction and stall accounting of one kernel from an ncu report.

ncu's `--page source --csv` lists the kernel's SASS with `Instructions Executed` and stall samples per instruction;
`nvdisasm -g` of the cubin gives the source line of every instruction (inlined code: the innermost line).  This joins
the two by instruction order and aggregates by (file, line).

  python tools/sass_lines.py <report.ncu-rep> <object.o|lib.so> <mangled-kernel-substring> [top] [launch-index]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
lines_of = None
for cub in sorted(os.listdir(tmp)):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur, inside, out = ("?", 0), False, []
    for l in dis.splitlines():
        if l.startswith(".text.") and l.rstrip().endswith(":"):
            inside = pat in l
            if inside and out:
                break
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    if out:
        lines_of = out
        break
if not lines_of:
    sys.exit(f"kernel matching {pat!r} not found in {obj}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# the report may hold several kernels: blocks start with a "Kernel Name" row
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
want = int(sys.argv[5]) if len(sys.argv) > 5 else 0
blk = [b for b in blocks if len(b["rows"]) == len(lines_of)]
if not blk:
    sys.exit(f"no kernel in the report has {len(lines_of)} instructions (report kernels: {[(b['name'][:60], len(b['rows'])) for b in blocks]})")
blk = blk[want]
h = blk["hdr"]
ci, cs = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for (off, key, text), r in zip(lines_of, blk["rows"]):
    n, s = int(r[ci] or 0), int(r[cs] or 0)
    a = agg[key]
    a[0] += n; a[1] += s; a[2] += 1
    tot_i += n; tot_s += s
print(f"# {blk['name'][:150]}")
print(f"# {len(lines_of)} SASS instructions, {tot_i:.4g} warp instructions executed, {tot_s} stall samples")
print(f"{'file:line':34s} {'sass':>5s} {'warp instr':>12s} {'share':>7s} {'stall smp':>10s} {'share':>7s}")
srcs = {}
for (fn, ln), (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    for d in ("celllistmap.jl_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, fn)
        if os.path.exists(p):
            srcs.setdefault(p, open(p).read().splitlines())
            if 0 < ln <= len(srcs[p]):
                text = srcs[p][ln - 1].strip()[:90]
    print(f"{fn + ':' + str(ln):34s} {k:5d} {n:12d} {100.0 * n / max(tot_i, 1):6.2f}% {s:10d} {100.0 * s / max(tot_s, 1):6.2f}%  {text}")

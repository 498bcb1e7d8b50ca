#!/usr/bin/env python
"""Newton's-third-law force sweep (k_sweep_n3) next to the full-shell sweep and the oracle on the C2 workload:
errors and kernel times.  Usage: python tools/check_n3.py [nside] [f32|f64|both]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402
from oracle import oracle as om  # noqa: E402  (checker only)

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
which = sys.argv[2] if len(sys.argv) > 2 else "both"
for dtype in [d for d, k in ((np.float32, "f32"), (np.float64, "f64")) if which in (k, "both")]:
    w = W.c2_argon(nside, dtype)
    n = w["x"].shape[0]
    o = om.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64))
    we, wf = o.lj(w["c6"], w["c12"], forces=True, nbatches=om.lib().ora_num_threads())
    res = {}
    for n3 in (0, 1):
        h = clm.Handle(3, dtype)
        h.set_option("n3", n3)
        h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
        h.set_positions(0, w["x"])
        e, f = np.zeros(1, dtype), np.zeros((n, 3), dtype)
        ms = []
        for _ in range(6):
            h.map_lj(w["c6"], w["c12"], e, f, profile=True)
            ms.append(h.stats().sweep_ms)
        res[n3] = (float(e[0]), f.copy(), min(ms[2:]), h.stats().map_ms)
        h.close()
    for n3 in (0, 1):
        e, f, ms, map_ms = res[n3]
        print(f"{np.dtype(dtype).name} n3={n3}: sweep {ms:.4f} ms (map {map_ms:.4f})  |E-Eo|/|Eo| = {abs(e - we) / abs(we):.3e}  "
              f"max|F-Fo|/max|Fo| = {np.abs(f - wf).max() / np.abs(wf).max():.3e}  sum F / (max F sqrt n) = {np.abs(f.astype(np.float64).sum(0)).max() / (np.abs(wf).max() * np.sqrt(n)):.3e}", flush=True)
    print(f"   n3 vs full shell: max|dF|/max|F| = {np.abs(res[1][1].astype(np.float64) - res[0][1]).max() / np.abs(wf).max():.3e}")

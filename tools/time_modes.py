#!/usr/bin/env python
"""Device times of the exactly-once (MODE_HALF) maps on the C2 workload next to the full-shell force map.
Usage: python tools/time_modes.py [nside]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm
import workloads as W
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for dtype in (np.float32, np.float64):
    w = W.c2_argon(nside, dtype)
    n = w["x"].shape[0]
    h = clm.Handle(3, dtype)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    h.set_positions(0, w["x"]); h.build()
    e, f = np.zeros(1, dtype), np.zeros((n, 3), dtype)
    sd, sd2, npairs = np.zeros(1, dtype), np.zeros(1, dtype), np.zeros(1, np.int64)
    i, j, d = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, dtype)
    runs = {
        "LJ energy+forces (full shell)": lambda: h.map_lj(w["c6"], w["c12"], e, f, profile=True),
        "LJ energy (exactly once)": lambda: h.map_lj(w["c6"], w["c12"], e, None, profile=True),
        "sum d, d2 (exactly once)": lambda: h.map_sum_d_d2(sd, sd2, npairs, profile=True),
        "minimum distance (exactly once)": lambda: h.map_mindist(i, j, d, profile=True),
    }
    for name, fn in runs.items():
        for _ in range(4):
            fn()
        print(f"{np.dtype(dtype).name} {name:34s} sweep {h.stats().sweep_ms:.3f} ms", flush=True)
    h.close()

#!/usr/bin/env python
"""Device time of the distance-histogram map on the C2 workload (1M argon-density particles).  Usage: python tools/time_hist.py [nbins]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm
import workloads as W
nbins = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for dtype in (np.float32, np.float64):
    w = W.c2_argon(100, dtype)
    h = clm.Handle(3, dtype)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    h.set_positions(0, w["x"]); h.build()
    c = np.zeros(nbins, np.int64)
    for it in range(4):
        h.map_dist_hist(w["cutoff"] / nbins, c, profile=True)
    st = h.stats()
    print(f"distance histogram {np.dtype(dtype).name} nbins={nbins}: pairs {int(c.sum())} sweep {st.sweep_ms:.3f} ms -> {c.sum() / (st.sweep_ms * 1e-3):.3e} pair-evals/s", flush=True)
    h.close()

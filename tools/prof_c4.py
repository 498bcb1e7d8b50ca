#!/usr/bin/env python
"""Profiling driver: pair-velocity map of the C4 workload (halotools-style, Float64) at a reduced size.
Usage: python tools/prof_c4.py [n] [dim] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = W.c4_galaxies(n, dim)
h = clm.Handle(dim, np.float64)
h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
h.set_positions(0, w["x"])
h.build()
c, sm = np.zeros(5, np.int64), np.zeros(5)
for it in range(reps):
    h.map_pairvel(w["v"], None, w["rbins"], c, sm, profile=True)
    st = h.stats()
    print(f"C4 {dim}-D n={n}: pairs {int(c.sum())} sweep {st.sweep_ms:.3f} ms build {st.build_ms:.3f} ms tiles {st.n_tiles} cells {st.n_cells} -> {c.sum() / (st.sweep_ms * 1e-3):.3e} pair-evals/s", flush=True)
h.close()

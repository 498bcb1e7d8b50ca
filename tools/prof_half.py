#!/usr/bin/env python
"""Profiling driver: the exactly-once (MODE_HALF) LJ energy map on the C2 workload.  Usage: python tools/prof_half.py [f32|f64] [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm
import workloads as W
dtype = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = W.c2_argon(100, dtype)
h = clm.Handle(3, dtype)
h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
h.set_positions(0, w["x"]); h.build()
e = np.zeros(1, dtype)
for it in range(reps):
    h.map_lj(w["c6"], w["c12"], e, None, profile=True)
    print(f"LJ energy (exactly once) {np.dtype(dtype).name}: sweep {h.stats().sweep_ms:.4f} ms E={e[0]:.6e}", flush=True)
h.close()

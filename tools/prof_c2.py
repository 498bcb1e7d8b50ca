#!/usr/bin/env python
"""Profiling driver: a few build + LJ energy+forces steps of the C2 workload (for ncu captures of k_sweep / k_bin).
Usage: python tools/prof_c2.py [nside] [f32|f64] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dtype = np.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else np.float32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
w = W.c2_argon(nside, dtype)
n = w["x"].shape[0]
tdt = torch.float32 if dtype == np.float32 else torch.float64
x_dev = torch.from_numpy(w["x"]).cuda()
e_dev = torch.zeros(1, dtype=tdt, device="cuda")
f_dev = torch.zeros((n, 3), dtype=tdt, device="cuda")
h = clm.Handle(3, dtype)
h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
if os.environ.get("CLM_N3"):
    h.set_option("n3", int(os.environ["CLM_N3"]))   # 1: Newton's-third-law sweep (default for Float32 only), 0: full shell
for it in range(steps):
    h.set_positions(0, x_dev)
    h.build()
    h.map_lj(w["c6"], w["c12"], e_dev, None if os.environ.get("CLM_ENERGY_ONLY") else f_dev, reset=True, profile=True)   # CLM_ENERGY_ONLY=1: the energy map
    st = h.stats()
    print(f"step {it}: sweep {st.sweep_ms:.4f} ms build {st.build_ms:.4f} ms tiles {st.n_tiles} E={float(e_dev[0]):.6e}", flush=True)
h.close()

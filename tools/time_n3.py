#!/usr/bin/env python
"""Kernel times of the LJ energy+forces map on the C2 workload: Newton's-third-law sweep (k_sweep_n3) and full-shell sweep,
for the library named by CLM_SO (variants built with CLM_NVCC_EXTRA).  No oracle: see tools/check_n3.py for the errors.
Usage: python tools/time_n3.py [nside] [f32|f64|both]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
which = sys.argv[2] if len(sys.argv) > 2 else "f32"
for dtype in [d for d, k in ((np.float32, "f32"), (np.float64, "f64")) if which in (k, "both")]:
    w = W.c2_argon(nside, dtype)
    n = w["x"].shape[0]
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    x_dev = torch.from_numpy(w["x"]).cuda()
    e_dev = torch.zeros(1, dtype=tdt, device="cuda")
    f_dev = torch.zeros((n, 3), dtype=tdt, device="cuda")
    for n3 in (1, 0, 2):     # 2: the Newton's-third-law sweep without the energy (energy_out = NULL)
        h = clm.Handle(3, dtype)
        h.set_option("n3", min(n3, 1))
        h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
        ts, bs, ms = [], [], []
        for it in range(12):
            h.set_positions(0, x_dev)
            h.map_lj(w["c6"], w["c12"], None if n3 == 2 else e_dev, f_dev, reset=True, profile=True)
            st = h.stats()
            if it >= 4:
                ts.append(st.sweep_ms); bs.append(st.build_ms); ms.append(st.map_ms)
        print(f"{os.path.basename(os.environ.get('CLM_SO', 'libclm_b200.so'))} {np.dtype(dtype).name} n3={n3}: sweep {np.median(ts):.4f} ms  build {np.median(bs):.4f} ms  "
              f"map {np.median(ms):.4f} ms  E={float(e_dev[0]):.7e}  |F|max={float(f_dev.abs().max()):.5e}", flush=True)
        h.close()

#!/usr/bin/env python
"""Kernel time of the LJ ENERGY-ONLY map (exactly once per pair) on the C2 workload: the lean partner-per-lane sweep
(k_sweep_n3 with a functor without force outputs; clm_set_option "n3" != 0) against the exactly-once k_sweep<MODE_HALF>
(option "n3" = 0), for the library named by CLM_SO (variants built with CLM_NVCC_EXTRA).
Usage: python tools/time_lj_energy.py [nside] [f32|f64|both]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
which = sys.argv[2] if len(sys.argv) > 2 else "both"
for dtype in [d for d, k in ((np.float32, "f32"), (np.float64, "f64")) if which in (k, "both")]:
    w = W.c2_argon(nside, dtype)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    x_dev = torch.from_numpy(w["x"]).cuda()
    e_dev = torch.zeros(1, dtype=tdt, device="cuda")
    res = {}
    for n3 in (1, 0):
        h = clm.Handle(3, dtype)
        h.set_option("n3", n3)
        h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
        ts, bs = [], []
        for it in range(12):
            h.set_positions(0, x_dev)
            h.map_lj(w["c6"], w["c12"], e_dev, None, reset=True, profile=True)
            st = h.stats()
            if it >= 4:
                ts.append(st.sweep_ms); bs.append(st.build_ms)
        res[n3] = float(e_dev[0])
        print(f"{os.path.basename(os.environ.get('CLM_SO', 'libclm_b200.so'))} {np.dtype(dtype).name} LJ energy only, n3={n3}: sweep {np.median(ts):.4f} ms  "
              f"build {np.median(bs):.4f} ms  E={res[n3]:.10e}", flush=True)
        h.close()
    print(f"   lean vs MODE_HALF energy: relative difference {abs(res[1] - res[0]) / abs(res[0]):.3e}", flush=True)

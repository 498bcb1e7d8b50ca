#!/usr/bin/env python
"""GPU tuning sweep: LJ-force sweep-kernel time of the C2 workload over the device-grid split (`sub`) and the
tile size (`tile_i`).  Usage: python tools/tune_sweep.py [nside] [f32|f64]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dtypes = [np.float32, np.float64] if len(sys.argv) <= 2 else [np.float32 if sys.argv[2] == "f32" else np.float64]
for dtype in dtypes:
    w = W.c2_argon(nside, dtype)
    n = w["x"].shape[0]
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    x_dev = torch.from_numpy(w["x"]).cuda()
    e_dev = torch.zeros(1, dtype=tdt, device="cuda")
    f_dev = torch.zeros((n, 3), dtype=tdt, device="cuda")
    for sub in (1, 2, 3, 4):
        for ti in (8,):
            h = clm.Handle(3, dtype)
            h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
            h.set_option("sub", sub)
            h.set_positions(0, x_dev)
            h.build()
            ts, bs = [], []
            for it in range(6):
                h.set_positions(0, x_dev)
                h.build()
                h.map_lj(w["c6"], w["c12"], e_dev, f_dev, reset=True, profile=True)
                st = h.stats()
                if it >= 2:
                    ts.append(st.sweep_ms)
                    bs.append(st.build_ms)
            print(f"{np.dtype(dtype).name} sub={sub} tile_i={ti:2d}: sweep {np.mean(ts):8.3f} ms  build {np.mean(bs):7.3f} ms  "
                  f"tiles {st.n_tiles}  cells {st.n_cells}  E={float(e_dev[0]):.6e}", flush=True)
            h.close()

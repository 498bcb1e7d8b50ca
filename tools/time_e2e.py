#!/usr/bin/env python
"""End-to-end step times of the C2 workload through the C ABI with pinned HOST buffers: synchronous calls against
pipelined frames (clm_set_positions_async + CLM_ASYNC).  Usage: python tools/time_e2e.py [nside] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for dtype in (np.float32, np.float64):
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    w = W.c2_argon(nside, dtype)
    n = w["x"].shape[0]
    h = clm.Handle(3, dtype)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    xs = [torch.from_numpy(w["x"]).pin_memory() for _ in range(2)]
    fs = [torch.zeros((n, 3), dtype=tdt).pin_memory() for _ in range(2)]
    es = [torch.zeros(1, dtype=tdt).pin_memory() for _ in range(2)]
    for _ in range(3):
        h.set_positions(0, xs[0].numpy())
        h.map_lj(w["c6"], w["c12"], es[0].numpy(), fs[0].numpy())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        h.set_positions(0, xs[k & 1].numpy())
        h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy())
    torch.cuda.synchronize()
    t_sync = (time.perf_counter() - t0) / steps
    e_sync = float(es[(steps - 1) & 1][0])
    for k in range(4):
        h.set_positions_async(0, xs[k & 1].numpy())
        h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy(), async_=True)
    h.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        h.set_positions_async(0, xs[k & 1].numpy())
        h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy(), async_=True)
    h.synchronize()
    t_pipe = (time.perf_counter() - t0) / steps
    print(f"{np.dtype(dtype).name}: synchronous {1e3 * t_sync:.3f} ms/step, pipelined {1e3 * t_pipe:.3f} ms/step, "
          f"energies {e_sync:.6e} / {float(es[(steps - 1) & 1][0]):.6e}", flush=True)
    h.close()

#!/usr/bin/env python
"""neighbour-list build time (BASELINE metric 2): C1 (10k, cutoff 0.1, F64) and larger unit-cube systems."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import celllistmap_b200 as clm
import workloads as W

for n, cutoff in ((10_000, 0.1), (100_000, 0.1), (1_000_000, 0.03)):
    w = W.c1_neighborlist(n)
    x = w["x"]
    nb = clm.InPlaceNeighborList(x=x, cutoff=cutoff, unitcell=w["unitcell"])
    h = nb.sys._h
    for _ in range(3):
        clm.update(nb, xpositions=x)
        lst = nb.neighborlist()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        clm.update(nb, xpositions=x)          # new coordinates -> H2D + UpdateCellList! + sweep + emission + D2H of the records
        lst = nb.neighborlist()
        ts.append(time.perf_counter() - t0)
    # device-resident: positions already on the GPU, list left on the GPU
    xd = torch.from_numpy(x).cuda()
    td = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.set_positions(0, xd)
        h.build()
        cnt = h.neighborlist_count(profile=True)
        td.append(time.perf_counter() - t0)
    st = h.stats()
    print(f"n={n} cutoff={cutoff}: pairs={len(lst)}  e2e {1e3*np.median(ts):.3f} ms  device-resident {1e3*np.median(td):.3f} ms  "
          f"(build {st.build_ms:.3f} ms, emission sweep {st.sweep_ms:.3f} ms, map {st.map_ms:.3f} ms)", flush=True)

#!/usr/bin/env python
"""Device times of BASELINE.json configs 3 and 4 (parity-test configurations; not bench lines): build + map through the
C ABI with host inputs resident, CUDA-event timings from clm_stats."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm
import workloads as W


def timed(fn, h, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    st = h.stats()
    return 1e3 * np.median(ts), st.build_ms, st.sweep_ms

w = W.c3_triclinic_cross(1_000_000, 1_000_000)
h = clm.Handle(3, np.float64)
h.set_box(clm._capi.TRICLINIC, w["unitcell"], w["cutoff"], 1)
h.set_positions(0, w["x"]); h.set_positions(1, w["y"]); h.build()
i, j, d = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1)
sd, sd2, n = np.zeros(1), np.zeros(1), np.zeros(1, np.int64)
h.map_sum_d_d2(sd, sd2, n)
t, b, s = timed(lambda: h.map_mindist(i, j, d, profile=True), h)
print(f"C3 triclinic cross 1M x 1M min-distance F64: pairs {int(n[0])}  map call {t:.3f} ms  sweep kernel {s:.3f} ms  build {b:.3f} ms  -> {int(n[0]) / (s * 1e-3):.3e} pair-evals/s (kernel)", flush=True)
h.close()
for dim in (3, 2):
    w = W.c4_galaxies(4_000_000, dim)
    h = clm.Handle(dim, np.float64)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    h.set_positions(0, w["x"]); h.build()
    c, sm = np.zeros(5, np.int64), np.zeros(5)
    t, b, s = timed(lambda: h.map_pairvel(w["v"], None, w["rbins"], c, sm, profile=True), h, reps=3)
    npairs = int(c.sum())
    print(f"C4 pair-velocity 4M galaxies {dim}-D F64: pairs {npairs}  map call {t:.3f} ms  sweep kernel {s:.3f} ms  build {b:.3f} ms  -> {npairs / (s * 1e-3):.3e} pair-evals/s (kernel)", flush=True)
    h.close()

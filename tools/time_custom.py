#!/usr/bin/env python
"""Run-time compiled user pair functions next to the compiled-in catalogue on the C2 workload (1M particles):
LJ energy (exactly once) and LJ energy + forces (full shell), written as CUDA C++ source and compiled with NVRTC."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celllistmap_b200 as clm
import workloads as W

ENERGY = """
struct UserLJ {
    static constexpr int NSCALAR = 1, NPART = 0, NAUX = 0, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        const T r2 = T(1) / p.d2, r6 = r2 * r2 * r2;
        out.add_scalar(0, r6 * (par[1] * r6 - par[0]));
    }
};
"""
FORCES = """
struct UserLJForces {   // u = c12/d^12 - c6/d^6;  f_i = -(12 c12/d^14 - 6 c6/d^8) (x_j - x_i)
    static constexpr int NSCALAR = 1, NPART = 3, NAUX = 0, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        const T r2 = T(1) / p.d2, r6 = r2 * r2 * r2;
        out.add_scalar(0, r6 * (par[1] * r6 - par[0]));
        const T fs = r2 * r6 * (T(12) * par[1] * r6 - T(6) * par[0]);
        for (int k = 0; k < 3; ++k) out.add_i(k, -fs * (p.y[k] - p.x[k]));
    }
};
"""
for dtype in (np.float32, np.float64):
    w = W.c2_argon(100, dtype)
    n = w["x"].shape[0]
    h = clm.Handle(3, dtype)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    h.set_positions(0, w["x"]); h.build()
    e, f = np.zeros(1, dtype), np.zeros((n, 3), dtype)
    t0 = time.perf_counter(); fe, _ = h.custom_compile(ENERGY, "UserLJ"); t1 = time.perf_counter(); ff, _ = h.custom_compile(FORCES, "UserLJForces"); t2 = time.perf_counter()
    sc, pp = np.zeros(1, dtype), np.zeros((n, 3), dtype)
    res = {}
    for name, fn in (("catalogue LJ energy", lambda: h.map_lj(w["c6"], w["c12"], e, None, profile=True)),
                     ("user      LJ energy", lambda: h.map_custom(fe, (w["c6"], w["c12"]), scalars=sc, profile=True)),
                     ("catalogue LJ energy+forces", lambda: h.map_lj(w["c6"], w["c12"], e, f, profile=True)),
                     ("user      LJ energy+forces", lambda: h.map_custom(ff, (w["c6"], w["c12"]), scalars=sc, per_particle=pp, profile=True))):
        for _ in range(4):
            fn()
        res[name] = h.stats().sweep_ms
        print(f"{np.dtype(dtype).name} {name:28s} sweep {res[name]:.3f} ms", flush=True)
    print(f"{np.dtype(dtype).name} NVRTC compile: {t1 - t0:.2f} s + {t2 - t1:.2f} s (first includes opening libnvrtc); user vs catalogue energy {sc[0]:.6e} / {e[0]:.6e}; "
          f"max |f_user - f_cat| / max|f| = {np.abs(pp - f).max() / np.abs(f).max():.2e}", flush=True)
    h.close()

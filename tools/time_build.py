#!/usr/bin/env python
"""Cell-list build time (UpdateCellList!: k_zero_ints, k_bin<count>, k_rows, k_bin<scatter>, k_twin) against the grid cap of
k_bin (clm_set_option "bin_blocks_per_sm": 0 = one block per 256 particles, N = at most N blocks per SM striding over the
particles).  CUDA events around the build (clm_stats.build_ms), L2 flushed in front of every step.
Usage: python tools/time_build.py [nside ...]        (100 -> 1M particles, 200 -> 8M)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import celllistmap_b200 as clm  # noqa: E402
import workloads as W  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [100, 200]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for nside in sizes:
    for dtype in (np.float32, np.float64):
        if nside > 100 and dtype == np.float64:
            continue
        w = W.c2_argon(nside, dtype)
        n = w["x"].shape[0]
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        x_dev = torch.from_numpy(w["x"]).cuda()
        e_dev = torch.zeros(1, dtype=tdt, device="cuda")
        f_dev = torch.zeros((n, 3), dtype=tdt, device="cuda")
        for cap in (0, 4, 8, 16, 0, 8):
            h = clm.Handle(3, dtype)
            h.set_option("bin_blocks_per_sm", cap)
            h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
            bs, ss, st_ms = [], [], []
            for it in range(14):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                h.set_positions(0, x_dev)
                h.map_lj(w["c6"], w["c12"], e_dev, f_dev, reset=True, profile=True)
                b.record()
                st = h.stats()
                torch.cuda.synchronize()
                if it >= 4:
                    bs.append(st.build_ms); ss.append(st.sweep_ms); st_ms.append(a.elapsed_time(b))
            print(f"n={n} {np.dtype(dtype).name} bin_blocks_per_sm={cap}: build {np.median(bs):.4f} ms (min {min(bs):.4f})  sweep {np.median(ss):.4f} ms  "
                  f"step(profiled) {np.median(st_ms):.4f} ms  E={float(e_dev[0]):.7e}", flush=True)
            h.close()

#!/usr/bin/env python
"""Pipelined end-to-end frames alternating between TWO handles (two ParticleSystems): the frames of a trajectory are
independent, so the cell-list build of frame k+1 (handle B, its own streams) can run next to the sweep of frame k (handle A)
if the sweep leaves room on the SMs (clm_set_option blocks_per_sm).  Wall clock per frame, C2 workload, with the L2 flush."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import celllistmap_b200 as clm
import workloads as W
dtype = np.float32; tdt = torch.float32
w = W.c2_argon(100, dtype); n = w["x"].shape[0]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
flush.zero_()
steps = 60
for nh in (1, 2, 3):
    for bps in ((0,) if nh == 1 else (0, -1, -2)):
        hs, sts = [], []
        for k in range(nh):
            h = clm.Handle(3, dtype)
            st = torch.cuda.Stream()
            h.set_stream(st.cuda_stream)
            h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
            if bps:
                h.set_option("blocks_per_sm", bps)
            hs.append(h); sts.append(st)
        xs = [torch.from_numpy(w["x"]).pin_memory() for _ in range(2 * nh)]
        fs = [torch.zeros((n, 3), dtype=tdt).pin_memory() for _ in range(2 * nh)]
        es = [torch.zeros(1, dtype=tdt).pin_memory() for _ in range(2 * nh)]
        def frame(k, fl):
            a = k % nh; b = (k // nh) & 1; q = a * 2 + b
            with torch.cuda.stream(sts[a]):
                if fl: flush.zero_()
                hs[a].set_positions_async(0, xs[q].numpy())
                hs[a].map_lj(w["c6"], w["c12"], es[q].numpy(), fs[q].numpy(), async_=True)
        for k in range(8): frame(k, 0)
        for h in hs: h.synchronize()
        for fl in (0, 1):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for k in range(steps): frame(k, fl)
            for h in hs: h.synchronize()
            torch.cuda.synchronize()
            print("handles=%d blocks_per_sm=%d flush=%d: wall/frame %.3f ms  E=%.7e" % (nh, bps, fl, 1e3 * (time.perf_counter() - t0) / steps, float(es[0][0])), flush=True)
        for h in hs: h.close()

# device-resident: the same alternation with positions and outputs on the device (no copies): throughput of independent steps
xd = torch.from_numpy(w["x"]).cuda()
for nh, bps in ((1, 0), (2, 0), (2, -1), (3, -1), (3, -2)):
    hs, sts, fd, ed = [], [], [], []
    for k in range(nh):
        h = clm.Handle(3, dtype); st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
        h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
        if bps: h.set_option("blocks_per_sm", bps)
        hs.append(h); sts.append(st)
        fd.append(torch.zeros((n, 3), dtype=tdt, device="cuda")); ed.append(torch.zeros(1, dtype=tdt, device="cuda"))
    def step(k, fl):
        a = k % nh
        with torch.cuda.stream(sts[a]):
            if fl: flush.zero_()
            hs[a].set_positions(0, xd)
            hs[a].map_lj(w["c6"], w["c12"], ed[a], fd[a])
    for k in range(8): step(k, 0)
    torch.cuda.synchronize()
    for fl in (0, 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(steps): step(k, fl)
        torch.cuda.synchronize()
        print("device-resident handles=%d blocks_per_sm=%d flush=%d: wall/step %.3f ms  E=%.7e" % (nh, bps, fl, 1e3 * (time.perf_counter() - t0) / steps, float(ed[0][0])), flush=True)
    for h in hs: h.close()

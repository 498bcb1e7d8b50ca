#!/usr/bin/env python
"""Where the pipelined end-to-end step goes: wall clock per frame of clm_set_positions_async + clm_map_lj(CLM_ASYNC) on the C2
workload with the copies switched off one at a time (clm_set_option "dbg": 2 = no force copy-out, 4 = no copy-in) and with /
without the 256 MiB L2 flush bench.py puts between frames, next to the device-resident step."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import celllistmap_b200 as clm
import workloads as W
dtype = np.float32; tdt = torch.float32
w = W.c2_argon(100, dtype); n = w["x"].shape[0]
h = clm.Handle(3, dtype)
st = torch.cuda.Stream()
h.set_stream(st.cuda_stream)
h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
xs = [torch.from_numpy(w["x"]).pin_memory() for _ in range(2)]
fs = [torch.zeros((n, 3), dtype=tdt).pin_memory() for _ in range(2)]
es = [torch.zeros(1, dtype=tdt).pin_memory() for _ in range(2)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
steps = 60
with torch.cuda.stream(st):
    flush.zero_()
    for k in range(6):
        h.set_positions_async(0, xs[k & 1].numpy()); h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy(), async_=True)
    h.synchronize()
    for dbg in (0, 8, 2, 4):
        for fl in (0, 1):
            h.set_option("dbg", dbg)
            host = []
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for k in range(steps):
                a = time.perf_counter()
                if fl: flush.zero_()
                h.set_positions_async(0, xs[k & 1].numpy())
                h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy(), async_=True, profile=True)
                host.append(time.perf_counter() - a)
            h.synchronize(); torch.cuda.synchronize()
            tot = (time.perf_counter() - t0) / steps
            s_ = h.stats()
            print("dbg=%d (%s%s) flush=%d: wall/frame %.3f ms   host enqueue/frame median %.3f ms   last frame: build %.3f sweep %.3f map %.3f ms" % (dbg, "no D2H " if dbg & 2 else ("D2H at once " if dbg & 8 else ""), "no H2D" if dbg & 4 else "", fl, 1e3 * tot, 1e3 * np.median(host), s_.build_ms, s_.sweep_ms, s_.map_ms), flush=True)
    h.set_option("dbg", 0)
    xd = torch.from_numpy(w["x"]).cuda(); fd = torch.zeros((n, 3), dtype=tdt, device="cuda"); ed = torch.zeros(1, dtype=tdt, device="cuda")
    for fl in (0, 1):
        for k in range(5):
            h.set_positions(0, xd); h.map_lj(w["c6"], w["c12"], ed, fd)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(steps):
            if fl: flush.zero_()
            h.set_positions(0, xd); h.map_lj(w["c6"], w["c12"], ed, fd)
        torch.cuda.synchronize()
        print("device-resident flush=%d: wall/step %.3f ms" % (fl, 1e3 * (time.perf_counter() - t0) / steps), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for k in range(20): flush.zero_()
    e1.record(st); torch.cuda.synchronize()
    print("flush alone: %.4f ms" % (e0.elapsed_time(e1) / 20))

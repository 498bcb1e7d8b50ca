#!/usr/bin/env python
"""CPU restatement (oracle port, all host threads) on BASELINE.json configs 1-4: build + map seconds and pair-evals/s.
The oracle is test infrastructure; this tool only reports the CPU baseline that SURVEY.md §8(d) asks for next to the GPU
numbers of tools/time_configs.py / bench.py.  Usage: python tools/cpu_configs.py [--small]"""
import os, sys, time, platform
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as om
import workloads as W

small = "--small" in sys.argv
nt = om.lib().ora_num_threads()
cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
print(f"CPU restatement of CellListMap.jl (C++/OpenMP, not Julia), {nt} threads, {cpu[0] if cpu else platform.processor()}", flush=True)


def timed(label, make, run, reps=3, count=None):
    """median build / map seconds; `count` (untimed) returns the in-cutoff pairs when `run` does not"""
    tb, tm, pairs = [], [], 0
    for _ in range(reps):
        t0 = time.perf_counter(); o = make(); t1 = time.perf_counter(); pairs = run(o); t2 = time.perf_counter()
        tb.append(t1 - t0); tm.append(t2 - t1)
        if count is not None:
            pairs = count(o)
        del o
    b, m = np.median(tb), np.median(tm)
    print(f"{label}: pairs {pairs}  build {1e3 * b:.1f} ms  map {1e3 * m:.1f} ms  -> {pairs / (b + m):.3e} pair-evals/s (build + map), {pairs / m:.3e} (map)", flush=True)


w = W.c1_neighborlist()
timed("C1 neighbour list 10k F64", lambda: om.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"]), lambda o: len(o.neighborlist(nbatches=nt)[0]), reps=5)
w2 = W.c2_argon(50 if small else 100, np.float32)
timed(f"C2 LJ energy+forces {w2['x'].shape[0]} F32", lambda: om.Oracle(w2["x"], w2["cutoff"], unitcell=w2["unitcell"], dtype=np.float32),
      lambda o: o.lj(w2["c6"], w2["c12"], forces=True, nbatches=nt), count=lambda o: o.sum_d_d2(nbatches=nt)[2])
n3 = 100_000 if small else 1_000_000
w3 = W.c3_triclinic_cross(n3, n3)
timed(f"C3 triclinic cross {n3} x {n3} min-distance F64", lambda: om.Oracle(w3["x"], w3["cutoff"], unitcell=w3["unitcell"], y=w3["y"]),
      lambda o: o.mindist(nbatches=nt), reps=2, count=lambda o: o.sum_d_d2(nbatches=nt)[2])
for dim in (3, 2):
    n4 = 200_000 if small else 4_000_000
    w4 = W.c4_galaxies(n4, dim)
    timed(f"C4 pair velocities {n4} galaxies {dim}-D F64", lambda: om.Oracle(w4["x"], w4["cutoff"], unitcell=w4["unitcell"]),
          lambda o: int(o.pairvel(w4["v"], w4["rbins"], nbatches=nt)[0].sum()), reps=1 if dim == 3 and not small else 2)

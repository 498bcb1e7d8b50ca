import sys, numpy as np
sys.path.insert(0, "/root/repo")
import celllistmap_b200 as clm, workloads as W
w = W.c3_triclinic_cross(1_000_000, 1_000_000)
h = clm.Handle(3, np.float64)
h.set_box(clm._capi.TRICLINIC, w["unitcell"], w["cutoff"], 1)
i, j, d = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1)
for k in range(4):
    h.set_positions(0, w["x"]); h.set_positions(1, w["y"]); h.map_mindist(i, j, d, profile=True)
    st = h.stats(); print("C3 build %.3f ms sweep %.3f ms  (%d, %d, %.17g)" % (st.build_ms, st.sweep_ms, i[0], j[0], d[0]))

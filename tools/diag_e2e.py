import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import celllistmap_b200 as clm
import workloads as W
dtype=np.float32; tdt=torch.float32
w = W.c2_argon(100, dtype); n = w["x"].shape[0]
h = clm.Handle(3, dtype)
st = torch.cuda.Stream()
h.set_stream(st.cuda_stream)
h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
xs = [torch.from_numpy(w["x"]).pin_memory() for _ in range(2)]
fs = [torch.zeros((n, 3), dtype=tdt).pin_memory() for _ in range(2)]
es = [torch.zeros(1, dtype=tdt).pin_memory() for _ in range(2)]
for k in range(6):
    h.set_positions_async(0, xs[k & 1].numpy()); h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy(), async_=True)
h.synchronize()
h.set_option("dbg", int(os.environ.get("DBG", "0")))   # both coordinate buffers hold valid data from here on
steps=40
evs=[]; host=[]
t0=time.perf_counter()
for k in range(steps):
    a=time.perf_counter()
    h.set_positions_async(0, xs[k & 1].numpy())
    b=time.perf_counter()
    h.map_lj(w["c6"], w["c12"], es[k & 1].numpy(), fs[k & 1].numpy(), async_=True)
    c=time.perf_counter()
    e=torch.cuda.Event(enable_timing=True); e.record(st); evs.append(e)
    host.append((b-a, c-b))
h.synchronize()
tot=(time.perf_counter()-t0)/steps
d=[evs[i].elapsed_time(evs[i+1]) for i in range(steps-1)]
print("wall/step %.3f ms; gpu main-stream step (event deltas) median %.3f ms; host set_pos %.3f ms, host map %.3f ms" % (1e3*tot, np.median(d), 1e3*np.median([x[0] for x in host]), 1e3*np.median([x[1] for x in host])))
# device-resident for comparison
xd=torch.from_numpy(w["x"]).cuda(); fd=torch.zeros((n,3),dtype=tdt,device="cuda"); ed=torch.zeros(1,dtype=tdt,device="cuda")
for k in range(5):
    h.set_positions(0, xd); h.map_lj(w["c6"], w["c12"], ed, fd)
torch.cuda.synchronize(); t0=time.perf_counter()
for k in range(steps):
    h.set_positions(0, xd); h.map_lj(w["c6"], w["c12"], ed, fd)
torch.cuda.synchronize()
print("device-resident wall/step %.3f ms" % (1e3*(time.perf_counter()-t0)/steps))

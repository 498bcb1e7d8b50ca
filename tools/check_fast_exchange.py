#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_fast_exchange.py : the single-sync halo exchange (clm_select_layers + fixed-
capacity NCCL messages) must deliver exactly the rows of the exact, collective exchange, and the same LJ energy."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import celllistmap_b200  # noqa: F401
from celllistmap_b200 import slab
import bench_multi
import workloads as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x_host, unitcell = bench_multi.slab_lattice(rank, world, 60, 20, np.float32)   # 20 planes = 6 whole cell layers per rank
s = slab.SlabSystem(unitcell, 12.0, dtype=np.float32)
x = torch.from_numpy(x_host).cuda()
f = torch.zeros_like(x)
s.update(x)                       # exact path (sizes the fast path)
slow = s.x_foreign.clone()
e_slow = float(s.map_lj(W.ARGON_C6, W.ARGON_C12, f))
f_slow = f.clone()
assert s._cap is not None
s.update(x)                       # fast path
fast = s.x_foreign.clone()
e_fast = float(s.map_lj(W.ARGON_C6, W.ARGON_C12, f))
key = lambda t: sorted(map(tuple, t.cpu().numpy().tolist()))
ok = key(slow) == key(fast)
print(f"[rank {rank}] foreign rows slow {slow.shape[0]} fast {fast.shape[0]} identical-set {ok}  energy slow {e_slow:.6e} fast {e_fast:.6e}  max|df| {float((f - f_slow).abs().max()):.3e}", flush=True)
assert ok and abs(e_slow - e_fast) <= 1e-6 * abs(e_slow)
dist.barrier()
dist.destroy_process_group()

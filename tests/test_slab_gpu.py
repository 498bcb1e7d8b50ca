"""Slab-decomposed path on the GPU: 2 and 3 ranks sharing cuda:0 (gloo plumbing, halo staged through the host), results
against the CPU oracle / the single-handle run: neighbour lists bit-exact (union over ranks), counts exact, LJ forces
and energy within tolerance.  On a multi-GPU box bench.py exercises the same code over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import workloads as W

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        import celllistmap_b200  # noqa: F401
        from celllistmap_b200 import slab
        dtype = np.float64 if case == "list" else np.float32
        if case == "list":
            w = W.c1_neighborlist(6000)
        else:
            w = W.c2_argon(24, dtype)
        s = slab.SlabSystem(w["unitcell"], w["cutoff"], dtype=dtype)
        xo, ids = s.partition(w["x"])
        s.update(xo, ids)
        if case == "list":
            rec = s.neighborlist()
            sd, sd2, n = s.sum_d_d2()
            h = s.dist_hist(w["cutoff"] / 10, 10).cpu().numpy()
            q.put((rank, rec["i"].tolist(), rec["j"].tolist(), rec["d"].tolist(), n, h.tolist()))
        else:
            f = torch.zeros((s.n_owned, 3), dtype=torch.float32, device="cuda")
            e = s.map_lj(w["c6"], w["c12"], f)
            q.put((rank, ids.cpu().numpy().tolist(), f.cpu().numpy().tolist(), float(e), s.n_foreign))
        s.close()
    finally:
        dist.destroy_process_group()


def _run(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 3])
def test_slab_neighborlist_bit_exact(oracle_mod, world):
    res = _run(world, "list")
    w = W.c1_neighborlist(6000)
    o = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"])
    wi, wj, wd = o.neighborlist()
    gi = np.concatenate([np.array(r[1], np.int64) for r in res])
    gj = np.concatenate([np.array(r[2], np.int64) for r in res])
    gd = np.concatenate([np.array(r[3], np.float64) for r in res])
    key = lambda a, b, d: sorted(zip(np.minimum(a, b).tolist(), np.maximum(a, b).tolist(), d.tolist()))
    assert key(gi, gj, gd) == key(wi, wj, wd), "union of the per-rank lists must equal the reference list bit for bit"
    for r in res:
        assert r[4] == len(wi)                                   # all_reduced pair count
        assert r[5] == o.dist_hist(w["cutoff"] / 10, 10).tolist()  # all_reduced histogram


@pytest.mark.parametrize("world", [2, 3])
def test_slab_lj_forces(oracle_mod, world):
    res = _run(world, "lj")
    w = W.c2_argon(24, np.float32)
    we, wf = oracle_mod.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64)).lj(w["c6"], w["c12"], forces=True)
    f = np.zeros_like(wf)
    seen = np.zeros(len(wf), bool)
    for r in res:
        ids = np.array(r[1]) - 1
        assert not seen[ids].any()
        seen[ids] = True
        f[ids] = np.array(r[2])
        assert abs(r[3] - we) <= 2e-5 * abs(we)
        assert r[4] > 0
    assert seen.all()
    assert np.abs(f - wf).max() <= 4e-5 * np.abs(wf).max()

"""Host logic of the multi-GPU slab decomposition on CPU: world_size-2 and -3 process groups over gloo.

No GPU and no product compute here: the cell-layer index (which the product gets from clm_cell_coords) is injected
as a small numpy function; what is tested is ownership, face selection and the halo exchange plumbing of
celllistmap.jl_b200/slab.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layers(x, L, n_inner, lcell):
    cs = L / n_inner
    return (np.floor(np.mod(x[:, 0], L) / cs).astype(np.int64) % n_inner) + lcell


def _worker(rank, world, port, n_inner, lcell, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import celllistmap_b200  # noqa: F401  (import shim)
        from celllistmap_b200 import slab
        L = 10.0
        rng = np.random.default_rng(5)
        x = rng.random((3000, 3)) * L
        ids = np.arange(1, 3001)
        c = _layers(x, L, n_inner, lcell)
        plan = slab.SlabPlan(n_inner, lcell, world)
        mine = plan.owner_of(c) == rank
        xo, io, co = torch.from_numpy(x[mine]), torch.from_numpy(ids[mine]), torch.from_numpy(c[mine])
        lo_m, hi_m = plan.face_masks(co, rank)
        # a per-particle side array (weights, velocities, user inputs) travels with the halo like the coordinates
        aux = np.stack([np.sin(ids), np.cos(ids)], 1)
        got_x, got_i, got_a = slab.exchange_halo([xo, io, torch.from_numpy(aux[mine])], lo_m, hi_m, plan, rank)
        same = bool(torch.equal(got_x, torch.from_numpy(x[got_i.numpy() - 1]))) and bool(torch.equal(got_a, torch.from_numpy(aux[got_i.numpy() - 1])))
        q.put((rank, ids[mine].tolist(), got_i.tolist(), same, plan.bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_inner,lcell", [(2, 10, 1), (3, 12, 2), (2, 4, 2)])
def test_halo_exchange_gloo(world, n_inner, lcell):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_inner, lcell, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L = 10.0
    x = np.random.default_rng(5).random((3000, 3)) * L
    c = _layers(x, L, n_inner, lcell)
    owned_all = []
    for rank, owned, foreign, same_bits, bounds in res:
        assert same_bits, "halo coordinates and side arrays must arrive bit-identical, row for row"
        owned_all += owned
        lo, hi = bounds[rank], bounds[rank + 1]
        assert all(lo <= c[i - 1] < hi for i in owned)
        # every particle within lcell cell layers (periodically) of the slab, and not owned, must have been received
        rel = (c - lcell)
        need = set()
        for i in range(3000):
            if lo <= c[i] < hi:
                continue
            d_up = (rel[i] - (hi - lcell)) % n_inner          # layers above the slab top (periodic)
            d_dn = ((lo - lcell) - 1 - rel[i]) % n_inner       # layers below the slab bottom (periodic)
            if d_up < lcell or d_dn < lcell:
                need.add(i + 1)
        assert need == set(foreign), (rank, len(need), len(foreign))
        assert len(foreign) == len(set(foreign)), "duplicate halo particles"
    assert sorted(owned_all) == list(range(1, 3001)), "every particle is owned by exactly one rank"


def test_plan_rejects_thin_slabs():
    import celllistmap_b200  # noqa: F401
    from celllistmap_b200 import slab
    with pytest.raises(ValueError):
        slab.SlabPlan(5, 2, 4)
    p = slab.SlabPlan(30, 1, 8)
    assert p.bounds[0] == 1 and p.bounds[-1] == 31
    assert (np.diff(p.bounds) >= 3).all()


def _rows_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import celllistmap_b200  # noqa: F401
        from celllistmap_b200 import slab
        # rank r owns ids r+1, r+1+world, ...; row k goes to every other rank q with (id + q) % 3 == 0: arbitrary peers,
        # some rows to several ranks, some to none (the general point-to-point exchange of the triclinic slabs)
        ids = torch.arange(rank + 1, 601, world)
        x = torch.stack([ids.double(), ids.double() ** 2, -ids.double()], 1)
        sends = {qq: ((ids + qq) % 3 == 0) for qq in range(world) if qq != rank}
        sends = {qq: m for qq, m in sends.items() if bool(m.any())}
        gx, gi = slab.exchange_rows([x, ids], sends, world, rank)
        q.put((rank, gi.tolist(), bool(torch.equal(gx, torch.stack([gi.double(), gi.double() ** 2, -gi.double()], 1)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_rows_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, same in res:
        assert same, "rows must arrive bit-identical, every payload in the same order"
        want = sorted(i for i in range(1, 601) if (i - 1) % world != rank and (i + rank) % 3 == 0)
        assert sorted(got) == want and len(got) == len(set(got))

"""Slab-decomposed path on the GPU: 2 and 3 ranks sharing cuda:0 (gloo plumbing, halo staged through the host), results
against the CPU oracle / the single-handle run: neighbour lists bit-exact (union over ranks), counts exact, LJ forces
and energy within tolerance; every case also through the single-sync ("fast") halo exchange, and -- when the box has two
GPUs -- over NCCL, one rank per GPU (the path bench.py times at N > 1)."""
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


CUSTOM_SRC = """
struct WeightedCount {   // scalar: sum par0 w_i w_j; per particle: sum_j w_j
    static constexpr int NSCALAR = 1, NPART = 1, NAUX = 1, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        out.add_scalar(0, par[0] * p.ai[0] * p.aj[0]);
        out.add_i(0, p.aj[0]);
    }
};
"""


def _worker(rank, world, port, case, q, fast=False, backend="gloo"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    dev = rank if backend == "nccl" else 0
    assert dev < ndev
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import celllistmap_b200  # noqa: F401
        from celllistmap_b200 import slab
        dtype = np.float64 if case in ("list", "aux", "cross", "tri") else np.float32
        if case in ("tri", "tri32"):
            # triclinic self-set system: halo by periodic images, owned + halo rows in global-id order with a foreign mask
            dt = np.float64 if case == "tri" else np.float32
            w = W.triclinic_argon(18, dt)
            s = slab.SlabSystem(w["unitcell"], w["cutoff"], dtype=dt)
            xo, ids = s.partition(w["x"])
            s.update(xo, ids)
            rec = s.neighborlist()
            sd, sd2, n = s.sum_d_d2()
            mi, mj, md = s.mindist()
            f = torch.zeros((s.n_owned, 3), dtype=torch.float64 if case == "tri" else torch.float32, device="cuda")
            e = s.map_lj(w["c6"], w["c12"], f)
            q.put((rank, rec["i"].tolist(), rec["j"].tolist(), rec["d"].tolist(), n, (mi, mj, md), ids.cpu().numpy().tolist(),
                   f.cpu().numpy().tolist(), float(e), s.n_foreign, s.n_owned))
            s.close()
            return
        if case == "cross":
            # two-set system: x particles sharded by slab, the y set (partners only) with its halo
            rng = np.random.default_rng(21)
            x, y = rng.random((4000, 3)), rng.random((3000, 3))
            s2 = slab.SlabSystem2(np.ones(3), 0.1, dtype=np.float64)
            xo, xi = s2.partition(x)
            yo, yi = s2.partition(y)
            s2.update(xo, yo, xi, yi)
            rec = s2.neighborlist()
            mi, mj, md = s2.mindist()
            sd, sd2, n = s2.sum_d_d2()
            q.put((rank, rec["i"].tolist(), rec["j"].tolist(), rec["d"].tolist(), n, (mi, mj, md), s2.n_y_halo))
            s2.close()
            return
        if case == "aux":
            # side arrays travel with the halo: Coulomb weights, pair velocities, a user pair function
            w = W.c1_neighborlist(5000)
            rng = np.random.default_rng(9)
            wts, vel = 0.5 + rng.random(5000), rng.random((5000, 3))
            s = slab.SlabSystem(w["unitcell"], w["cutoff"], dtype=dtype)
            xo, ids = s.partition(w["x"])
            own = ids.cpu().numpy() - 1
            s.update(xo, ids, aux=wts[own])
            if fast:      # second update through the single-sync exchange (fixed-capacity messages sized by the first one)
                s.force_fast = True
                s.update(xo, ids, aux=wts[own])
                assert s.fast_exchanges == 1
            f = torch.zeros((s.n_owned, 3), dtype=torch.float64, device="cuda")
            e = s.map_coulomb(-9.8, f)
            sc, pp, _, _ = s.map_custom(CUSTOM_SRC, "WeightedCount", params=(2.0,))
            mi, mj, md = s.mindist()
            s.update(xo, ids, aux=vel[own])
            counts, sums = s.pairvel(np.array([0.0, 0.02, 0.04, 0.06, 0.08, 0.1]))
            q.put((rank, own.tolist(), f.cpu().numpy().tolist(), float(e), sc.cpu().numpy().tolist(), pp.cpu().numpy().tolist(),
                   (mi, mj, md), counts.cpu().numpy().tolist(), sums.cpu().numpy().tolist()))
            s.close()
            return
        if case == "list":
            w = W.c1_neighborlist(6000)
        else:
            w = W.c2_argon(24, dtype)
        s = slab.SlabSystem(w["unitcell"], w["cutoff"], dtype=dtype)
        xo, ids = s.partition(w["x"])
        s.update(xo, ids)
        if fast:
            s.force_fast = True
            s.update(xo, ids)
            assert s.fast_exchanges == 1
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


def _run(world, case, fast=False, backend="gloo"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q, fast, backend)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


MODES = [(2, False, "gloo"), (3, False, "gloo"), (2, True, "gloo"), (3, True, "gloo"), (2, True, "nccl")]


def _skip_unless_possible(world, backend):
    import torch
    if backend == "nccl" and torch.cuda.device_count() < world:
        pytest.skip("the NCCL variant needs one GPU per rank (run with gpurun --gpus 2; output kept under profiles/)")


@pytest.mark.parametrize("world,fast,backend", MODES)
def test_slab_neighborlist_bit_exact(oracle_mod, world, fast, backend):
    _skip_unless_possible(world, backend)
    res = _run(world, "list", fast, backend)
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


@pytest.mark.parametrize("world,fast,backend", MODES)
def test_slab_lj_forces(oracle_mod, world, fast, backend):
    _skip_unless_possible(world, backend)
    res = _run(world, "lj", fast, backend)
    w = W.c2_argon(24, np.float32)
    we, wf = oracle_mod.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64)).lj(w["c6"], w["c12"], forces=True)
    f = np.zeros_like(wf)
    seen = np.zeros(len(wf), bool)
    for r in res:
        ids = np.array(r[1]) - 1
        assert not seen[ids].any()
        seen[ids] = True
        f[ids] = np.array(r[2])
        assert abs(r[3] - we) <= 1e-5 * abs(we)
        assert r[4] > 0
    assert seen.all()
    # the bar holds against the oracle in the same precision (bit-identical wrapped coordinates on every rank); the
    # distance to Float64 arithmetic is reported (conditioning of Float32 coordinates, tests/parity_util.py)
    from parity_util import force_report
    f32 = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"], dtype=np.float32).lj(w["c6"], w["c12"], forces=True)[1]
    err_same, _, _ = force_report(f"slab LJ {world} ranks f32", f, f32, wf)
    assert err_same <= 1e-5


@pytest.mark.parametrize("world,fast,backend", MODES)
def test_slab_side_arrays(oracle_mod, world, fast, backend):
    """Coulomb energy + forces, a user pair function, the minimum distance and the pair-velocity histogram of a
    slab-decomposed system equal the single-process results (side arrays exchanged with the halo).  fast: through the
    single-sync exchange (clm_select_layers + fixed-capacity messages + collective overflow flag), which carries the global
    ids and the side arrays too; backend nccl: the path the multi-GPU bench times."""
    _skip_unless_possible(world, backend)
    res = _run(world, "aux", fast, backend)
    w = W.c1_neighborlist(5000)
    rng = np.random.default_rng(9)
    wts, vel = 0.5 + rng.random(5000), rng.random((5000, 3))
    o = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"])
    we, wf = o.coulomb(-9.8, wts, forces=True)
    i, j, d = o.neighborlist()
    i, j = i - 1, j - 1
    want_sc = 2.0 * (wts[i] * wts[j]).sum()
    want_pp = np.zeros(5000)
    np.add.at(want_pp, i, wts[j])
    np.add.at(want_pp, j, wts[i])
    rbins = np.array([0.0, 0.02, 0.04, 0.06, 0.08, 0.1])
    wc, ws = o.pairvel(vel, rbins)
    k = np.argmin(d)
    f, pp = np.zeros((5000, 3)), np.zeros(5000)
    for r in res:
        own = np.array(r[1])
        f[own] = np.array(r[2])
        pp[own] = np.array(r[5])[:, 0]
        assert abs(r[3] - we) <= 1e-10 * abs(we)
        assert abs(r[4][0] - want_sc) <= 1e-10 * want_sc
        mi, mj, md = r[6]
        assert md == d[k] and {mi, mj} == {int(i[k]) + 1, int(j[k]) + 1}
        assert r[7] == wc.tolist()
        assert np.abs(np.array(r[8]) - ws).max() <= 1e-10 * np.abs(ws).max()
    assert np.abs(f - wf).max() <= 1e-9 * np.abs(wf).max()
    assert np.abs(pp - want_pp).max() <= 1e-10 * want_pp.max()


# ---------------------------------------------------------------------------------------------------------
# the same decomposition driven by the library itself over NCCL (clm_comm_init / clm_slab_update / clm_comm_allreduce_sum):
# what a Julia or C host without its own communication layer calls.  One GPU per rank; world = 1 runs on any box.
def _capi_worker(rank, world, uid, q):
    import torch
    torch.cuda.set_device(rank)
    import celllistmap_b200 as clm
    w = W.c2_argon(24, np.float32)
    h = clm.Handle(3, np.float32, device=rank)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    h.comm_init(uid, rank, world)
    lo, hi = h.slab_range()
    c = h.cell_coords(w["x"], 0)
    c = c.cpu().numpy() if hasattr(c, "cpu") else np.asarray(c)
    top = hi + (1 if rank == world - 1 else 0)      # the last rank also owns the rounding layer at the top
    mine = np.nonzero((c >= lo) & (c < top))[0]
    for step in range(2):                           # second step: warm halo capacity
        h.slab_update(w["x"][mine])
        e, f = np.zeros(1, np.float32), np.zeros((len(mine), 3), np.float32)
        h.map_lj(w["c6"], w["c12"], e, f)
        sd, sd2, npairs = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.int64)
        h.map_sum_d_d2(sd, sd2, npairs)
        e64 = e.astype(np.float64)
        h.comm_allreduce_sum(e64)
        h.comm_allreduce_sum(npairs)
    n_own, n_for, r, wd = h.slab_info()
    assert (r, wd, n_own) == (rank, world, len(mine))
    q.put((rank, mine.tolist(), f.tolist(), float(e64[0]), int(npairs[0]), n_for))
    h.comm_destroy()
    h.close()


@pytest.mark.parametrize("world", [1, 2])
def test_slab_through_the_c_abi_over_nccl(oracle_mod, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("one GPU per rank (run with gpurun --gpus 2; output kept under profiles/)")
    import celllistmap_b200 as clm
    uid = clm._capi.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_capi_worker, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=150) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    w = W.c2_argon(24, np.float32)
    o64 = oracle_mod.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64))
    we, wf = o64.lj(w["c6"], w["c12"], forces=True)
    wn = o64.sum_d_d2()[2]
    f32 = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"], dtype=np.float32).lj(w["c6"], w["c12"], forces=True)[1]
    f = np.zeros_like(wf)
    seen = np.zeros(len(wf), bool)
    for r in res:
        ids = np.array(r[1])
        assert not seen[ids].any()
        seen[ids] = True
        f[ids] = np.array(r[2])
        assert abs(r[3] - we) <= 1e-5 * abs(we)
        assert r[4] == wn
        assert (r[5] > 0) == (world > 1)
    assert seen.all()
    from parity_util import force_report
    err_same, _, _ = force_report(f"C-ABI slab LJ {world} rank(s) f32", f, f32, wf)
    assert err_same <= 1e-5


@pytest.mark.parametrize("world", [2, 3])
def test_slab_two_set_system(oracle_mod, world):
    """cross-set slabs (src/internals/cross.jl:8-25): the union of the per-rank cross lists is the reference list bit for
    bit, the global minimum distance and pair count are the single-process ones."""
    res = _run(world, "cross")
    rng = np.random.default_rng(21)
    x, y = rng.random((4000, 3)), rng.random((3000, 3))
    o = oracle_mod.Oracle(x, 0.1, unitcell=np.ones(3), y=y)
    wi, wj, wd = o.neighborlist()
    gi = np.concatenate([np.array(r[1], np.int64) for r in res])
    gj = np.concatenate([np.array(r[2], np.int64) for r in res])
    gd = np.concatenate([np.array(r[3], np.float64) for r in res])
    assert sorted(zip(gi.tolist(), gj.tolist(), gd.tolist())) == sorted(zip(wi.tolist(), wj.tolist(), wd.tolist()))
    k = int(np.argmin(wd))
    for r in res:
        assert r[4] == len(wi)
        assert r[5][2] == wd[k] and (r[5][0], r[5][1]) == (int(wi[k]), int(wj[k]))
        assert r[6] > 0


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["tri", "tri32"])
def test_slab_triclinic(oracle_mod, world, case):
    """triclinic slabs (src/internals/self.jl:164-184: i real, index_i < index_j): owned + halo particles reach the engine in
    GLOBAL-id order with the halo rows flagged (clm_set_foreign_mask), so every pair is evaluated by the owner of its
    smaller-index particle from the same periodic image as on one GPU: the union of the per-rank lists equals the
    reference list bit for bit (distances included), forces and energy within tolerance."""
    dt = np.float64 if case == "tri" else np.float32
    res = _run(world, case)
    w = W.triclinic_argon(18, dt)
    o = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"], dtype=dt)
    wi, wj, wd = o.neighborlist()
    gi = np.concatenate([np.array(r[1], np.int64) for r in res])
    gj = np.concatenate([np.array(r[2], np.int64) for r in res])
    gd = np.concatenate([np.array(r[3], dt) for r in res])
    key = lambda a, b, d: sorted(zip(np.minimum(a, b).tolist(), np.maximum(a, b).tolist(), d.tolist()))
    assert key(gi, gj, gd) == key(wi, wj, wd), "union of the per-rank lists must equal the reference list bit for bit"
    k = int(np.argmin(wd))
    o64 = oracle_mod.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64))
    we, wf = o64.lj(w["c6"], w["c12"], forces=True)
    f = np.zeros_like(wf)
    seen = np.zeros(len(wf), bool)
    n_halo = 0
    esame, fsame = o.lj(w["c6"], w["c12"], forces=True)     # the oracle in the precision of the ranks
    # Float32: the slab path sweeps the full shell, so f_i and f_j of a pair come from two different periodic images whose
    # Float32 coordinates round differently -- the error against the same-precision oracle is bounded by the conditioning of
    # Float32 coordinates (the oracle's own Float32-vs-Float64 distance), not by 1e-5; both are reported and asserted
    e_cond = abs(esame - we) / abs(we)
    e_tol = 1e-10 if dt == np.float64 else max(1e-5, e_cond)
    for r in res:
        assert r[4] == len(wi)
        assert r[5][2] == wd[k] and {r[5][0], r[5][1]} == {int(wi[k]), int(wj[k])}
        ids = np.array(r[6]) - 1
        assert not seen[ids].any()
        seen[ids] = True
        f[ids] = np.array(r[7])
        print(f"[parity] triclinic slab {case} {world} ranks: |E - E_oracle64| / |E| = {abs(r[8] - we) / abs(we):.3e} (oracle in the same precision: {e_cond:.3e})")
        assert abs(r[8] - we) <= e_tol * abs(we)
        n_halo += r[9]
        print(f"[slab] triclinic rank {r[0]}: {r[10]} owned, {r[9]} halo particles of {len(wf)}")   # (a box of a few cells per side: most of it is halo)
    assert seen.all() and n_halo > 0
    from parity_util import force_report
    err_same, err_64, ref_64 = force_report(f"triclinic slab LJ {world} ranks {case}", f, fsame, wf)
    assert err_same <= (1e-10 if dt == np.float64 else max(1e-5, 2.0 * ref_64))

"""Oracle cell-list path vs its independent O(N^2) twin of the reference's map_naive!
(test/modules/Testing.jl:76-104), re-creating the reference's property tests:
check_random_cells (Testing.jl:162-237), pathological_coordinates (:36-60),
test_pathological (:563-599), margin/lcell tests (test/internals/tests.jl:595-684).
"""
import numpy as np
import pytest

from golden import kats as K


def pair_sets_equal(om, a, b, cutoff, band=1e-11):
    """Pair sets must be identical except for pairs whose distance is within `band` of the cutoff
    (the two paths wrap coordinates differently, so d differs in the last bits)."""
    pa, da = om.sorted_pairs(*a)
    pb, db = om.sorted_pairs(*b)
    ka = {tuple(p) for p, d in zip(pa.tolist(), da) if abs(d - cutoff) > band * cutoff}
    kb = {tuple(p) for p, d in zip(pb.tolist(), db) if abs(d - cutoff) > band * cutoff}
    assert len(pa) == len(np.unique(pa, axis=0)), "duplicate pairs in cell-list result"
    return ka == kb


def random_cell(rng, N, triclinic):
    M = np.zeros((N, N))
    if triclinic:
        M[:] = -10 + 20 * rng.random((N, N))
    else:
        M[np.diag_indices(N)] = -10 + 20 * rng.random(N)
    return M


@pytest.mark.parametrize("N", [2, 3])
@pytest.mark.parametrize("triclinic", [False, True])
@pytest.mark.parametrize("lcell", [1, 2, 3])
def test_random_cells(oracle_mod, N, triclinic, lcell):
    om = oracle_mod
    rng = np.random.default_rng(1000 * N + 10 * lcell + int(triclinic))
    ntrial = 0
    attempts = 0
    while ntrial < 25 and attempts < 20000:
        attempts += 1
        M = random_cell(rng, N, triclinic)
        cutoff = 1 + rng.random()
        M_arg = M if triclinic else list(np.diag(M))
        x = 10 * rng.random((rng.integers(10, 21), N)) - 50
        try:
            o = om.Oracle(x, cutoff, unitcell=M_arg, lcell=lcell)
        except om.OracleError as e:
            assert e.code == 2  # unit cell check failed -> the reference `continue`s
            continue
        if np.prod(o.box()["nc"]) > 100000:
            continue
        cl = o.neighborlist()
        if len(cl[0]) == 0:
            continue
        ntrial += 1
        assert pair_sets_equal(om, cl, o.neighborlist(algo=om.ALGO_NAIVE), cutoff), (M, cutoff, x)
        assert pair_sets_equal(om, cl, o.neighborlist(algo=om.ALGO_CELLLIST_NOPROJ), cutoff, band=0.0)
        assert pair_sets_equal(om, cl, o.neighborlist(nbatches=3), cutoff, band=0.0)
        s_cl = o.sum_d_d2()
        s_nv = o.sum_d_d2(algo=om.ALGO_NAIVE)
        if s_cl[2] == s_nv[2]:
            assert abs(s_cl[1] - s_nv[1]) <= 1e-9 * abs(s_nv[1])
    assert ntrial == 25


def pathological_coordinates(rng, n):
    sides = np.array([250.0, 250.0, 250.0])
    x = sides * rng.random((n, 3))
    nf, pf = np.nextafter(0.0, 1.0), np.nextafter(0.0, -1.0)
    r = lambda: sides[2] * rng.random()
    x[0] = -sides / 2
    x[1] = -sides / 2 + [nf, nf, r()]
    x[2] = -sides / 2 + [pf, pf, r()]
    x[3] = sides / 2 + [nf, nf, r()]
    x[4] = sides / 2 + [pf, pf, r()]
    x[9] = sides
    x[10] = sides + [nf, nf, r()]
    x[11] = sides + [pf, pf, r()]
    x[12] = [nf, nf, r()]
    x[13] = [pf, pf, r()]
    x[14] = 0.0
    x[99] = [sides[0] / 2, -sides[1] / 2, 2 * sides[2]]
    y = sides * rng.random((n, 3))
    return x, y, sides, 10.0


@pytest.mark.parametrize("lcell", [1, 2])
def test_pathological_coordinates_self_and_cross(oracle_mod, lcell):
    om = oracle_mod
    rng = np.random.default_rng(321)
    x, y, sides, cutoff = pathological_coordinates(rng, 1500)
    o = om.Oracle(x, cutoff, unitcell=list(sides), lcell=lcell)
    assert pair_sets_equal(om, o.neighborlist(), o.neighborlist(algo=om.ALGO_NAIVE), cutoff)
    o2 = om.Oracle(x[:400], cutoff, unitcell=list(sides), y=y, lcell=lcell)
    a = o2.neighborlist()
    b = o2.neighborlist(algo=om.ALGO_NAIVE)
    pa, da = om.sorted_pairs(*a, ordered=True)
    pb, db = om.sorted_pairs(*b, ordered=True)
    assert np.array_equal(pa, pb)
    np.testing.assert_allclose(da, db, rtol=1e-9)
    # triclinic box with the same (orthorhombic) matrix must give the same pairs (tests.jl:121-275 pattern)
    o3 = om.Oracle(x, cutoff, unitcell=np.diag(sides), lcell=lcell)
    assert pair_sets_equal(om, o.neighborlist(), o3.neighborlist(), cutoff)


def test_pathological_2d_matrices(oracle_mod):
    om = oracle_mod
    rng = np.random.default_rng(7)
    mats = list(K.PATHOLOGICAL_2D_CELLS) + [np.array([[-1.2, 0.2], [0.2, 1.2]]), np.array([[-1.2, 0.2], [0.2, -1.2]])]
    for M in mats:
        for lcell in range(1, 6):
            for _ in range(20):
                x = rng.random((2, 2))
                o = om.Oracle(x, 0.2, unitcell=M, lcell=lcell)
                assert pair_sets_equal(om, o.neighborlist(), o.neighborlist(algo=om.ALGO_NAIVE), 0.2), (M, lcell, x)
    # lattice-like point sets of test/internals/tests.jl:465-470 (distances exactly at 0.2 are excluded by the band)
    for M in K.PATHOLOGICAL_2D_CELLS:
        for x in (100 * rng.random((100, 2)),
                  np.array([[0.1 * i + 0.1 * j, 0.2 * j] for i in range(6) for j in range(6)]),
                  np.array([[0.1 * i, 0.1 * j] for i in range(6) for j in range(6)])):
            o = om.Oracle(x, 0.2, unitcell=M)
            assert pair_sets_equal(om, o.neighborlist(), o.neighborlist(algo=om.ALGO_NAIVE), 0.2, band=1e-9)


def test_triclinic_exactly_once_and_margins(oracle_mod):
    om = oracle_mod
    t = K.TRICLINIC_ONCE
    o = om.Oracle(np.array(t["x"]), t["cutoff"], unitcell=t["unitcell"])
    assert o.sum_d_d2()[2] == o.sum_d_d2(algo=om.ALGO_NAIVE)[2]
    rng = np.random.default_rng(11)
    for lcell in (1, 2, 3, 5):  # tests.jl:604-639
        sides = np.array([20.0, 20.0, 20.0])
        x = sides * rng.random((200, 3))
        cs = sides / (2 * lcell + 1)
        extra = [[cs[0] * i + 0.1, cs[1] * j + 0.1, cs[2] * k + 0.1] for i in range(1, lcell + 2) for j in range(1, lcell + 2) for k in range(1, lcell + 2)]
        x = np.vstack([x, np.array(extra)])
        o = om.Oracle(x, 2.5, unitcell=list(sides), lcell=lcell)
        a, b = o.sum_d_d2(), o.sum_d_d2(algo=om.ALGO_NAIVE)
        assert a[2] == b[2] and abs(a[0] - b[0]) <= 1e-10 * b[0]
    for lcell in (1, 2, 3):  # tests.jl:641-658
        uc = np.array([[15.0, 5.0, 0.0], [0.0, 15.0, 3.0], [0.0, 0.0, 15.0]])
        x = (uc @ rng.random((3, 150))).T
        o = om.Oracle(x, 2.0, unitcell=uc, lcell=lcell)
        a, b = o.sum_d_d2(), o.sum_d_d2(algo=om.ALGO_NAIVE)
        assert a[2] == b[2] and abs(a[0] - b[0]) <= 1e-10 * b[0]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_catalogue_celllist_vs_naive(oracle_mod, dtype):
    """Every catalogue functor: cell-list path (serial and batched) vs the naive twin."""
    om = oracle_mod
    rng = np.random.default_rng(5)
    n, L, rc = 600, 12.0, 2.0
    # jittered lattice avoids r -> 0 blow-ups of LJ
    g = np.stack(np.meshgrid(*[np.arange(9)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] * (L / 9)
    x = (g + 0.3 * rng.random((n, 3))).astype(dtype)
    tol = 2e-4 if dtype == np.float32 else 1e-10
    uc = np.array([[L, 0.0, 0.0], [1.5, L, 0.0], [0.7, 1.1, L]]).T
    for unitcell in ([L, L, L], uc, None):
        o = om.Oracle(x, rc, unitcell=unitcell, dtype=dtype)
        for nb in (0, 4):
            e, f = o.lj(1.0, 1.0, forces=True, nbatches=nb)
            en, fn = o.lj(1.0, 1.0, forces=True, algo=om.ALGO_NAIVE)
            assert abs(e - en) <= tol * abs(en)
            assert np.abs(f - fn).max() <= tol * np.abs(fn).max()
            w = (1 + rng.random(n)).astype(dtype)
            e, f = o.coulomb(-9.8, w, forces=True, nbatches=nb)
            en, fn = o.coulomb(-9.8, w, forces=True, algo=om.ALGO_NAIVE)
            assert abs(e - en) <= tol * abs(en)
            assert np.abs(f - fn).max() <= tol * np.abs(fn).max()
            assert np.array_equal(o.dist_hist(0.25, 8, nbatches=nb), o.dist_hist(0.25, 8, algo=om.ALGO_NAIVE))
            v = rng.random((n, 3)).astype(dtype)
            c, s = o.pairvel(v, [0.0, 0.5, 1.0, 1.5, 2.0], nbatches=nb)
            cn, sn = o.pairvel(v, [0.0, 0.5, 1.0, 1.5, 2.0], algo=om.ALGO_NAIVE)
            assert np.array_equal(c, cn)
            assert np.abs(s - sn).max() <= 50 * tol * np.abs(sn).max() + 1e-3 * (dtype == np.float32)
            i, j, d = o.mindist(nbatches=nb)
            i2, j2, d2 = o.mindist(algo=om.ALGO_NAIVE)
            assert {i, j} == {i2, j2} and abs(d - d2) <= tol * d2

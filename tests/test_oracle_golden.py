"""Pins the CPU oracle (oracle/) against the reference's own golden vectors and KATs.

These tests are the 'parity pin' of SURVEY.md §8(c): every value is transcribed from the
reference's tests/doctests/docs (tests/golden/kats.py cites file:line) and the inputs are the
reference's own fixtures converted by tests/golden/make_golden.py.
"""
import os

import numpy as np
import pytest

from golden import kats as K

HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = np.load(os.path.join(HERE, "golden", "namd_frames.npz"))
ARGON = np.load(os.path.join(HERE, "golden", "argon_cubic.npy"))


@pytest.mark.parametrize("frame", list(K.NAMD_CASES))
@pytest.mark.parametrize("lcell", K.NAMD_LCELLS)
def test_namd_lj_energy(oracle_mod, frame, lcell):
    uc, gold = K.NAMD_CASES[frame]
    x = FRAMES[frame].astype(np.float64)  # Chemfiles hands Float64 copies of the float32 DCD data
    o = oracle_mod.Oracle(x, K.NAMD_CUTOFF, unitcell=uc, lcell=lcell)
    for kw in (dict(), dict(nbatches=4), dict(algo=oracle_mod.ALGO_CELLLIST_NOPROJ)):
        e = o.lj(K.NAMD_C6, K.NAMD_C12, **kw)
        assert abs(e - gold) <= 1e-11 * abs(gold), (frame, lcell, kw, e, gold)  # reference test uses isapprox (rtol 1.5e-8)


def test_argon_sum_d2(oracle_mod):
    o = oracle_mod.Oracle(ARGON, K.ARGON_CUTOFF, unitcell=K.ARGON_UNITCELL)
    sd, sd2, n = o.sum_d_d2()
    assert n == K.ARGON_NL_PERIODIC[0]
    assert abs(sd2 - K.ARGON_SUM_D2) <= 1e-12 * K.ARGON_SUM_D2
    o2 = oracle_mod.Oracle(ARGON[:50], K.ARGON_CUTOFF, unitcell=K.ARGON_UNITCELL, y=ARGON[50:])
    _, sd2c, _ = o2.sum_d_d2()
    assert abs(sd2c - K.ARGON_SUM_D2_CROSS) <= 1e-12 * K.ARGON_SUM_D2_CROSS


def _first_of(i, j, d, pair):
    m = ((i == pair[0]) & (j == pair[1])) | ((i == pair[1]) & (j == pair[0]))
    assert m.sum() == 1
    return d[m][0]


def test_argon_neighborlists(oracle_mod):
    # These doctests were generated with a PDBTools that returns Float32 coordinates (promoted to
    # Float64 by the Float64 cutoff): with float32-rounded inputs the oracle reproduces the three
    # printed distances BIT-EXACTLY.
    A32 = ARGON.astype(np.float32).astype(np.float64)
    n, first = K.ARGON_NL_NONPERIODIC
    i, j, d = oracle_mod.Oracle(A32, K.ARGON_CUTOFF).neighborlist()
    assert len(i) == n and _first_of(i, j, d, first[:2]) == first[2]
    n, first = K.ARGON_NL_PERIODIC
    i, j, d = oracle_mod.Oracle(A32, K.ARGON_CUTOFF, unitcell=K.ARGON_UNITCELL).neighborlist()
    assert len(i) == n and _first_of(i, j, d, first[:2]) == first[2]
    n, first = K.ARGON_NL_CROSS
    i, j, d = oracle_mod.Oracle(A32[:50], K.ARGON_CUTOFF, y=A32[50:]).neighborlist()
    assert len(i) == n
    m = (i == first[0]) & (j == first[1])
    assert m.sum() == 1 and d[m][0] == first[2]


def test_argon_min_distance_and_inverse_distance_forces(oracle_mod):
    o = oracle_mod.Oracle(ARGON, K.ARGON_CUTOFF, unitcell=K.ARGON_UNITCELL)
    _, _, d = o.mindist()
    assert d == K.ARGON_MIN_DIST  # bit-exact with Float64-parsed coordinates
    # energy sum 1/d and forces f_i = sum (x_j - x_i)/d^3  == Coulomb functor with k = -1, unit weights
    w = np.ones(len(ARGON))
    e, f = o.coulomb(-1.0, w, forces=True)
    assert abs(-e - K.ARGON_SUM_INV_D) <= 1e-12 * K.ARGON_SUM_INV_D
    for idx, gold in K.ARGON_FORCES_INV_D.items():
        np.testing.assert_allclose(f[idx], gold, rtol=1e-11, atol=1e-15)


@pytest.mark.parametrize("kat", K.GRID_KATS, ids=lambda k: str(k[3]))
def test_grid_kats(oracle_mod, kat):
    uc, cutoff, lcell, nc, cs = kat
    dim = len(nc)
    o = oracle_mod.Oracle(np.zeros((1, dim)), cutoff, unitcell=uc, lcell=lcell)
    b = o.box()
    assert list(b["nc"]) == nc
    np.testing.assert_allclose(b["cell_size"], cs, rtol=1e-14)


def test_show_celllist_and_computing_box_kats(oracle_mod):
    k = K.SHOW_CELLLIST_KAT
    x = np.random.default_rng(1).random((k["n"], 3))
    st = oracle_mod.Oracle(x, k["cutoff"], unitcell=k["unitcell"]).stats()
    assert st["n_cells_real"] == k["n_cells_real"] and st["n_particles"] == k["n_particles"]
    c = K.COMPUTING_BOX_KAT
    b = oracle_mod.Oracle(np.zeros((1, 3)), c["cutoff"], unitcell=c["unitcell"]).box()
    np.testing.assert_allclose(b["cb_min"], c["lo"], rtol=1e-15)
    np.testing.assert_allclose(b["cb_max"], c["hi"], rtol=1e-15)
    # non-periodic unit cell = extent + 2.1 cutoff (test/API/ParticleSystem.jl:145-147)
    x = np.random.default_rng(2).random((50, 3))
    b = oracle_mod.Oracle(x, 0.1).box()
    np.testing.assert_allclose(np.diag(b["input_unit_cell"]), x.max(0) - x.min(0) + 2.1 * 0.1, rtol=1e-15)


@pytest.mark.parametrize("kat", K.BOUNDARY_KATS, ids=lambda k: f"rc{k[1]:.3g}-n{len(k[0])}-{k[3]}")
def test_boundary_kats(oracle_mod, kat):
    x, cutoff, uc, npairs = kat
    x = np.array(x, dtype=np.float64)
    for algo in (oracle_mod.ALGO_CELLLIST, oracle_mod.ALGO_CELLLIST_NOPROJ):
        i, j, d = oracle_mod.Oracle(x, cutoff, unitcell=uc).neighborlist(algo=algo)
        assert len(i) == npairs, (x, cutoff, uc, algo)
    if npairs == 1 and cutoff < 1.0 and uc == [2.0, 2.0]:
        assert d[0] == K.BOUNDARY_D_KAT


def test_bug84_unique_and_few_particles(oracle_mod):
    for pts in (K.BUG84_3D, K.BUG84_2D):
        i, j, d = oracle_mod.Oracle(np.array(pts, dtype=np.float32), 7.0, dtype=np.float32).neighborlist()
        sp = oracle_mod.sorted_pairs(i, j)
        assert len(np.unique(sp, axis=0)) == len(sp) and np.all(sp[:, 0] != sp[:, 1])
    x, y, rc, exp = K.FEW_CROSS
    i, j, d = oracle_mod.Oracle(np.array(x), rc, y=np.array(y)).neighborlist()
    assert (i[0], j[0]) == exp[:2] and abs(d[0] - exp[2]) < 1e-12
    z, rc, exp = K.FEW_SELF
    i, j, d = oracle_mod.Oracle(np.array(z), rc).neighborlist()
    assert len(i) == 1 and {int(i[0]), int(j[0])} == set(exp[:2]) and abs(d[0] - exp[2]) < 1e-12
    # empty systems (neighborlists.jl:257-260)
    i, j, d = oracle_mod.Oracle(np.zeros((0, 3)), 0.1, unitcell=[1.0, 1.0, 1.0]).neighborlist()
    assert len(i) == 0


def test_nan_is_rejected(oracle_mod):
    x = np.random.default_rng(0).random((10, 3))
    x[6, 1] = np.nan
    with pytest.raises(oracle_mod.OracleError) as ei:  # "Invalid coordinates" + 1-based index (CellOperations.jl:9-17)
        oracle_mod.Oracle(x, 0.1, unitcell=[1.0, 1.0, 1.0])
    assert "Invalid coordinates" in str(ei.value) and "index 7" in str(ei.value)
    with pytest.raises(oracle_mod.OracleError):  # unit cell check (Box.jl:243)
        oracle_mod.Oracle(x[:5], 0.6, unitcell=[1.0, 1.0, 1.0])


@pytest.mark.parametrize("kat", K.CELL_LIMITS_KATS, ids=lambda k: str(k[0]))
def test_align_cell_and_cell_limits_kats(oracle_mod, kat):
    """align_cell + cell_limits of the reference (src/internals/CellOperations.jl:337-479) through the oracle's Box:
    a triclinic box's computing box is (lo - lcell*cs, hi + lcell*cs) with cs = cutoff/lcell and origin = 0
    (Box.jl:252), so the limits of the aligned cell are recovered from it."""
    m, lo, hi, exact = kat
    m = np.array(m, dtype=np.float64)
    n = m.shape[0]
    cutoff = 0.05
    o = oracle_mod.Oracle(np.zeros((1, n)), cutoff, unitcell=m, triclinic=True)
    b = o.box()
    got_lo, got_hi = b["cb_min"] + cutoff, b["cb_max"] - cutoff
    if exact:
        assert np.allclose(got_lo, lo, rtol=0, atol=1e-15) and np.allclose(got_hi, hi, rtol=0, atol=1e-15)
    else:
        assert np.allclose(got_lo, lo, rtol=1e-12, atol=1e-10) and np.allclose(got_hi, hi, rtol=1e-12, atol=1e-10)
    # the aligned cell keeps the lattice: first (longest) vector along +x, same volume
    a = b["aligned_unit_cell"]
    assert abs(abs(np.linalg.det(a)) - abs(np.linalg.det(m))) <= 1e-12 * abs(np.linalg.det(m))
    k = int(np.argmax(np.linalg.norm(m, axis=0)))
    assert np.allclose(a[:, k][1:], 0.0, atol=1e-12) and a[0, k] > 0


@pytest.mark.parametrize("kat", K.ALIGN_CELL_KATS, ids=["l0l1", "-l0l1"])
def test_align_cell_2d_kats(oracle_mod, kat):
    """align_cell in 2-D (src/internals/CellOperations.jl:353-375): aligned matrix and rotation"""
    m, want_a, want_r = (np.array(v, dtype=np.float64) for v in kat)
    b = oracle_mod.Oracle(np.zeros((1, 2)), 0.05, unitcell=m, triclinic=True).box()
    assert np.allclose(b["aligned_unit_cell"], want_a, atol=1e-14)
    assert np.allclose(b["rotation"], want_r, atol=1e-14)
    assert np.allclose(b["rotation"] @ b["inv_rotation"], np.eye(2), atol=1e-14)


def test_align_cell_3d_random_rotations(oracle_mod):
    """a rotated diag(3, 2, 1) cell is brought back with its longest vector along +x and the plane of the other two
    containing the x axis (test/internals/CellOperations.jl:138-150)"""
    rng = np.random.default_rng(4)
    m = np.diag([3.0, 2.0, 1.0])
    for _ in range(5):
        q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        b = oracle_mod.Oracle(np.zeros((1, 3)), 0.05, unitcell=q @ m, triclinic=True).box()
        a = b["aligned_unit_cell"]
        assert np.allclose(a[:, 0], m[:, 0], atol=1e-12)
        assert np.allclose(np.cross([1.0, 0.0, 0.0], np.cross(a[:, 1], a[:, 2])), 0.0, atol=1e-10)

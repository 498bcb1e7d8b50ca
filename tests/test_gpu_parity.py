"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): neighbour pair sets bit-exact after (i,j) sorting -- and, because the device
sweep visits every cell pair from the same home cell as the reference and repeats its arithmetic unfused, the
distances are bit-exact too; counts / histograms exact; energies, forces and sums within 1e-10 (Float64) /
1e-5 (Float32) relative.
"""
import numpy as np
import pytest

import workloads as W
from golden import kats as K

pytestmark = pytest.mark.gpu

from parity_util import RTOL, conditioning_bound, force_report


@pytest.fixture(scope="module")
def clm():
    import celllistmap_b200 as c
    return c


def canon(rec_or_tuple):
    if isinstance(rec_or_tuple, tuple):
        i, j, d = rec_or_tuple
    else:
        i, j, d = rec_or_tuple["i"], rec_or_tuple["j"], rec_or_tuple["d"]
    a, b = np.minimum(i, j), np.maximum(i, j)
    o = np.lexsort((b, a))
    return a[o], b[o], np.asarray(d)[o]


def canon_cross(rec_or_tuple):
    if isinstance(rec_or_tuple, tuple):
        i, j, d = rec_or_tuple
    else:
        i, j, d = rec_or_tuple["i"], rec_or_tuple["j"], rec_or_tuple["d"]
    o = np.lexsort((j, i))
    return np.asarray(i)[o], np.asarray(j)[o], np.asarray(d)[o]


def assert_lists_identical(got, want, cross=False):
    f = canon_cross if cross else canon
    gi, gj, gd = f(got)
    wi, wj, wd = f(want)
    assert gi.shape == wi.shape, f"pair count differs: {gi.shape[0]} vs oracle {wi.shape[0]}"
    assert np.array_equal(gi, wi) and np.array_equal(gj, wj), "pair sets differ"
    assert np.array_equal(gd.view(np.uint8), wd.view(np.uint8)), "distances are not bit-identical"


def random_system(rng, n, dim, kind, dtype, scale=10.0):
    """positions spilling outside the cell (exercises wrapping), cell of `kind`, cutoff < half the smallest height."""
    if kind == "ortho":
        uc = (scale * (0.8 + 0.4 * rng.random(dim))).astype(dtype)
    elif kind == "triclinic":
        uc = np.diag(scale * (0.8 + 0.4 * rng.random(dim)))
        uc += np.triu(scale * 0.3 * (rng.random((dim, dim)) - 0.3), 1)
        uc = uc.astype(dtype)
    else:
        uc = None
    x = (scale * (1.6 * rng.random((n, dim)) - 0.3)).astype(dtype)
    return x, uc


# ---------------------------------------------------------------------------------------------------------
# neighbour lists: bit-exact vs the oracle
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("lcell", [1, 2, 3])
def test_neighborlist_self_bit_exact(clm, oracle_mod, dtype, kind, dim, lcell):
    rng = np.random.default_rng(7 + 13 * dim + lcell)
    n = 3000 if dim == 3 else 1500
    x, uc = random_system(rng, n, dim, kind, dtype)
    cutoff = 1.2
    got = clm.neighborlist(xpositions=x, cutoff=cutoff, unitcell=uc, lcell=lcell)
    want = oracle_mod.Oracle(x, cutoff, unitcell=uc, lcell=lcell, dtype=dtype).neighborlist()
    assert len(want[0]) > 100
    assert_lists_identical(got, want)


@pytest.mark.parametrize("kind,dim,dtype,lcell", [("ortho", 3, np.float64, 9), ("triclinic", 3, np.float32, 8), ("nonperiodic", 2, np.float64, 12),
                                                  ("triclinic", 2, np.float64, 15), ("ortho", 3, np.float32, 5)])
def test_neighborlist_large_lcell_bit_exact(clm, oracle_mod, kind, dim, dtype, lcell):
    """lcell beyond the usual 1-3 (the reference accepts any lcell >= 1, src/internals/Box.jl:192): cells of cutoff / lcell,
    (2 lcell + 1)^(N-1) stencil rows per tile; no sub-cell split beyond lcell = 7."""
    rng = np.random.default_rng(31 + lcell)
    x, uc = random_system(rng, 2000 if dim == 3 else 1200, dim, kind, dtype)
    cutoff = 1.2
    got = clm.neighborlist(xpositions=x, cutoff=cutoff, unitcell=uc, lcell=lcell)
    want = oracle_mod.Oracle(x, cutoff, unitcell=uc, lcell=lcell, dtype=dtype).neighborlist()
    assert len(want[0]) > 100
    assert_lists_identical(got, want)
    if kind != "nonperiodic" and dim == 3 and dtype == np.float64:     # the force sweeps walk the same stencil tables
        n = x.shape[0]
        sys_ = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, lcell=lcell, output=clm.EnergyAndForces(0.0, np.zeros((n, 3), dtype)))
        out = clm.pairwise(clm.LJEnergyAndForces(1.0e-3, 1.0e-6), sys_)
        we, wf = oracle_mod.Oracle(x, cutoff, unitcell=uc, lcell=lcell, dtype=dtype).lj(1.0e-3, 1.0e-6, forces=True)
        assert abs(out.energy - we) <= 1e-10 * abs(we)
        assert np.abs(np.asarray(out.forces) - wf).max() <= 1e-10 * np.abs(wf).max()
    if kind == "ortho" and dtype == np.float32:
        # Float32 forces take the Newton's-third-law sweep, which walks the stencil tables with its own code: a missed or doubled
        # pair would be an O(1) error; random close pairs make anything tighter than the pair-set check a conditioning question
        n = x.shape[0]
        sys_ = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, lcell=lcell, output=clm.EnergyAndForces(0.0, np.zeros((n, 3), dtype)))
        out = clm.pairwise(clm.LJEnergyAndForces(1.0e-3, 1.0e-6), sys_)
        we, wf = oracle_mod.Oracle(x.astype(np.float64), cutoff, unitcell=uc.astype(np.float64), lcell=lcell).lj(1.0e-3, 1.0e-6, forces=True)
        assert abs(out.energy - we) <= 1e-3 * abs(we)
        assert np.abs(np.asarray(out.forces) - wf).max() <= 1e-3 * np.abs(wf).max()


def test_lcell_beyond_the_stencil_table_is_refused(clm):
    x = np.random.default_rng(1).random((100, 3))
    with pytest.raises(Exception, match="lcell"):
        clm.neighborlist(xpositions=x, cutoff=0.1, unitcell=[1.0, 1.0, 1.0], lcell=16)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("dim", [2, 3])
def test_neighborlist_cross_bit_exact(clm, oracle_mod, dtype, kind, dim):
    rng = np.random.default_rng(99 + dim)
    x, uc = random_system(rng, 1200, dim, kind, dtype)
    y, _ = random_system(rng, 2100, dim, kind, dtype)
    cutoff = 1.1
    got = clm.neighborlist(xpositions=x, ypositions=y, cutoff=cutoff, unitcell=uc)
    want = oracle_mod.Oracle(x, cutoff, unitcell=uc, y=y, dtype=dtype).neighborlist()
    assert len(want[0]) > 100
    assert_lists_identical(got, want, cross=True)


def test_c1_neighborlist_config(clm, oracle_mod):
    """BASELINE.json configs[0]: 10k random 3-D particles, orthorhombic box, cutoff 0.1 L, Float64."""
    w = W.c1_neighborlist()
    got = clm.neighborlist(xpositions=w["x"], cutoff=w["cutoff"], unitcell=w["unitcell"])
    want = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"]).neighborlist()
    assert 200_000 < len(want[0]) < 220_000
    assert_lists_identical(got, want)
    # no duplicates
    a, b, _ = canon(got)
    assert len(np.unique(np.stack([a, b], 1), axis=0)) == len(a)


def test_inplace_neighborlist_reuses_a_page_locked_buffer(clm, oracle_mod):
    """InPlaceNeighborList (src/API/neighborlist.jl:84-111) overwrites its record array in place; from the second list on the
    array is page-locked (clm_host_register) so that the copy-out is a direct DMA transfer.  Lists stay bit-identical across
    the switch, across position updates, and across a regrowth of the array (which unpins the old one)."""
    w = W.c1_neighborlist(4000)
    nb = clm.InPlaceNeighborList(x=w["x"], cutoff=w["cutoff"], unitcell=w["unitcell"])
    want = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"]).neighborlist()
    for k in range(3):
        clm.update(nb, xpositions=w["x"])
        assert_lists_identical(nb.neighborlist().copy(), want)
        assert nb._pinned == (k >= 1)
    x2 = np.mod(w["x"] + 0.013, 1.0)
    clm.update(nb, xpositions=x2)
    assert_lists_identical(nb.neighborlist().copy(), oracle_mod.Oracle(x2, w["cutoff"], unitcell=w["unitcell"]).neighborlist())
    assert nb._pinned
    clm.update(nb, xpositions=x2, cutoff=1.5 * w["cutoff"])          # ~3.4x the pairs: the array regrows, unpinned until reused again
    assert_lists_identical(nb.neighborlist().copy(), oracle_mod.Oracle(x2, 1.5 * w["cutoff"], unitcell=w["unitcell"]).neighborlist())
    assert not nb._pinned
    assert_lists_identical(nb.neighborlist().copy(), oracle_mod.Oracle(x2, 1.5 * w["cutoff"], unitcell=w["unitcell"]).neighborlist())
    assert nb._pinned


# ---------------------------------------------------------------------------------------------------------
# reductions
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("two", [False, True])
def test_sum_d_d2(clm, oracle_mod, dtype, kind, two):
    rng = np.random.default_rng(5)
    x, uc = random_system(rng, 4000, 3, kind, dtype)
    y = random_system(rng, 3000, 3, kind, dtype)[0] if two else None
    sys = clm.ParticleSystem(xpositions=x, ypositions=y, unitcell=uc, cutoff=1.3, output=None)
    sd, sd2, n = clm.pairwise(clm.SumDistances(), sys)
    o = oracle_mod.Oracle(x, 1.3, unitcell=uc, y=y, dtype=dtype)
    wsd, wsd2, wn = o.sum_d_d2()
    assert n == wn
    tol = RTOL[np.dtype(dtype)]
    # the oracle accumulates in T; compare against an exact-ish sum from its neighbour list
    d = o.neighborlist()[2].astype(np.float64)
    assert abs(sd - d.sum()) <= tol * d.sum()
    assert abs(sd2 - (d * d).sum()) <= 10 * tol * (d * d).sum()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("dim", [2, 3])
def test_lj_energy_forces(clm, oracle_mod, dtype, kind, dim):
    rng = np.random.default_rng(11 + dim)
    # jittered lattice: no r -> 0 blow-ups
    m = 14 if dim == 3 else 40
    g = np.stack(np.meshgrid(*[np.arange(m)] * dim, indexing="ij"), -1).reshape(-1, dim).astype(np.float64)
    x = (g + 0.25 * (rng.random(g.shape) - 0.5)).astype(dtype)
    rng.shuffle(x)
    if kind == "ortho":
        uc = np.full(dim, float(m), dtype)
    elif kind == "triclinic":
        uc = (np.eye(dim) * m + np.triu(np.full((dim, dim), 0.2 * m), 1)).astype(dtype)
    else:
        uc = None
    cutoff, c6, c12 = 2.7, 4.0, 4.0
    n = x.shape[0]
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff,
                             output=clm.EnergyAndForces(0.0, np.zeros((n, dim), dtype)))
    out = clm.pairwise(clm.LJEnergyAndForces(c6, c12), sys)
    o64 = oracle_mod.Oracle(x.astype(np.float64), cutoff, unitcell=None if uc is None else uc.astype(np.float64))
    we, wf = o64.lj(c6, c12, forces=True)
    tol = RTOL[np.dtype(dtype)]
    escale = np.abs(oracle_mod.Oracle(x.astype(np.float64), cutoff, unitcell=None if uc is None else uc.astype(np.float64)).lj(c6, -c12))
    assert abs(out.energy - we) <= tol * max(abs(we), escale)
    # forces: the bar holds against the oracle in the SAME precision (bit-identical wrapped coordinates); the distance to
    # Float64 arithmetic is reported and bounded by the conditioning of the coordinates (tests/parity_util.py)
    fscale = np.abs(wf).max()
    ftol = tol * fscale
    f_same = wf if dtype == np.float64 else oracle_mod.Oracle(x, cutoff, unitcell=uc, dtype=dtype).lj(c6, c12, forces=True)[1]
    err_same, err_64, _ = force_report(f"LJ {kind} {dim}-D {np.dtype(dtype).name}", out.forces, f_same, wf)
    assert err_same <= tol
    assert err_64 <= max(tol, conditioning_bound(dtype, 1.3 * m, 0.75, 12))
    # energy-only map (reference's exactly-once sweep) agrees too
    sys2 = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, output=0.0)
    e2 = clm.pairwise(clm.LJEnergy(c6, c12), sys2)
    assert abs(e2 - we) <= tol * max(abs(we), escale)
    # reset=False accumulates on the previous value (API/pairwise.jl:52-54)
    e3 = clm.pairwise(clm.LJEnergy(c6, c12), sys2, reset=False)
    assert abs(e3 - 2 * we) <= 2 * tol * max(abs(we), escale)
    f_before = out.forces.copy()
    clm.pairwise(clm.LJEnergyAndForces(c6, c12), sys, reset=False)
    assert np.abs(sys.output.forces - 2 * f_before).max() <= ftol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lj_cross_forces(clm, oracle_mod, dtype):
    rng = np.random.default_rng(3)
    m = 12
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    pts = (g + 0.25 * (rng.random(g.shape) - 0.5)).astype(dtype)
    rng.shuffle(pts)
    x, y = pts[:700], pts[700:]
    uc = np.full(3, float(m), dtype)
    sys = clm.ParticleSystem(xpositions=x, ypositions=y, unitcell=uc, cutoff=2.5,
                             output=clm.EnergyAndForces(0.0, np.zeros((700, 3), dtype)))
    out = clm.pairwise(clm.LJEnergyAndForces(4.0, 4.0), sys)
    we, wf = oracle_mod.Oracle(x.astype(np.float64), 2.5, unitcell=uc.astype(np.float64), y=y.astype(np.float64)).lj(4.0, 4.0, forces=True)
    tol = RTOL[np.dtype(dtype)]
    # mixed-sign energy: the scale of the sum is the sum of the magnitudes (attractive + repulsive parts)
    escale = abs(oracle_mod.Oracle(x.astype(np.float64), 2.5, unitcell=uc.astype(np.float64), y=y.astype(np.float64)).lj(4.0, -4.0))
    print(f"[parity] LJ cross {np.dtype(dtype).name}: |E - E_oracle64| / sum|terms| = {abs(out.energy - we) / escale:.3e}")
    assert abs(out.energy - we) <= tol * max(abs(we), escale)
    f_same = wf if dtype == np.float64 else oracle_mod.Oracle(x, 2.5, unitcell=uc, y=y, dtype=dtype).lj(4.0, 4.0, forces=True)[1]
    err_same, err_64, _ = force_report(f"LJ cross {np.dtype(dtype).name}", out.forces, f_same, wf)
    assert err_same <= tol
    assert err_64 <= max(tol, conditioning_bound(dtype, 1.3 * m, 0.75, 12))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic"])
def test_coulomb(clm, oracle_mod, dtype, kind):
    rng = np.random.default_rng(8)
    m = 12
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    x = (g + 0.3 * (rng.random(g.shape) - 0.5)).astype(dtype)
    rng.shuffle(x)
    w = (0.5 + rng.random(x.shape[0])).astype(dtype)
    uc = np.full(3, float(m), dtype) if kind == "ortho" else (np.eye(3) * m + np.triu(np.full((3, 3), 2.0), 1)).astype(dtype)
    n = x.shape[0]
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=2.6, output=clm.EnergyAndForces(0.0, np.zeros((n, 3), dtype)))
    out = clm.pairwise(clm.CoulombEnergyAndForces(-9.8, w), sys)
    we, wf = oracle_mod.Oracle(x.astype(np.float64), 2.6, unitcell=uc.astype(np.float64)).coulomb(-9.8, w.astype(np.float64), forces=True)
    tol = RTOL[np.dtype(dtype)]
    print(f"[parity] Coulomb {kind} {np.dtype(dtype).name}: |E - E_oracle64| / |E| = {abs(out.energy - we) / abs(we):.3e}")
    assert abs(out.energy - we) <= tol * abs(we)
    f_same = wf if dtype == np.float64 else oracle_mod.Oracle(x, 2.6, unitcell=uc, dtype=dtype).coulomb(-9.8, w, forces=True)[1]
    err_same, err_64, _ = force_report(f"Coulomb {kind} {np.dtype(dtype).name}", out.forces, f_same, wf)
    assert err_same <= tol
    assert err_64 <= max(tol, conditioning_bound(dtype, 1.3 * m, 0.7, 2))
    e = clm.pairwise(clm.CoulombEnergy(-9.8, w), clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=2.6, output=0.0))
    assert abs(e - we) <= tol * abs(we)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nbins", [5, 40])
def test_distance_histogram_exact(clm, oracle_mod, dtype, nbins):
    rng = np.random.default_rng(21)
    x, uc = random_system(rng, 5000, 3, "ortho", dtype)
    cutoff = 2.0
    width = cutoff / nbins
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, output=np.zeros(nbins, np.int64))
    h = clm.pairwise(clm.DistanceHistogram(width), sys)
    want = oracle_mod.Oracle(x, cutoff, unitcell=uc, dtype=dtype).dist_hist(width, nbins)
    assert np.array_equal(h, want)
    h2 = clm.pairwise(clm.DistanceHistogram(width), sys, reset=False).copy()
    assert np.array_equal(h2, 2 * want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("width,nbins", [(0.25, 8), (0.5, 4), (0.1, 20), (1.0 / 3.0, 6), (0.0625, 32)])
def test_histograms_distances_on_bin_edges(clm, oracle_mod, dtype, width, nbins):
    """lattice positions (spacing 0.5, exactly representable): many distances sit EXACTLY on bin edges, which is where
    the table-driven bin search of the device (d2 thresholds of the correctly rounded sqrt / division) must agree with
    the direct floor(sqrt(d2) / width) and searchsortedfirst(rbins, sqrt(d2)) of the reference"""
    g = np.arange(12) * 0.5
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(dtype)
    uc = np.array([6.0, 6.0, 6.0], dtype)
    cutoff = 2.0
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, output=np.zeros(nbins, np.int64))
    h = clm.pairwise(clm.DistanceHistogram(dtype(width)), sys)
    o = oracle_mod.Oracle(x, cutoff, unitcell=uc, dtype=dtype)
    assert np.array_equal(h, o.dist_hist(dtype(width), nbins))
    assert h.sum() > 10000
    # pair-velocity bins: edges at multiples of the same width (right-closed bins)
    nb = min(nbins, int(cutoff / width))
    rbins = (np.arange(nb + 1) * dtype(width)).astype(dtype)
    v = np.random.default_rng(3).random(x.shape).astype(dtype)
    sys2 = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, output=(np.zeros(nb, np.int64), np.zeros(nb, dtype)))
    counts, _ = clm.pairwise(clm.PairwiseVelocities(rbins, v), sys2)
    assert np.array_equal(counts, o.pairvel(v, rbins)[0])


def test_pairwise_velocities_pathological_edges(clm, oracle_mod):
    """bin edges whose squares underflow (the host cannot tabulate exact d2 thresholds for them: the map falls back to the
    direct sqrt form) and more edges than fit the kernel parameters: counts stay exact"""
    rng = np.random.default_rng(23)
    x, uc = random_system(rng, 2500, 3, "ortho", np.float64)
    v = rng.random(x.shape)
    o = oracle_mod.Oracle(x, 2.0, unitcell=uc)
    for rbins in (np.array([0.0, 1e-200, 0.4, 0.8, 1.2, 1.6, 2.0]), np.linspace(0.0, 2.0, 41), np.linspace(0.0, 2.0, 14)):
        nb = len(rbins) - 1
        sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=2.0, output=(np.zeros(nb, np.int64), np.zeros(nb)))
        counts, sums = clm.pairwise(clm.PairwiseVelocities(rbins, v), sys)
        wc, ws = o.pairvel(v, rbins)
        assert np.array_equal(counts, wc)
        assert np.abs(sums - ws).max() <= 1e-10 * max(np.abs(ws).max(), 1.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", ["ortho", "triclinic"])
def test_pairwise_velocities(clm, oracle_mod, dtype, dim, kind):
    rng = np.random.default_rng(17)
    x, uc = random_system(rng, 3000, dim, kind, dtype)
    v = rng.random(x.shape).astype(dtype)
    rbins = np.array([0.0, 0.4, 0.8, 1.2, 1.6, 2.0], dtype)
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=2.0, output=(np.zeros(5, np.int64), np.zeros(5, dtype)))
    counts, sums = clm.pairwise(clm.PairwiseVelocities(rbins, v), sys)
    # the oracle in the SAME precision sees the same pair set bit for bit: counts are exact; sums are accumulated in a
    # different order (and against Float64 to bound the Float32 error)
    wc, ws = oracle_mod.Oracle(x, 2.0, unitcell=uc, dtype=dtype).pairvel(v, rbins)
    assert np.array_equal(counts, wc)
    # sums of SIGNED terms dv.r/|r| (uncorrelated velocities: the sum itself is ~ sqrt(count), not a scale): the bar is
    # relative to the sum of the term magnitudes, ~0.3 per pair for velocities uniform in [0,1)^dim.  Reference: the oracle
    # in the same precision -- the SAME pair set; Float64 arithmetic on Float32 inputs moves a pair that sits within
    # rounding of the cutoff across it, which changes a bin's sum by one whole term
    err = np.abs(sums - ws) / (0.3 * np.maximum(counts, 1))
    print(f"[parity] pair velocities {kind} {dim}-D {np.dtype(dtype).name}: max |sum - oracle| / sum|terms| = {err.max():.3e}")
    assert err.max() <= RTOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("two", [False, True])
def test_minimum_distance(clm, oracle_mod, dtype, kind, two):
    rng = np.random.default_rng(23)
    x, uc = random_system(rng, 2500, 3, kind, dtype)
    y = random_system(rng, 1500, 3, kind, dtype)[0] if two else None
    sys = clm.ParticleSystem(xpositions=x, ypositions=y, unitcell=uc, cutoff=1.5, output=clm.MinimumDistance())
    md = clm.pairwise(clm.MinimumDistanceMap(), sys)
    wi, wj, wd = oracle_mod.Oracle(x, 1.5, unitcell=uc, y=y, dtype=dtype).mindist()
    assert md.d == wd
    assert (md.i, md.j) == (wi, wj) if two else {md.i, md.j} == {wi, wj}


# ---------------------------------------------------------------------------------------------------------
# golden fixtures of the reference
def test_golden_argon(clm):
    x = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "argon_cubic.npy"))
    sys = clm.ParticleSystem(xpositions=x, unitcell=K.ARGON_UNITCELL, cutoff=K.ARGON_CUTOFF, output=None)
    sd, sd2, n = clm.pairwise(clm.SumDistances(), sys)
    assert n == K.ARGON_NL_PERIODIC[0]
    assert abs(sd2 - K.ARGON_SUM_D2) <= 1e-12 * K.ARGON_SUM_D2
    # the doctest distances were printed from Float32-rounded coordinates (see tests/test_oracle_golden.py)
    x = x.astype(np.float32).astype(np.float64)
    nl = clm.neighborlist(xpositions=x, cutoff=K.ARGON_CUTOFF)
    assert len(nl) == K.ARGON_NL_NONPERIODIC[0]
    a, b, d = canon(nl)
    i0, j0, d0 = K.ARGON_NL_NONPERIODIC[1]
    k = np.where((a == i0) & (b == j0))[0]
    assert len(k) == 1 and d[k[0]] == d0
    nlc = clm.neighborlist(xpositions=x[:50], ypositions=x[50:], cutoff=K.ARGON_CUTOFF)
    assert len(nlc) == K.ARGON_NL_CROSS[0]
    x = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "argon_cubic.npy"))
    md = clm.pairwise(clm.MinimumDistanceMap(), clm.ParticleSystem(xpositions=x, unitcell=K.ARGON_UNITCELL, cutoff=K.ARGON_CUTOFF,
                                                                  output=clm.MinimumDistance()))
    assert abs(md.d - K.ARGON_MIN_DIST) < 1e-14


@pytest.mark.parametrize("frame", sorted(K.NAMD_CASES))
@pytest.mark.parametrize("lcell", K.NAMD_LCELLS)
def test_golden_namd_energies(clm, frame, lcell):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "namd_frames.npz"))
    x = z[frame].astype(np.float64)
    uc, golden = K.NAMD_CASES[frame]
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=K.NAMD_CUTOFF, output=0.0, lcell=lcell)
    e = clm.pairwise(clm.LJEnergy(K.NAMD_C6, K.NAMD_C12), sys)
    assert abs(e - golden) <= 1e-9 * abs(golden)


def test_grid_kats(clm):
    for uc, cutoff, lcell, nc, cs in K.GRID_KATS:
        dim = len(nc)
        sys = clm.ParticleSystem(xpositions=np.zeros((1, dim)), unitcell=uc, cutoff=cutoff, output=0.0, lcell=lcell)
        b = sys.box
        assert b.nc.tolist() == nc
        assert np.allclose(b.cell_size, cs, rtol=1e-14)
    k = K.SHOW_CELLLIST_KAT
    x = np.random.default_rng(1).random((k["n"], 3))
    sys = clm.ParticleSystem(xpositions=x, unitcell=k["unitcell"], cutoff=k["cutoff"], output=0.0)
    st = sys.stats()
    assert st.n_cells_real[0] == k["n_cells_real"] and st.n_total[0] == k["n_particles"]
    c = K.COMPUTING_BOX_KAT
    lo, hi = clm.get_computing_box(clm.ParticleSystem(xpositions=np.zeros((1, 3)), unitcell=c["unitcell"], cutoff=c["cutoff"], output=0.0))
    assert np.allclose(lo, c["lo"], atol=1e-15) and np.allclose(hi, c["hi"], atol=1e-15)
    # align_cell + cell_limits of the reference (test/internals/CellOperations.jl:159-194) through the product's Box
    for m, want_lo, want_hi, _ in K.CELL_LIMITS_KATS:
        m = np.array(m, dtype=np.float64)
        rc = 0.05
        lo, hi = clm.get_computing_box(clm.ParticleSystem(xpositions=np.zeros((1, m.shape[0])), unitcell=m, cutoff=rc, output=0.0))
        assert np.allclose(lo + rc, want_lo, rtol=1e-12, atol=1e-10) and np.allclose(hi - rc, want_hi, rtol=1e-12, atol=1e-10)


def test_boundary_kats(clm):
    for x, cutoff, uc, npairs in K.BOUNDARY_KATS:
        nl = clm.neighborlist(xpositions=np.array(x), cutoff=cutoff, unitcell=uc)
        assert len(nl) == npairs, (x, cutoff, uc)
    x, y, cutoff, (i, j, d) = K.FEW_CROSS
    nl = clm.neighborlist(xpositions=np.array(x), ypositions=np.array(y), cutoff=cutoff)
    assert len(nl) == 1 and (nl["i"][0], nl["j"][0]) == (i, j) and abs(nl["d"][0] - d) < 1e-12
    x, cutoff, (i, j, d) = K.FEW_SELF
    nl = clm.neighborlist(xpositions=np.array(x), cutoff=cutoff)
    assert len(nl) == 1 and {nl["i"][0], nl["j"][0]} == {i, j}
    for pts in (K.BUG84_3D, K.BUG84_2D):
        nl = clm.neighborlist(xpositions=np.array(pts, np.float32), cutoff=np.float32(7.0))
        a, b, _ = canon(nl)
        assert len(np.unique(np.stack([a, b], 1), axis=0)) == len(a)


# ---------------------------------------------------------------------------------------------------------
# API behaviour and error mapping
def test_errors_and_updates(clm, oracle_mod):
    x = np.random.default_rng(2).random((500, 3))
    with pytest.raises(ValueError, match="positions` OR `xpositions"):
        clm.ParticleSystem(unitcell=[1, 1, 1], cutoff=0.1, output=0.0)
    with pytest.raises(ValueError, match="Unit cell matrix does not satisfy"):
        clm.ParticleSystem(xpositions=x, unitcell=[1, 1, 1], cutoff=0.6, output=0.0)
    bad = x.copy()
    bad[41, 1] = np.nan
    with pytest.raises(ValueError, match="Invalid coordinates.*42"):
        clm.ParticleSystem(xpositions=bad, unitcell=[1, 1, 1], cutoff=0.1, output=0.0)
    with pytest.raises(ValueError, match="lcell"):
        clm.ParticleSystem(xpositions=x, unitcell=[1, 1, 1], cutoff=0.1, output=0.0, lcell=0)
    sys = clm.ParticleSystem(xpositions=x, unitcell=[1, 1, 1], cutoff=0.1, output=None)
    with pytest.raises(ValueError, match="ypositions can only"):
        clm.update(sys, ypositions=x)
    n1 = clm.pairwise(clm.SumDistances(), sys)[2]
    assert n1 == oracle_mod.Oracle(x, 0.1, unitcell=[1.0, 1, 1]).sum_d_d2()[2]
    # update!: new coordinates (different count), cutoff and unit cell; nothing recomputed until the next map
    x2 = np.random.default_rng(3).random((800, 3)) * 2
    clm.update(sys, xpositions=x2, cutoff=0.2, unitcell=[2, 2, 2])
    n2 = clm.pairwise(clm.SumDistances(), sys)[2]
    assert n2 == oracle_mod.Oracle(x2, 0.2, unitcell=[2.0, 2, 2]).sum_d_d2()[2]
    # element mutation sets the updated flag (ParticleSystemPositions)
    sys.xpositions[0] = sys.xpositions[1] + 0.01
    x3 = np.array(sys.xpositions)
    assert clm.pairwise(clm.SumDistances(), sys)[2] == oracle_mod.Oracle(x3, 0.2, unitcell=[2.0, 2, 2]).sum_d_d2()[2]
    nps = clm.ParticleSystem(xpositions=x, cutoff=0.1, output=None)
    with pytest.raises(ValueError, match="non-periodic"):
        clm.update(nps, unitcell=[1, 1, 1])


def test_tiny_and_empty_systems(clm):
    for n in (0, 1, 2):
        x = np.random.default_rng(n).random((n, 3))
        sys = clm.ParticleSystem(xpositions=x, unitcell=[1, 1, 1], cutoff=0.1, output=None)
        assert clm.pairwise(clm.SumDistances(), sys)[2] == 0 or n == 2
        assert len(clm.neighborlist(xpositions=x, cutoff=0.1, unitcell=[1, 1, 1])) in (0, 1)
    x = np.random.default_rng(5).random((100, 3))
    nl = clm.neighborlist(xpositions=x, ypositions=np.zeros((0, 3)), cutoff=0.3, unitcell=[1, 1, 1])
    assert len(nl) == 0


def test_inplace_neighborlist_reuse(clm, oracle_mod):
    rng = np.random.default_rng(31)
    x = rng.random((2000, 3))
    nb = clm.InPlaceNeighborList(x=x, cutoff=0.1, unitcell=[1, 1, 1])
    assert_lists_identical(clm.neighborlist_(nb).copy(), oracle_mod.Oracle(x, 0.1, unitcell=[1.0, 1, 1]).neighborlist())
    x2 = rng.random((3000, 3))
    clm.update(nb, xpositions=x2, cutoff=0.12)
    assert_lists_identical(clm.neighborlist_(nb).copy(), oracle_mod.Oracle(x2, 0.12, unitcell=[1.0, 1, 1]).neighborlist())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
def test_energy_only_lean_sweep_equals_mode_half(clm, oracle_mod, dtype, kind):
    """LJ energy without forces: the lean partner-per-lane sweep (option "n3" = 1; the Float32 default) visits the pair set of
    k_sweep<MODE_HALF / MODE_TRI> (option "n3" = 0; the Float64 default), so the two energies agree to the rounding of the
    order-free sums, and both meet the north_star bar against the oracle in Float64 (scale: the sum of the term magnitudes, as
    in test_lj_energy_forces)."""
    rng = np.random.default_rng(5)
    x, uc = random_system(rng, 6000, 3, kind, dtype, scale=12.6)
    cutoff, c6, c12 = 2.7, 4.0, 4.0
    uc64 = None if uc is None else uc.astype(np.float64)
    want = float(oracle_mod.Oracle(x.astype(np.float64), cutoff, unitcell=uc64).lj(c6, c12, forces=False))
    escale = abs(float(oracle_mod.Oracle(x.astype(np.float64), cutoff, unitcell=uc64).lj(c6, -c12, forces=False)))
    got = {}
    for n3 in (1, 0):
        sys_ = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff, output=0.0)
        sys_._h.set_option("n3", n3)
        got[n3] = float(clm.pairwise(clm.LJEnergy(c6, c12), sys_))
    scale = max(abs(want), escale)
    tol = RTOL[np.dtype(dtype)]
    # distance to Float64 arithmetic: bounded by the conditioning of the wrapped Float32 coordinates at the closest pair, which
    # dominates a random system's energy (tests/parity_util.py); the bar proper is the agreement of the two sweeps
    dmin = float(oracle_mod.Oracle(x.astype(np.float64), cutoff, unitcell=uc64).neighborlist()[2].min())
    bar64 = max(tol, conditioning_bound(dtype, 1.3 * 12.6, dmin, 12))
    print(f"[parity] energy-only {kind} {np.dtype(dtype).name}: lean vs MODE_HALF {abs(got[1] - got[0]) / scale:.2e}; lean {abs(got[1] - want) / scale:.2e}, "
          f"MODE_HALF {abs(got[0] - want) / scale:.2e} of the Float64 oracle (closest pair {dmin:.3g}, conditioning bound {bar64:.1e})")
    assert abs(got[1] - got[0]) <= tol * scale
    assert abs(got[1] - want) <= bar64 * scale and abs(got[0] - want) <= bar64 * scale


@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("cap", [1, 3])
def test_capped_bin_grid_gives_the_same_list(clm, oracle_mod, kind, cap):
    """clm_set_option "bin_blocks_per_sm": the binning kernel strides over the particles with a capped grid (120 000 particles on
    148 x cap blocks of 256: several trips per block, a partial last one); a tuning knob must not change the result."""
    rng = np.random.default_rng(97)
    x, uc = random_system(rng, 120000, 3, kind, np.float64, scale=30.0)
    nb = clm.InPlaceNeighborList(x=x, cutoff=1.0, unitcell=uc)
    nb.sys._h.set_option("bin_blocks_per_sm", cap)
    clm.update(nb, xpositions=x)
    got = clm.neighborlist_(nb).copy()
    assert_lists_identical(got, oracle_mod.Oracle(x, 1.0, unitcell=uc).neighborlist())


# ---------------------------------------------------------------------------------------------------------
# full-size property checks (BASELINE.json configs[1]): what the domain offers independent of size
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c2_full_size_properties(clm, dtype):
    w = W.c2_argon(100, dtype)
    n = w["x"].shape[0]
    sys = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"],
                             output=clm.EnergyAndForces(0.0, np.zeros((n, 3), dtype)))
    out = clm.pairwise(clm.LJEnergyAndForces(w["c6"], w["c12"]), sys)
    f = out.forces.astype(np.float64)
    fmax = np.abs(f).max()
    # Newton's third law: the net force vanishes
    assert np.abs(f.sum(0)).max() <= (1e-9 if dtype == np.float64 else 2e-3) * fmax * np.sqrt(n)
    # energy of the exactly-once sweep == half-summed full-shell energy
    e_once = clm.pairwise(clm.LJEnergy(w["c6"], w["c12"]), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=0.0))
    assert abs(e_once - out.energy) <= RTOL[np.dtype(dtype)] * abs(e_once)
    # pair count == analytic expectation within statistics, and equals the list length of the neighbour list
    sd, sd2, npairs = clm.pairwise(clm.SumDistances(), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=None))
    expect = 0.5 * n * W.ARGON_RHO * 4.0 / 3.0 * np.pi * w["cutoff"] ** 3
    assert abs(npairs - expect) < 0.01 * expect
    # translation invariance under a lattice vector + permutation invariance of the scalar results
    perm = np.random.default_rng(0).permutation(n)
    x2 = (w["x"][perm].astype(np.float64) + np.array([w["L"], -w["L"], 0.0])).astype(dtype)
    sd_b, sd2_b, npairs_b = clm.pairwise(clm.SumDistances(), clm.ParticleSystem(xpositions=x2, unitcell=w["unitcell"], cutoff=w["cutoff"], output=None))
    if dtype == np.float64:
        assert abs(npairs_b - npairs) <= 2 and abs(sd2_b - sd2) <= 1e-9 * sd2


# ---------------------------------------------------------------------------------------------------------
# capacity estimates that are too small: the engine must notice and repeat build / emission with the exact size
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_record_capacity_retry_clustered_corner(clm, oracle_mod, dtype):
    """every particle sits in a corner of the cell and has 7 periodic images: 8x the records of the uniform estimate"""
    rng = np.random.default_rng(77)
    x = (0.45 * rng.random((6000, 3))).astype(dtype)
    uc = np.array([10.0, 10.0, 10.0], dtype)
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=1.0, output=None)
    assert sys.stats().n_total[0] == 8 * 6000
    sd, sd2, n = clm.pairwise(clm.SumDistances(), sys)
    o = oracle_mod.Oracle(x, 1.0, unitcell=uc, dtype=dtype)
    assert n == o.sum_d_d2()[2]
    # the retry also happens when the build is only enqueued behind a map (fresh system, first call is a map)
    nl = clm.neighborlist(xpositions=x, cutoff=1.0, unitcell=uc)
    assert_lists_identical(nl, o.neighborlist())
    h = clm.Handle(3, dtype)
    h.set_box(clm._capi.ORTHORHOMBIC, uc, 1.0, 1)
    h.set_positions(0, x)
    e, f = np.zeros(1, dtype), np.zeros((6000, 3), dtype)
    h.map_lj(1e-4, 1e-8, e, f)          # first call on a dirty handle: enqueue, overflow, repeat
    we, wf = oracle_mod.Oracle(x.astype(np.float64), 1.0, unitcell=uc.astype(np.float64)).lj(1e-4, 1e-8, forces=True)
    # random positions in a corner: some pairs are nearly on top of each other and the largest force is conditioned by
    # rounding the coordinates alone; the bar holds against the oracle in the same precision
    f_same = wf if dtype == np.float64 else oracle_mod.Oracle(x, 1.0, unitcell=uc, dtype=dtype).lj(1e-4, 1e-8, forces=True)[1]
    err_same, _, _ = force_report(f"LJ clustered corner {np.dtype(dtype).name}", f, f_same, wf)
    assert err_same <= RTOL[np.dtype(dtype)]
    h.close()


def test_neighborlist_capacity_retry_clustered(clm, oracle_mod):
    """a dense blob in a large periodic cell has far more pairs than the uniform-density size hint"""
    rng = np.random.default_rng(78)
    x = 50.0 + 2.0 * rng.random((3000, 3))
    nl = clm.neighborlist(xpositions=x, cutoff=1.5, unitcell=[100.0, 100.0, 100.0])
    want = oracle_mod.Oracle(x, 1.5, unitcell=[100.0, 100.0, 100.0]).neighborlist()
    assert len(want[0]) > 1_000_000
    assert_lists_identical(nl, want)


def test_device_outputs_accumulate(clm, oracle_mod):
    """CLM_OUT_DEVICE without CLM_RESET adds onto the caller's device buffers"""
    import torch
    w = W.c2_argon(12, np.float64, cutoff=8.0)
    n = w["x"].shape[0]
    h = clm.Handle(3, np.float64)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    h.set_positions(0, torch.from_numpy(w["x"]).cuda())
    e = torch.full((1,), 5.0, dtype=torch.float64, device="cuda")
    f = torch.ones((n, 3), dtype=torch.float64, device="cuda")
    h.map_lj(w["c6"], w["c12"], e, f, reset=False)
    h.map_lj(w["c6"], w["c12"], e, f, reset=False)
    h.synchronize()
    we, wf = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"]).lj(w["c6"], w["c12"], forces=True)
    assert abs(float(e) - (5.0 + 2 * we)) <= 1e-10 * abs(we)
    assert np.abs(f.cpu().numpy() - (1.0 + 2 * wf)).max() <= 1e-10 * np.abs(wf).max()
    h.close()


# ---------------------------------------------------------------------------------------------------------
# at-cutoff band (north_star: pairs within 1 ulp of the cutoff are "reported separately"; the reference documents that
# such pairs may fall on either side, docs/src/neighborlists.md:12, the test being d2 <= cutoff_sqr, vicinal_cells.jl:35-36)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cutoff_band_reported(clm, oracle_mod, dtype):
    # lattice with spacing 0.5 and cutoff 2.0: every site has 6 partners at d2 == 4.0 EXACTLY (+-4 steps along an axis) and
    # no other pair within an ulp of it -> 3 band pairs per site, all of them inside the list (d2 <= cutoff^2)
    g = np.arange(12) * 0.5
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(dtype)
    uc = np.array([6.0, 6.0, 6.0], dtype)
    nb = clm.InPlaceNeighborList(x=x, cutoff=2.0, unitcell=uc)
    nl = nb.neighborlist()
    want = oracle_mod.Oracle(x, 2.0, unitcell=uc, dtype=dtype).neighborlist()
    assert_lists_identical(nl.copy(), want)
    on_cutoff = int((want[2] == dtype(2.0)).sum())
    assert on_cutoff == 3 * x.shape[0]
    assert nb.n_cutoff_band == on_cutoff
    # the reduction map reports the same band
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=2.0, output=None)
    clm.pairwise(clm.SumDistances(), sys)
    assert sys.stats().n_cutoff_band == on_cutoff
    # a cutoff one ulp below loses exactly those pairs, and they are still reported as the band
    rc = np.nextafter(dtype(2.0), dtype(0.0))
    nb2 = clm.InPlaceNeighborList(x=x, cutoff=rc, unitcell=uc)
    nl2 = nb2.neighborlist()
    want2 = oracle_mod.Oracle(x, rc, unitcell=uc, dtype=dtype).neighborlist()
    assert_lists_identical(nl2.copy(), want2)
    assert len(nl2) == len(nl) - on_cutoff
    # random positions: no pair within an ulp of the cutoff, the band is empty
    xr = np.random.default_rng(4).random((4000, 3)).astype(dtype)
    nb3 = clm.InPlaceNeighborList(x=xr, cutoff=0.11, unitcell=np.ones(3, dtype))
    nb3.neighborlist()
    assert nb3.n_cutoff_band == 0


# ---------------------------------------------------------------------------------------------------------
# pipelined frames (clm_set_positions_async + CLM_ASYNC): copy-in of frame k+1, compute of frame k and copy-out of
# frame k-1 overlap; the results must be those of the synchronous calls, frame by frame
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pipelined_frames_equal_synchronous(clm, dtype):
    import torch
    w = W.c2_argon(24, dtype)
    n = w["x"].shape[0]
    rng = np.random.default_rng(12)
    frames = [(w["x"] + (0.05 * k * rng.standard_normal(w["x"].shape)).astype(dtype)) for k in range(7)]
    frames[3] = frames[3][: n - 1000]   # a frame with a different particle count
    want = []
    hs = clm.Handle(3, dtype)
    hs.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    for x in frames:
        hs.set_positions(0, x)
        e, f = np.zeros(1, dtype), np.zeros((x.shape[0], 3), dtype)
        hs.map_lj(w["c6"], w["c12"], e, f)
        want.append((e.copy(), f.copy()))
    hs.close()
    h = clm.Handle(3, dtype)
    h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    xs = [torch.from_numpy(x).pin_memory() for x in frames]
    es = [torch.zeros(1, dtype=tdt).pin_memory() for _ in frames]
    fs = [torch.zeros((x.shape[0], 3), dtype=tdt).pin_memory() for x in frames]
    k = 0
    while k < len(frames):
        try:
            h.set_positions_async(0, xs[k].numpy())
            h.map_lj(w["c6"], w["c12"], es[k].numpy(), fs[k].numpy(), async_=True)
            k += 1
        except clm._capi.ClmError as err:
            # the record capacity of the PREVIOUS frame was too small (only possible on the first frames): repeat it
            assert err.code == 6   # CLM_ERR_CAPACITY
            k -= 1
    h.synchronize()
    for k, (we, wf) in enumerate(want):
        # the Float32 force sweep adds partner forces with floating-point reductions (order not fixed): a few ulp
        tol = 1e-6 if dtype == np.float32 else 1e-13
        assert abs(float(es[k][0]) - float(we[0])) <= tol * abs(float(we[0])), k
        assert np.abs(fs[k].numpy() - wf).max() <= tol * np.abs(wf).max(), k
    h.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("handles", [1, 2, 3])
def test_frame_pipeline_equals_synchronous(clm, dtype, handles):
    """clm.FramePipeline: independent frames through several handles in turn (the build of one frame next to the sweep of the
    previous one): every frame's energy and forces are those of a synchronous call."""
    import torch
    w = W.c2_argon(24, dtype)
    n = w["x"].shape[0]
    rng = np.random.default_rng(13)
    frames = [(w["x"] + (0.05 * k * rng.standard_normal(w["x"].shape)).astype(dtype)) for k in range(9)]
    frames[4] = frames[4][: n - 777]
    want = []
    hs = clm.Handle(3, dtype)
    hs.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
    for x in frames:
        hs.set_positions(0, x)
        e, f = np.zeros(1, dtype), np.zeros((x.shape[0], 3), dtype)
        hs.map_lj(w["c6"], w["c12"], e, f)
        want.append((e.copy(), f.copy()))
    hs.close()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    xs = [torch.from_numpy(x).pin_memory() for x in frames]
    es = [torch.zeros(1, dtype=tdt).pin_memory() for _ in frames]
    fs = [torch.zeros((x.shape[0], 3), dtype=tdt).pin_memory() for x in frames]
    pipe = clm.FramePipeline(3, dtype, w["unitcell"], w["cutoff"], handles=handles)
    for k in range(len(frames)):
        pipe.submit_lj(w["c6"], w["c12"], xs[k].numpy(), es[k].numpy(), fs[k].numpy())
    pipe.synchronize()
    for k, (we, wf) in enumerate(want):
        tol = 1e-6 if dtype == np.float32 else 1e-13
        assert abs(float(es[k][0]) - float(we[0])) <= tol * abs(float(we[0])), k
        assert np.abs(fs[k].numpy() - wf).max() <= tol * np.abs(wf).max(), k
    pipe.close()


@pytest.mark.parametrize("handles", [1, 2])
def test_frame_pipeline_repeats_a_frame_whose_record_capacity_was_too_small(clm, handles):
    """The record capacity of a build is estimated without a host round trip (uniform density); a frame with many more
    periodic images than that -- here every particle sits in a corner of the cell, 7 images each -- overflows it.  The overflow
    of a pipelined frame is only seen when the next frame of that handle is submitted (or at synchronize()): FramePipeline
    repeats the frame.  Results: those of synchronous calls."""
    import torch
    dtype = np.float64
    rng = np.random.default_rng(3)
    L, n, cutoff = 10.0, 6000, 1.0
    uniform = (L * rng.random((n, 3))).astype(dtype)
    corner = (0.8 * cutoff * rng.random((n, 3))).astype(dtype)      # every image index in {-1,0,1}^3 lands in the computing box
    frames = [uniform, corner, uniform + 0.01, corner + 0.01, corner + 0.02]
    uc = np.full(3, L, dtype)
    want = []
    hs = clm.Handle(3, dtype)
    hs.set_box(clm._capi.ORTHORHOMBIC, uc, cutoff, 1)
    for x in frames:
        hs.set_positions(0, x)
        e, f = np.zeros(1, dtype), np.zeros((n, 3), dtype)
        hs.map_lj(1.0e-3, 1.0e-6, e, f)
        want.append((e.copy(), f.copy()))
    hs.close()
    xs = [torch.from_numpy(x).pin_memory() for x in frames]
    es = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in frames]
    fs = [torch.zeros((n, 3), dtype=torch.float64).pin_memory() for _ in frames]
    pipe = clm.FramePipeline(3, dtype, uc, cutoff, handles=handles)
    for k in range(len(frames)):
        pipe.submit_lj(1.0e-3, 1.0e-6, xs[k].numpy(), es[k].numpy(), fs[k].numpy())
    pipe.synchronize()
    for k, (we, wf) in enumerate(want):
        assert abs(float(es[k][0]) - float(we[0])) <= 1e-12 * abs(float(we[0])), k
        assert np.abs(fs[k].numpy() - wf).max() <= 1e-12 * np.abs(wf).max(), k
    pipe.close()


# ---------------------------------------------------------------------------------------------------------
# non-periodic systems reuse the box of the previous build while the new coordinates stay inside the limits it was made
# from (_limits_fit_in_box, src/internals/ParticleSystem.jl:165-174; checked on the device, no host round trip) and get
# a new box when they do not: the neighbour list is the oracle's either way
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("two_sets", [False, True])
def test_nonperiodic_box_reuse_and_regrow(clm, oracle_mod, dtype, two_sets):
    rng = np.random.default_rng(77)
    x0 = rng.random((3000, 3)).astype(dtype)
    y0 = rng.random((2000, 3)).astype(dtype) if two_sets else None
    nb = clm.InPlaceNeighborList(x=x0, y=y0, cutoff=0.08)
    key = lambda lst: sorted(zip(lst["i"].tolist(), lst["j"].tolist(), lst["d"].tolist()))

    def check(x, y):
        lst = nb.neighborlist()
        i, j, d = oracle_mod.Oracle(x, 0.08, y=y, dtype=dtype).neighborlist()
        if two_sets:
            want = sorted(zip(i.tolist(), j.tolist(), d.tolist()))
            got = key(lst)
        else:
            want = sorted(zip(np.minimum(i, j).tolist(), np.maximum(i, j).tolist(), d.tolist()))
            got = sorted(zip(np.minimum(lst["i"], lst["j"]).tolist(), np.maximum(lst["i"], lst["j"]).tolist(), lst["d"].tolist()))
        assert got == want

    check(x0, y0)
    box0 = nb.sys._h.get_box()
    # (a) the coordinates contract towards the centre: they fit, the box is kept
    x1 = (0.5 + 0.8 * (x0 - 0.5)).astype(dtype)
    y1 = None if y0 is None else (0.5 + 0.8 * (y0 - 0.5)).astype(dtype)
    clm.update(nb, xpositions=x1, ypositions=y1)
    check(x1, y1)
    box1 = nb.sys._h.get_box()
    assert list(box1.nc) == list(box0.nc) and list(box1.origin) == list(box0.origin)
    # (b) they expand beyond the old limits: new limits, new box (one retry inside the call)
    x2 = (0.5 + 1.7 * (x0 - 0.5)).astype(dtype)
    y2 = None if y0 is None else (0.5 + 1.7 * (y0 - 0.5)).astype(dtype)
    clm.update(nb, xpositions=x2, ypositions=y2)
    check(x2, y2)
    box2 = nb.sys._h.get_box()
    assert list(box2.origin) != list(box0.origin)
    # (c) a single stray particle is enough
    x3 = x2.copy()
    x3[17] += dtype(5.0)
    clm.update(nb, xpositions=x3, ypositions=y2)
    check(x3, y2)

"""BASELINE.json configs 3 and 4 at FULL size on the GPU: the oracle where it finishes in seconds (C3), otherwise
size-independent properties of the domain (C4: independent-functor checksums, analytic pair statistics, exact
power-of-two linearity, Galilean invariance)."""
import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clm():
    import celllistmap_b200 as c
    return c


def test_c3_triclinic_cross_minimum_distance_full_size(clm, oracle_mod):
    """config 3: triclinic unit cell, 1M x 1M particles, cross-set minimum-distance map, Float64."""
    w = W.c3_triclinic_cross(1_000_000, 1_000_000)
    sys = clm.ParticleSystem(xpositions=w["x"], ypositions=w["y"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=clm.MinimumDistance())
    md = clm.pairwise(clm.MinimumDistanceMap(), sys)
    assert md.i >= 1 and md.j >= 1 and 0 < md.d < w["cutoff"]
    # the reported pair really is at that distance (minimum image recomputed on the host)
    v = clm.wrap_relative_to(w["y"][md.j - 1], w["x"][md.i - 1], w["unitcell"]) - w["x"][md.i - 1]
    assert abs(np.linalg.norm(v) - md.d) <= 1e-9
    # swapping the roles of the two sets gives the same pair
    sys2 = clm.ParticleSystem(xpositions=w["y"], ypositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=clm.MinimumDistance())
    md2 = clm.pairwise(clm.MinimumDistanceMap(), sys2)
    assert (md2.i, md2.j) == (md.j, md.i) and abs(md2.d - md.d) <= 1e-12
    # the CPU oracle on the same full-size input (all host threads)
    nt = oracle_mod.lib().ora_num_threads()
    oi, oj, od = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"], y=w["y"]).mindist(nbatches=nt)
    assert (md.i, md.j) == (oi, oj) and md.d == od


@pytest.mark.parametrize("dim", [3, 2])
def test_c4_pairwise_velocities_full_size(clm, dim):
    """config 4: halotools-style mean pairwise velocity, 4M galaxies, 3-D and 2-D."""
    w = W.c4_galaxies(4_000_000, dim)
    n, L, rb = w["x"].shape[0], w["L"], w["rbins"]
    out = (np.zeros(5, np.int64), np.zeros(5, np.float64))
    sys = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=out)
    counts, sums = clm.pairwise(clm.PairwiseVelocities(rb, w["v"]), sys)
    counts, sums = counts.copy(), sums.copy()
    # checksum against an independent functor: total count == number of in-cutoff pairs (all bins are inside the cutoff)
    sd, sd2, npairs = clm.pairwise(clm.SumDistances(), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=None))
    assert counts.sum() == npairs
    # analytic expectation for a uniform random field, bin by bin (Poisson noise ~ 1/sqrt(count) << 1e-3)
    shell = (4.0 / 3.0 * np.pi * (rb[1:] ** 3 - rb[:-1] ** 3)) if dim == 3 else np.pi * (rb[1:] ** 2 - rb[:-1] ** 2)
    expect = 0.5 * n * (n - 1) * shell / L ** dim
    assert np.all(np.abs(counts - expect) <= 5 * np.sqrt(expect) + 1e-4 * expect)
    # same histogram from the distance-histogram functor with unit bin width (bins right-closed vs left-closed differ only at integers)
    h = clm.pairwise(clm.DistanceHistogram(1.0), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=np.zeros(5, np.int64)))
    assert np.abs(h - counts).max() <= 2
    # linearity: scaling the velocities by 2 (exact in floating point) scales every sum exactly; counts unchanged
    out2 = (np.zeros(5, np.int64), np.zeros(5, np.float64))
    c2, s2 = clm.pairwise(clm.PairwiseVelocities(rb, 2.0 * w["v"]), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=out2))
    assert np.array_equal(c2, counts)
    assert np.abs(s2 - 2.0 * sums).max() <= 1e-9 * np.abs(sums).max()
    # Galilean invariance: a constant velocity offset cancels in v_i - v_j
    out3 = (np.zeros(5, np.int64), np.zeros(5, np.float64))
    c3, s3 = clm.pairwise(clm.PairwiseVelocities(rb, w["v"] + 0.5), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=out3))
    assert np.array_equal(c3, counts)
    assert np.abs(s3 - sums).max() <= 1e-6 * np.sqrt(float(counts.max()))
    # uncorrelated velocities: the mean pairwise velocity is consistent with zero
    assert np.all(np.abs(sums / counts) < 5.0 / np.sqrt(counts) + 1e-3)


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[1] (the config the metric is quoted on) at FULL size against the oracle, at the north_star
# tolerances: pair count equal, energy within 1e-5 (Float32) / 1e-10 (Float64), forces within the same bars against the
# oracle in the same precision; the distance to Float64 arithmetic is reported and bounded by the conditioning of
# Float32 coordinates (tests/parity_util.py).
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n3", [0, 1])
def test_c2_full_size_vs_oracle(clm, oracle_mod, dtype, n3):
    from parity_util import RTOL, conditioning_bound, force_report
    w = W.c2_argon(100, dtype)
    n = w["x"].shape[0]
    nt = oracle_mod.lib().ora_num_threads()
    tol = RTOL[np.dtype(dtype)]
    o64 = oracle_mod.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64))
    e64, f64 = o64.lj(w["c6"], w["c12"], forces=True, nbatches=nt)
    if dtype == np.float64:
        o_same, f_same = o64, f64
    else:
        o_same = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"], dtype=dtype)
        f_same = o_same.lj(w["c6"], w["c12"], forces=True, nbatches=nt)[1]
    sys = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"],
                             output=clm.EnergyAndForces(0.0, np.zeros((n, 3), dtype)))
    sys._h.set_option("n3", n3)   # 0: full-shell sweep (Float64 default), 1: Newton's-third-law sweep (Float32 default)
    out = clm.pairwise(clm.LJEnergyAndForces(w["c6"], w["c12"]), sys)
    # pair set: the exactly-once sweep sees the oracle's pairs (same precision: bit-identical coordinates)
    sd, sd2, npairs = clm.pairwise(clm.SumDistances(), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=None))
    assert npairs == o_same.sum_d_d2(nbatches=nt)[2]
    e_err = abs(out.energy - e64) / abs(e64)
    print(f"[parity] C2 1M {np.dtype(dtype).name} n3={n3}: |E - E_oracle64| / |E| = {e_err:.3e}")
    assert e_err <= tol
    err_same, err_64, _ = force_report(f"C2 1M {np.dtype(dtype).name} n3={n3}", out.forces, f_same, f64)
    if dtype == np.float32 and n3 == 0:
        # the full-shell sweep evaluates a pair that crosses the periodic boundary twice, from two different image pairs
        # (i real / j image and j real / i image), whose Float32 coordinates round differently; the reference (and the
        # Newton's-third-law sweep, the Float32 default) evaluates it once.  That path only meets the conditioning bound.
        assert err_same <= conditioning_bound(dtype, w["L"], 0.5 * W.ARGON_RHO ** (-1.0 / 3.0), 12)
    else:
        assert err_same <= tol
    a = W.ARGON_RHO ** (-1.0 / 3.0)
    assert err_64 <= max(tol, conditioning_bound(dtype, w["L"], 0.5 * a, 12))
    # energy-only map (the reference's exactly-once sweep, unfused arithmetic)
    e_once = clm.pairwise(clm.LJEnergy(w["c6"], w["c12"]), clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=0.0))
    assert abs(e_once - e64) <= tol * abs(e64)


@pytest.mark.parametrize("dim,n", [(3, 1_000_000), (2, 4_000_000)])
def test_c4_pairwise_velocities_vs_oracle(clm, oracle_mod, dim, n):
    """config 4 against the oracle where it finishes in seconds: 3-D at 1M galaxies (1.96e8 pairs), 2-D at the full 4M"""
    w = W.c4_galaxies(n, dim)
    nt = oracle_mod.lib().ora_num_threads()
    out = (np.zeros(5, np.int64), np.zeros(5, np.float64))
    sys = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=out)
    counts, sums = clm.pairwise(clm.PairwiseVelocities(w["rbins"], w["v"]), sys)
    wc, ws = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"]).pairvel(w["v"], w["rbins"], nbatches=nt)
    assert np.array_equal(counts, wc)
    # sums of signed terms (the mean pairwise velocity of an uncorrelated field is ~0): the scale of a bin is the sum of
    # |terms| ~ count * <|dv . r|/r> ~ 0.3 * count
    scale = 0.3 * counts.astype(np.float64)
    err = np.abs(sums - ws) / scale
    print(f"[parity] C4 {dim}-D {n}: max |sum - oracle| / (0.3 count) = {err.max():.3e}")
    assert err.max() <= 1e-10


def test_c5_single_gpu_vs_oracle(clm, oracle_mod):
    """config 5's particle system (64M in the multi-GPU bench) at the size the oracle finishes in seconds: 8M particles,
    Float32, one GPU: energy and forces against the oracle"""
    from parity_util import force_report
    w = W.c2_argon(200, np.float32)
    n = w["x"].shape[0]
    nt = oracle_mod.lib().ora_num_threads()
    sys = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"],
                             output=clm.EnergyAndForces(0.0, np.zeros((n, 3), np.float32)))
    out = clm.pairwise(clm.LJEnergyAndForces(w["c6"], w["c12"]), sys)
    o64 = oracle_mod.Oracle(w["x"].astype(np.float64), w["cutoff"], unitcell=w["unitcell"].astype(np.float64))
    e64, f64 = o64.lj(w["c6"], w["c12"], forces=True, nbatches=nt)
    del o64
    f32 = oracle_mod.Oracle(w["x"], w["cutoff"], unitcell=w["unitcell"], dtype=np.float32).lj(w["c6"], w["c12"], forces=True, nbatches=nt)[1]
    e_err = abs(out.energy - e64) / abs(e64)
    print(f"[parity] C5-like 8M f32: |E - E_oracle64| / |E| = {e_err:.3e}")
    assert e_err <= 1e-5
    err_same, _, _ = force_report("C5-like 8M f32", out.forces, f32, f64)
    assert err_same <= 1e-5

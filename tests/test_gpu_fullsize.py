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

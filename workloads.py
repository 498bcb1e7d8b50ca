"""Bit-reproducible synthetic inputs of the BASELINE.json configs.

RNG: splitmix64 stream, u = (next() >> 11) * 2^-53, seed 321 (+k for sub-streams); Float32 inputs are the
Float64 values rounded once.  Language neutral: a Julia/C++ twin only needs the 64-bit integer mix below.
Used by tests/ and bench.py; never by the product path.
"""
import numpy as np

SEED = 321
_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed, n):
    """first n outputs of the splitmix64 stream started at `seed` (uint64 array)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + _GAMMA * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed, shape):
    n = int(np.prod(shape))
    return ((splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53).reshape(shape)


def shuffle_perm(seed, n):
    """Fisher-Yates order drawn from the stream; implemented as argsort of the stream (a permutation drawn from the
    same seed -- ties have probability ~n^2/2^64)."""
    return np.argsort(splitmix64(seed, n), kind="stable")


# LJ parameters of argon in the reference's GROMACS comparison (test/applications/gromacs/compare_with_gromacs.jl:10-11)
ARGON_C6 = 0.00622127e6
ARGON_C12 = 9.69576e6
ARGON_RHO = 0.0213          # atoms / A^3 (liquid argon)


def c1_neighborlist(n=10_000, dtype=np.float64):
    """C1: n uniform points in the unit cube, unitcell [1,1,1], cutoff 0.1 (src/API/neighborlist.jl:52-67)."""
    x = uniform(SEED, (n, 3)).astype(dtype)
    return dict(x=x, unitcell=np.ones(3, dtype), cutoff=0.1)


def c2_argon(nside=100, dtype=np.float32, cutoff=12.0):
    """C2/C5: nside^3 simple-cubic sites at liquid-argon density, each jittered by uniform(-a/4, a/4) per axis,
    order shuffled; cubic PBC."""
    a = ARGON_RHO ** (-1.0 / 3.0)
    L = a * nside
    n = nside ** 3
    g = np.arange(nside, dtype=np.float64)
    sites = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(n, 3) * a
    jitter = (uniform(SEED, (n, 3)) - 0.5) * (0.5 * a)
    x = sites + jitter + 0.25 * a
    x = x[shuffle_perm(SEED + 1, n)]
    return dict(x=np.ascontiguousarray(x.astype(dtype)), unitcell=np.full(3, L, dtype), cutoff=cutoff, c6=ARGON_C6,
                c12=ARGON_C12, L=L)


def c3_triclinic_cross(nx=1_000_000, ny=1_000_000, dtype=np.float64, cutoff=12.0):
    """C3: triclinic cell s*[80 0 30; 30 80 0; 0 40 80] (compare_with_namd.jl:119-121) scaled to argon density,
    two sets with fractional coordinates u^3 mapped through the cell matrix."""
    M0 = np.array([[80.0, 0.0, 30.0], [30.0, 80.0, 0.0], [0.0, 40.0, 80.0]])
    vol = (nx + ny) / ARGON_RHO
    s = (vol / abs(np.linalg.det(M0))) ** (1.0 / 3.0)
    M = s * M0
    x = uniform(SEED, (nx, 3)) @ M.T
    y = uniform(SEED + 2, (ny, 3)) @ M.T
    return dict(x=np.ascontiguousarray(x.astype(dtype)), y=np.ascontiguousarray(y.astype(dtype)), unitcell=M.astype(dtype),
                cutoff=cutoff)


def triclinic_argon(nside=18, dtype=np.float64, cutoff=12.0):
    """nside^3 jittered lattice sites (fractional coordinates) in the triclinic cell of C3 scaled to argon density: a
    well-conditioned triclinic self-set system (no close pairs) for force parity, order shuffled."""
    M0 = np.array([[80.0, 0.0, 30.0], [30.0, 80.0, 0.0], [0.0, 40.0, 80.0]])
    n = nside ** 3
    s = ((n / ARGON_RHO) / abs(np.linalg.det(M0))) ** (1.0 / 3.0)
    M = s * M0
    g = np.arange(nside, dtype=np.float64)
    frac = (np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(n, 3) + 0.4 + 0.2 * uniform(SEED + 5, (n, 3))) / nside
    x = (frac @ M.T)[shuffle_perm(SEED + 6, n)]
    return dict(x=np.ascontiguousarray(x.astype(dtype)), unitcell=M.astype(dtype), cutoff=cutoff, c6=ARGON_C6, c12=ARGON_C12)


def c4_galaxies(n=4_000_000, dim=3, dtype=np.float64):
    """C4: halotools-style pair-velocity input (test/examples/pairwise_velocities.jl:38-49): density 10^5/20.274^3
    per Mpc^3 (3-D) or its 2/3 power (2-D), velocities uniform in [0,1), r-bins 0..5, cutoff 5."""
    rho3 = 1.0e5 / 20.274 ** 3
    rho = rho3 if dim == 3 else rho3 ** (2.0 / 3.0)
    L = (n / rho) ** (1.0 / dim)
    x = L * uniform(SEED, (n, dim))
    v = uniform(SEED + 1, (n, dim))
    return dict(x=np.ascontiguousarray(x.astype(dtype)), v=np.ascontiguousarray(v.astype(dtype)),
                unitcell=np.full(dim, L, dtype), cutoff=5.0, rbins=np.arange(6, dtype=dtype), L=L)

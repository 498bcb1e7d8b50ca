"""User pair functions compiled at run time (clm_custom_compile / clm_map_custom): the GPU
counterpart of passing an arbitrary closure to pairwise!(f, sys).  Checked against numpy evaluations of the same
function over the ORACLE's neighbour list (pair set and distances bit-identical to the reference's algorithm) and
against the compiled-in catalogue.  Tolerances: 1e-10 (Float64) / 1e-5 (Float32) relative, counts exact."""
import numpy as np
import pytest

from test_gpu_parity import random_system

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float64): 1e-10, np.dtype(np.float32): 1e-5}


@pytest.fixture(scope="module")
def clm():
    import celllistmap_b200 as c
    return c


SCALAR_SRC = """
struct InvDist {   // sum 1/d, sum w_i w_j d2, pair count
    static constexpr int NSCALAR = 3, NPART = 0, NAUX = 1, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        out.add_scalar(0, par[0] / p.d());
        out.add_scalar(1, p.ai[0] * p.aj[0] * p.d2);
        out.add_scalar(2, T(1));
    }
};
"""

PART_SRC = """
struct Gravity {   // per particle: (neighbour count, sum_j w_i w_j (x_j - x_i)/d^3); scalar: sum w_i w_j / d
    static constexpr int NSCALAR = 1, NPART = 4, NAUX = 1, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        const T d = p.d();
        const T g = par[0] * p.ai[0] * p.aj[0] / (d * p.d2);
        out.add_scalar(0, par[0] * p.ai[0] * p.aj[0] / d);
        out.add_i(0, T(1));
        for (int k = 0; k < 3; ++k) out.add_i(k + 1, g * (p.y[k] - p.x[k]));
    }
};
"""

HIST_SRC = """
struct RadialHist {   // counts[floor(d/width)] += 1, sums[...] += d; scalar: number of pairs with i < j (all of them, once)
    static constexpr int NSCALAR = 1, NPART = 0, NAUX = 0, HIST = 1;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        out.add_hist((int)floor(p.d() / par[0]), p.d());
        out.add_scalar(0, T(1));
    }
};
"""

LJ_SRC = """
struct MyLJ {   // c12/d2^6 - c6/d2^3, par = {c6, c12}
    static constexpr int NSCALAR = 1, NPART = 0, NAUX = 0, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        const T r6 = T(1) / (p.d2 * p.d2 * p.d2);
        out.add_scalar(0, r6 * (par[1] * r6 - par[0]));
    }
};
"""


def min_image(x, y, i, j, uc):
    """minimum-image vectors y[j] - x[i] (float64) for the 0-based pair arrays i, j"""
    v = y[j].astype(np.float64) - x[i].astype(np.float64)
    if uc is None:
        return v
    M = np.diag(uc.astype(np.float64)) if uc.ndim == 1 else uc.astype(np.float64)
    f = np.linalg.solve(M, v.T)
    f -= np.round(f)
    return (M @ f).T


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("two", [False, True])
def test_custom_scalars(clm, oracle_mod, dtype, kind, two):
    rng = np.random.default_rng(21)
    x, uc = random_system(rng, 3000, 3, kind, dtype)
    y = random_system(rng, 2500, 3, kind, dtype)[0] if two else None
    wx = (0.5 + rng.random(len(x))).astype(dtype)
    wy = (0.5 + rng.random(len(y))).astype(dtype) if two else None
    sys = clm.ParticleSystem(xpositions=x, ypositions=y, unitcell=uc, cutoff=1.3, output=clm.CustomOutput(scalars=np.zeros(3, dtype)))
    f = clm.CustomPairFunction(SCALAR_SRC, "InvDist", params=(2.0,), aux=wx, aux_y=wy)
    out = clm.pairwise(f, sys)
    assert f.info(sys) == {"nscalar": 3, "npart": 0, "naux": 1, "hist": 0}
    i, j, d = oracle_mod.Oracle(x, 1.3, unitcell=uc, y=y, dtype=dtype).neighborlist()
    d = d.astype(np.float64)
    wj = (wy if two else wx).astype(np.float64)[j - 1]
    want = np.array([(2.0 / d).sum(), (wx.astype(np.float64)[i - 1] * wj * d * d).sum(), float(len(d))])
    tol = RTOL[np.dtype(dtype)]
    assert out.scalars[2] == want[2]
    assert np.all(np.abs(out.scalars[:2] - want[:2]) <= tol * np.abs(want[:2])), (out.scalars, want)
    # reset = false accumulates on the values found in the output (API/pairwise.jl:52-54)
    out2 = clm.pairwise(f, sys, reset=False)
    assert np.all(np.abs(out2.scalars - 2 * want) <= 2 * tol * np.abs(want))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("two", [False, True])
def test_custom_per_particle(clm, oracle_mod, dtype, kind, two):
    rng = np.random.default_rng(33)
    x, uc = random_system(rng, 2500, 3, kind, dtype)
    y = random_system(rng, 2000, 3, kind, dtype)[0] if two else None
    wx = (0.5 + rng.random(len(x))).astype(dtype)
    wy = (0.5 + rng.random(len(y))).astype(dtype) if two else None
    n = len(x)
    sys = clm.ParticleSystem(xpositions=x, ypositions=y, unitcell=uc, cutoff=1.25,
                             output=clm.CustomOutput(scalars=np.zeros(1, dtype), per_particle=np.zeros((n, 4), dtype)))
    f = clm.CustomPairFunction(PART_SRC, "Gravity", params=(-9.8,), aux=wx, aux_y=wy)
    out = clm.pairwise(f, sys)
    i, j, d = oracle_mod.Oracle(x, 1.25, unitcell=uc, y=y, dtype=dtype).neighborlist()
    i, j, d = i - 1, j - 1, d.astype(np.float64)
    yy = y if two else x
    v = min_image(x, yy, i, j, uc)
    ww = -9.8 * wx.astype(np.float64)[i] * (wy if two else wx).astype(np.float64)[j]
    g = (ww / d ** 3)[:, None] * v
    want = np.zeros((n, 4))
    np.add.at(want[:, 0], i, 1.0)
    np.add.at(want[:, 1:], i, g)
    if not two:   # self sets: the pair acts on both particles (f[i] += df; f[j] -= df)
        np.add.at(want[:, 0], j, 1.0)
        np.add.at(want[:, 1:], j, -g)
    tol = RTOL[np.dtype(dtype)]
    assert np.array_equal(out.per_particle[:, 0], want[:, 0].astype(dtype))
    scale = np.abs(want[:, 1:]).max()
    # forces are sums of terms of mixed sign: absolute error relative to the largest component.  `want` is Float64
    # arithmetic on the same inputs: for Float32 inputs the distance to it is bounded by the conditioning of the wrapped
    # coordinates (a d^-2 force at the closest pair: 3 * 4 ulp(coordinate) / d_min, tests/parity_util.py)
    from parity_util import conditioning_bound
    err = np.abs(out.per_particle[:, 1:] - want[:, 1:]).max() / scale
    bound = max(tol, conditioning_bound(dtype, 13.0, float(d.min()), 2))
    print(f"[parity] custom gravity {kind} two={two} {np.dtype(dtype).name}: max|F - F64| / max|F| = {err:.3e} (bound {bound:.3e})")
    assert err <= bound
    e = (ww / d).sum()
    assert abs(out.scalars[0] - e) <= tol * abs(e)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic"])
@pytest.mark.parametrize("nbins", [8, 100])
def test_custom_histogram(clm, oracle_mod, dtype, kind, nbins):
    rng = np.random.default_rng(44)
    x, uc = random_system(rng, 3000, 3, kind, dtype)
    cutoff = 1.3
    width = dtype(cutoff / nbins * 1.07)    # a few trailing bins stay empty; nothing falls outside
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=cutoff,
                             output=clm.CustomOutput(scalars=np.zeros(1, dtype), hist_counts=np.zeros(nbins, np.int64), hist_sums=np.zeros(nbins, dtype)))
    out = clm.pairwise(clm.CustomPairFunction(HIST_SRC, "RadialHist", params=(width,)), sys)
    i, j, d = oracle_mod.Oracle(x, cutoff, unitcell=uc, dtype=dtype).neighborlist()
    b = np.floor(d / width).astype(np.int64)    # same T arithmetic as the device (IEEE division, floor)
    keep = b < nbins
    want_c = np.bincount(b[keep], minlength=nbins)
    want_s = np.bincount(b[keep], weights=d[keep].astype(np.float64), minlength=nbins)
    assert np.array_equal(out.hist_counts, want_c)
    assert out.scalars[0] == len(d)
    tol = RTOL[np.dtype(dtype)]
    assert np.all(np.abs(out.hist_sums - want_s) <= tol * np.maximum(want_s, 1.0))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_custom_matches_catalogue_lj(clm, dtype):
    import workloads as W
    w = W.c2_argon(16, dtype, cutoff=12.0)
    sys = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=clm.CustomOutput(scalars=np.zeros(1, dtype)))
    e_custom = clm.pairwise(clm.CustomPairFunction(LJ_SRC, "MyLJ", params=(w["c6"], w["c12"])), sys).scalars[0]
    sys2 = clm.ParticleSystem(xpositions=w["x"], unitcell=w["unitcell"], cutoff=w["cutoff"], output=0.0)
    e_cat = clm.pairwise(clm.LJEnergy(w["c6"], w["c12"]), sys2)
    assert abs(e_custom - e_cat) <= 4 * RTOL[np.dtype(dtype)] * abs(e_cat)


def test_custom_2d_and_update(clm, oracle_mod):
    """2-D system; the compiled functor is reused after update!(sys; positions, cutoff)"""
    rng = np.random.default_rng(5)
    x, uc = random_system(rng, 2000, 2, "ortho", np.float64)
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=1.0, output=clm.CustomOutput(scalars=np.zeros(3)))
    f = clm.CustomPairFunction(SCALAR_SRC, "InvDist", params=(1.0,), aux=np.ones(len(x)))
    for cutoff in (1.0, 1.4):
        clm.update(sys, cutoff=cutoff)
        out = clm.pairwise(f, sys)
        d = oracle_mod.Oracle(x, cutoff, unitcell=uc).neighborlist()[2]
        assert out.scalars[2] == len(d)
        assert abs(out.scalars[0] - (1.0 / d).sum()) <= 1e-10 * (1.0 / d).sum()
        assert abs(out.scalars[1] - (d * d).sum()) <= 1e-10 * (d * d).sum()


def test_custom_errors(clm):
    rng = np.random.default_rng(6)
    x, uc = random_system(rng, 500, 3, "ortho", np.float64)
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=1.0, output=clm.CustomOutput(scalars=np.zeros(3)))
    with pytest.raises(ValueError, match="NVRTC compilation of the pair function failed"):
        clm.pairwise(clm.CustomPairFunction(SCALAR_SRC.replace("p.d2", "p.no_such_field"), "InvDist", aux=np.ones(len(x))), sys)
    with pytest.raises(ValueError, match="identifier"):
        clm.pairwise(clm.CustomPairFunction(SCALAR_SRC, "Inv Dist", aux=np.ones(len(x))), sys)
    with pytest.raises(ValueError, match="aux is required"):
        clm.pairwise(clm.CustomPairFunction(SCALAR_SRC, "InvDist"), sys)
    sys.output = clm.CustomOutput(scalars=np.zeros(2))
    with pytest.raises(ValueError, match="scalars"):
        clm.pairwise(clm.CustomPairFunction(SCALAR_SRC, "InvDist", aux=np.ones(len(x))), sys)
    with pytest.raises(TypeError):
        clm.pairwise(lambda pair, out: out, sys)


# ---------------------------------------------------------------------------------------------------------
# custom reducers (src/API/parallel_custom.jl:196-214): scalar outputs reduced with min / max instead of +
MINMAX_SRC = """
struct Extremes {   // 0: smallest d2, 1: largest w_i w_j d, 2: pair count (a + output next to the two reducers)
    static constexpr int NSCALAR = 3, NPART = 0, NAUX = 1, HIST = 0;
    static constexpr unsigned SCALAR_MIN = 1u, SCALAR_MAX = 2u;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        out.min_scalar(0, p.d2);
        out.max_scalar(1, p.ai[0] * p.aj[0] * p.d());
        out.add_scalar(2, T(1));
    }
};
"""
MINMAX_PART_SRC = """
struct NearestAndCount {   // per particle: neighbour count; scalar 0: smallest d over all pairs (full-shell mode)
    static constexpr int NSCALAR = 1, NPART = 1, NAUX = 0, HIST = 0;
    static constexpr unsigned SCALAR_MIN = 1u;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        out.min_scalar(0, p.d());
        out.add_i(0, T(1));
    }
};
"""


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ortho", "triclinic", "nonperiodic"])
@pytest.mark.parametrize("two", [False, True])
def test_custom_min_max_reducers(clm, oracle_mod, dtype, kind, two):
    rng = np.random.default_rng(5)
    x, uc = random_system(rng, 3000, 3, kind, dtype)
    y = random_system(rng, 2500, 3, kind, dtype)[0] if two else None
    wx = (0.5 + rng.random(len(x))).astype(dtype)
    wy = (0.5 + rng.random(len(y))).astype(dtype) if two else None
    sys = clm.ParticleSystem(xpositions=x, ypositions=y, unitcell=uc, cutoff=1.3, output=clm.CustomOutput(scalars=np.zeros(3, dtype)))
    f = clm.CustomPairFunction(MINMAX_SRC, "Extremes", aux=wx, aux_y=wy)
    out = clm.pairwise(f, sys)
    i, j, d = oracle_mod.Oracle(x, 1.3, unitcell=uc, y=y, dtype=dtype).neighborlist()
    wj = (wy if two else wx)[j - 1]
    # the minimum of the exact d2 values is exact; d2 = d*d up to the square root's rounding
    assert abs(float(out.scalars[0]) - float(d.min()) ** 2) <= 4 * np.finfo(dtype).eps * float(d.min()) ** 2
    want_max = (wx[i - 1].astype(np.float64) * wj.astype(np.float64) * d.astype(np.float64)).max()
    assert abs(float(out.scalars[1]) - want_max) <= RTOL[np.dtype(dtype)] * want_max
    assert out.scalars[2] == len(d)
    # reset = false: the values found in the output take part in the reduction (min / max), the + output accumulates
    sys.output.scalars[:] = [dtype(1e-9), dtype(1e9), dtype(5)]
    out2 = clm.pairwise(f, sys, reset=False)
    assert out2.scalars[0] == dtype(1e-9) and out2.scalars[1] == dtype(1e9) and out2.scalars[2] == len(d) + 5
    # no pair at all: the identities
    far = (x[:1] + dtype(3.0)).astype(dtype)
    if two:
        sys0 = clm.ParticleSystem(xpositions=x[:1], ypositions=far, unitcell=uc, cutoff=1.3, output=clm.CustomOutput(scalars=np.zeros(3, dtype)))
        f0 = clm.CustomPairFunction(MINMAX_SRC, "Extremes", aux=wx[:1], aux_y=wx[:1])
    else:
        sys0 = clm.ParticleSystem(xpositions=np.concatenate([x[:1], far]), unitcell=uc, cutoff=1.3, output=clm.CustomOutput(scalars=np.zeros(3, dtype)))
        f0 = clm.CustomPairFunction(MINMAX_SRC, "Extremes", aux=wx[:2])
    o0 = clm.pairwise(f0, sys0)
    assert o0.scalars[0] == np.inf and o0.scalars[1] == -np.inf and o0.scalars[2] == 0


def test_custom_min_reducer_with_per_particle_output(clm, oracle_mod):
    rng = np.random.default_rng(6)
    x, uc = random_system(rng, 2500, 3, "ortho", np.float64)
    n = len(x)
    sys = clm.ParticleSystem(xpositions=x, unitcell=uc, cutoff=1.2, output=clm.CustomOutput(scalars=np.zeros(1), per_particle=np.zeros((n, 1))))
    out = clm.pairwise(clm.CustomPairFunction(MINMAX_PART_SRC, "NearestAndCount"), sys)
    i, j, d = oracle_mod.Oracle(x, 1.2, unitcell=uc).neighborlist()
    assert out.scalars[0] == d.min()          # not halved: only + outputs of the full-shell mode are
    cnt = np.bincount(np.concatenate([i, j]) - 1, minlength=n)
    assert np.array_equal(out.per_particle[:, 0], cnt)

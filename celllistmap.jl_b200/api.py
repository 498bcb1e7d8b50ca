"""Host-side mirror of the reference's operator interface for the cutoff-pair path.

Same names, argument meaning and error behaviour as CellListMap.jl 0.10 (`!` dropped, as Python
identifiers cannot carry it):

    ParticleSystem(...)          src/API/ParticleSystem.jl:142-202
    pairwise(f, sys; reset)      pairwise!   src/API/pairwise.jl:48-63
    update(sys; ...)             update!     src/API/updating.jl:165-187
    resize_output(sys, n)        resize_output!  src/API/updating.jl:18-25
    neighborlist(...)            src/API/neighborlist.jl:314-340
    InPlaceNeighborList(...)     src/API/neighborlist.jl:84-111 ; update (:159-169) ; neighborlist_ (neighborlist!, :217-231)
    get_computing_box(sys)       src/API/get_computing_box.jl:11
    wrap_relative_to             src/internals/CellOperations.jl:102-127

`f` is NOT an arbitrary closure (SURVEY.md §2 row 13 is out of scope): it is one of the compiled-in catalogue
functors below, each the device twin of a pair function the reference's tests/docs define.
All compute happens in libclm_b200.so on the GPU; this module only moves arguments across the C ABI.
Particle indices are 1-based, as in the reference.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import ClmError, Handle, nl_dtype

__all__ = [
    "ParticleSystem", "ParticleSystemPositions", "pairwise", "update", "resize_output", "neighborlist",
    "InPlaceNeighborList", "neighborlist_", "get_computing_box", "Box", "NeighborPair", "DimensionMismatch",
    "LJEnergy", "LJForces", "LJEnergyAndForces", "CoulombEnergy", "CoulombForces", "CoulombEnergyAndForces",
    "DistanceHistogram", "PairwiseVelocities", "MinimumDistanceMap", "SumDistances", "MinimumDistance",
    "EnergyAndForces", "wrap_relative_to", "CustomPairFunction", "CustomOutput", "FramePipeline",
]


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (src/internals/CellLists.jl:740-751)."""


def _raise(e):
    """clm_status -> the exception class the reference throws for the same condition."""
    if e.code in (1, 2, 3):
        raise ValueError(e.message) from None          # ArgumentError
    if e.code == 5:
        raise DimensionMismatch(e.message) from None
    raise RuntimeError(str(e)) from None               # ErrorException


# ------------------------------------------------------------------------------------------------------
# NeighborPair (src/API/NeighborPair.jl:19-33): what a pair function sees.  The catalogue functors consume
# it on the device; this record documents the fields and is what `neighborlist` rows unpack to.
class NeighborPair:
    __slots__ = ("i", "j", "x", "y", "d2")

    def __init__(self, i, j, x, y, d2):
        self.i, self.j, self.x, self.y, self.d2 = i, j, x, y, d2

    @property
    def d(self):
        return float(np.sqrt(self.d2))


class MinimumDistance:
    """(i, j, d) output of the minimum-distance map (docs/src/ParticleSystem/examples.md:117-132)."""
    __slots__ = ("i", "j", "d")

    def __init__(self, i=0, j=0, d=np.inf):
        self.i, self.j, self.d = int(i), int(j), d

    def __repr__(self):
        return f"MinimumDistance(i={self.i}, j={self.j}, d={self.d})"


class EnergyAndForces:
    """compound output (docs/src/ParticleSystem/examples.md:64-89)."""
    __slots__ = ("energy", "forces")

    def __init__(self, energy, forces):
        self.energy, self.forces = energy, forces


# ------------------------------------------------------------------------------------------------------
# functor catalogue.  Each functor knows how to run itself through the C ABI and how to fold the result
# into `sys.output` with the reference's reset / accumulate semantics (API/pairwise.jl:52-54).
class _Functor:
    def run(self, sys, reset):
        raise NotImplementedError


class SumDistances(_Functor):
    """test functor f1/f2 (test/modules/Testing.jl:23-26): output = (sum d, sum d2, number of pairs)."""

    def run(self, sys, reset):
        h, T = sys._h, sys.dtype
        sd, sd2, n = np.zeros(1, T), np.zeros(1, T), np.zeros(1, np.int64)
        if not reset and sys.output is not None:
            sd[0], sd2[0], n[0] = sys.output
        h.map_sum_d_d2(sd, sd2, n, reset=reset)
        return (sd[0], sd2[0], int(n[0]))


class LJEnergy(_Functor):
    """u += c12/d2^6 - c6/d2^3 (test/applications/gromacs/compare_with_gromacs.jl:9-13); output: scalar."""

    def __init__(self, c6, c12):
        self.c6, self.c12 = c6, c12

    @classmethod
    def from_eps_sigma(cls, eps, sigma):
        """eps*((sig/d)^12 - 2 (sig/d)^6) (test/applications/namd/compare_with_namd.jl:7-13)."""
        return cls(2.0 * eps * sigma ** 6, eps * sigma ** 12)

    def run(self, sys, reset):
        e = np.zeros(1, sys.dtype)
        if not reset:
            e[0] = sys.output
        sys._h.map_lj(self.c6, self.c12, e, None, reset=reset)
        return e[0]


class LJForces(LJEnergy):
    """f[i] += df; f[j] -= df (docs/src/ParticleSystem/examples.md:41-47); output: (n, N) array, updated in place."""

    def run(self, sys, reset):
        f = _check_force_output(sys, sys.output)
        e = np.zeros(1, sys.dtype)
        sys._h.map_lj(self.c6, self.c12, e, f, reset=reset)
        return f


class LJEnergyAndForces(LJEnergy):
    def run(self, sys, reset):
        out = sys.output
        f = _check_force_output(sys, out.forces)
        e = np.array([0 if reset else out.energy], sys.dtype)
        sys._h.map_lj(self.c6, self.c12, e, f, reset=reset)
        out.energy = e[0]
        return out


class CoulombEnergy(_Functor):
    """u += k*w_i*w_j/d (test/examples/gravitational_potential.jl:30-34)."""

    def __init__(self, k, weights, weights_y=None):
        self.k, self.wx, self.wy = k, weights, weights_y

    def _w(self, sys):
        wx = np.ascontiguousarray(self.wx, dtype=sys.dtype)
        if wx.shape[0] != len(sys.xpositions):
            raise DimensionMismatch("weights must have one entry per particle")
        wy = None
        if sys.ypositions is not None:
            if self.wy is None:
                raise ValueError("weights_y is required for a two-set system")
            wy = np.ascontiguousarray(self.wy, dtype=sys.dtype)
            if wy.shape[0] != len(sys.ypositions):
                raise DimensionMismatch("weights_y must have one entry per particle of the second set")
        return wx, wy

    def run(self, sys, reset):
        wx, wy = self._w(sys)
        e = np.zeros(1, sys.dtype)
        if not reset:
            e[0] = sys.output
        sys._h.map_coulomb(self.k, wx, wy, e, None, reset=reset)
        return e[0]


class CoulombForces(CoulombEnergy):
    """F_i += k w_i w_j (x_i - x_j)/d^3 (test/examples/gravitational_force.jl:38-44)."""

    def run(self, sys, reset):
        wx, wy = self._w(sys)
        f = _check_force_output(sys, sys.output)
        e = np.zeros(1, sys.dtype)
        sys._h.map_coulomb(self.k, wx, wy, e, f, reset=reset)
        return f


class CoulombEnergyAndForces(CoulombEnergy):
    def run(self, sys, reset):
        wx, wy = self._w(sys)
        out = sys.output
        f = _check_force_output(sys, out.forces)
        e = np.array([0 if reset else out.energy], sys.dtype)
        sys._h.map_coulomb(self.k, wx, wy, e, f, reset=reset)
        out.energy = e[0]
        return out


class DistanceHistogram(_Functor):
    """hist[floor(Int, d/width) + 1] += 1 (test/examples/distance_histogram.jl:22-26); output: int64 array."""

    def __init__(self, width=1.0):
        self.width = width

    def run(self, sys, reset):
        out = sys.output
        if not (isinstance(out, np.ndarray) and out.dtype == np.int64 and out.ndim == 1 and out.flags["C_CONTIGUOUS"]):
            raise TypeError("DistanceHistogram needs a contiguous 1-D int64 array as output")
        sys._h.map_dist_hist(self.width, out, reset=reset)
        return out


class PairwiseVelocities(_Functor):
    """halotools-style mean pairwise velocity (test/examples/pairwise_velocities.jl:17-24): output = (counts int64[nb],
    sums T[nb]) with nb = len(rbins) - 1; mean = sums / counts (:75-77)."""

    def __init__(self, rbins, velocities, velocities_y=None):
        self.rbins, self.vx, self.vy = np.asarray(rbins), velocities, velocities_y

    def run(self, sys, reset):
        counts, sums = sys.output
        nb = self.rbins.shape[0] - 1
        if counts.shape != (nb,) or sums.shape != (nb,) or counts.dtype != np.int64 or sums.dtype != sys.dtype:
            raise TypeError("PairwiseVelocities needs output = (int64[nbins], T[nbins])")
        vx = np.ascontiguousarray(self.vx, dtype=sys.dtype)
        if vx.shape != (len(sys.xpositions), sys.dim):
            raise DimensionMismatch("velocities must be (n, N)")
        vy = None
        if sys.ypositions is not None:
            if self.vy is None:
                raise ValueError("velocities_y is required for a two-set system")
            vy = np.ascontiguousarray(self.vy, dtype=sys.dtype)
            if vy.shape != (len(sys.ypositions), sys.dim):
                raise DimensionMismatch("velocities_y must be (n_y, N)")
        sys._h.map_pairvel(vx, vy, self.rbins, counts, sums, reset=reset)
        return sys.output


class MinimumDistanceMap(_Functor):
    """pair of smallest distance (test/examples/nearest_neighbor.jl:9-16, :43); output: MinimumDistance."""

    def run(self, sys, reset):
        i, j, d = np.zeros(1, np.int64), np.zeros(1, np.int64), np.full(1, np.inf, sys.dtype)
        if not reset and isinstance(sys.output, MinimumDistance):
            i[0], j[0], d[0] = sys.output.i, sys.output.j, sys.output.d
        sys._h.map_mindist(i, j, d, reset=reset)
        return MinimumDistance(i[0], j[0], d[0])


class CustomOutput:
    """compound output of a CustomPairFunction: `scalars` T[NSCALAR], `per_particle` (n, NPART) T, `hist_counts`
    int64[nbins], `hist_sums` T[nbins]; fields the functor does not declare stay None."""
    __slots__ = ("scalars", "per_particle", "hist_counts", "hist_sums")

    def __init__(self, scalars=None, per_particle=None, hist_counts=None, hist_sums=None):
        self.scalars, self.per_particle, self.hist_counts, self.hist_sums = scalars, per_particle, hist_counts, hist_sums


class CustomPairFunction(_Functor):
    """A user pair function: the GPU counterpart of passing an arbitrary closure to pairwise!(f, sys)
    (src/API/pairwise.jl:48-63).  `source` is CUDA C++ defining a stateless struct `name` (interface in
    include/clm_b200.h, csrc/clm_custom.cuh); it is compiled at run time with NVRTC into the same sweep kernel the
    catalogue uses.  `params` (<= 16 numbers) are handed to the functor as par[]; `aux` / `aux_y` are (n, NAUX)
    per-particle side arrays of the two sets.  Output: a CustomOutput with the default `+` reduction
    (src/API/parallel_custom.jl:53-54, :116-123, :213)."""

    def __init__(self, source, name, params=(), aux=None, aux_y=None):
        self.source, self.name, self.params, self.aux, self.aux_y = source, name, tuple(params), aux, aux_y

    def check(self, dtype=np.float64):
        """compile-only check (needs libnvrtc, no device); returns the NVRTC log."""
        try:
            return _capi.custom_check(self.source, self.name, dtype)
        except ClmError as e:
            _raise(e)

    def _get(self, sys):
        # compiled functors belong to the handle (one NVRTC compilation per handle, sweep mode and functor source)
        cache = sys._h.__dict__.setdefault("_custom_cache", {})
        key = (self.source, self.name)
        if key not in cache:
            cache[key] = sys._h.custom_compile(self.source, self.name)
        return cache[key]

    def info(self, sys):
        i = self._get(sys)[1]
        return {"nscalar": i.nscalar, "npart": i.npart, "naux": i.naux, "hist": i.hist}

    def run(self, sys, reset):
        fid, info = self._get(sys)
        out, T, n = sys.output, sys.dtype, len(sys.xpositions)
        if not isinstance(out, CustomOutput):
            raise TypeError("CustomPairFunction needs a CustomOutput as output")
        ok = lambda a, shape, dt: isinstance(a, np.ndarray) and a.dtype == dt and a.shape == shape and a.flags["C_CONTIGUOUS"]
        if info.nscalar and not ok(out.scalars, (info.nscalar,), T):
            raise DimensionMismatch(f"output.scalars must be a contiguous ({info.nscalar},) {T} array")
        if info.npart and not ok(out.per_particle, (n, info.npart), T):
            raise DimensionMismatch(f"output.per_particle must be a contiguous ({n}, {info.npart}) {T} array (resize it with resize_output)")
        if info.hist:
            if not (isinstance(out.hist_counts, np.ndarray) and out.hist_counts.ndim == 1 and ok(out.hist_counts, out.hist_counts.shape, np.int64)
                    and ok(out.hist_sums, out.hist_counts.shape, T)):
                raise DimensionMismatch("output.hist_counts / hist_sums must be contiguous int64[nbins] / T[nbins] arrays")
        ax = ay = None
        if info.naux:
            if self.aux is None:
                raise ValueError("the functor reads per-particle side arrays: aux is required")
            ax = np.ascontiguousarray(self.aux, dtype=T).reshape(-1, info.naux)
            if ax.shape[0] != n:
                raise DimensionMismatch("aux must have one row per particle")
            if sys.ypositions is not None:
                if self.aux_y is None:
                    raise ValueError("aux_y is required for a two-set system")
                ay = np.ascontiguousarray(self.aux_y, dtype=T).reshape(-1, info.naux)
                if ay.shape[0] != len(sys.ypositions):
                    raise DimensionMismatch("aux_y must have one row per particle of the second set")
        sys._h.map_custom(fid, self.params, ax, ay, out.scalars if info.nscalar else None, out.per_particle if info.npart else None,
                          out.hist_counts if info.hist else None, out.hist_sums if info.hist else None, reset=reset)
        return out


def _check_force_output(sys, f):
    n = len(sys.xpositions)
    if not (isinstance(f, np.ndarray) and f.dtype == sys.dtype and f.shape == (n, sys.dim) and f.flags["C_CONTIGUOUS"]):
        raise DimensionMismatch(f"force output must be a contiguous ({n}, {sys.dim}) {sys.dtype} array "
                                "(resize it with resize_output)")
    return f


# ------------------------------------------------------------------------------------------------------
class ParticleSystemPositions:
    """Owning copy of the coordinates plus the `updated` flag any mutation sets
    (src/API/ParticleSystemPositions.jl:14-101)."""

    def __init__(self, x, dim, dtype):
        self._a = _as_positions(x, dim, dtype).copy()
        self.updated = True

    def __len__(self):
        return self._a.shape[0]

    def __getitem__(self, k):
        return self._a[k]

    def __setitem__(self, k, v):
        self._a[k] = v
        self.updated = True

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    @property
    def array(self):
        """read-only view of the stored coordinates (mutations must go through __setitem__ / assign)."""
        v = self._a.view()
        v.flags.writeable = False
        return v

    def assign(self, new_x):
        """element-wise copy into the existing storage, resizing if needed (src/API/updating.jl:79-113)."""
        new_x = _as_positions(new_x, self._a.shape[1], self._a.dtype)
        if new_x.shape[0] != self._a.shape[0]:
            self._a = new_x.copy()
        else:
            self._a[...] = new_x
        self.updated = True

    def resize(self, n):
        a = np.zeros((n, self._a.shape[1]), self._a.dtype)
        m = min(n, self._a.shape[0])
        a[:m] = self._a[:m]
        self._a = a
        self.updated = True

    def append(self, rows):
        self._a = np.concatenate([self._a, _as_positions(rows, self._a.shape[1], self._a.dtype)], axis=0)
        self.updated = True

    def empty(self):
        self.resize(0)


def _get_dim(unitcell, x, y):
    """dimension inference (src/internals/auxiliary_functions.jl:8-37)."""
    if unitcell is not None:
        uc = np.asarray(unitcell)
        if uc.ndim == 2 and uc.shape[0] != uc.shape[1]:
            raise ValueError("Unit cell matrix must be square.")
        return int(uc.shape[0])
    for a in (x, y):
        if a is None:
            continue
        arr = np.asarray(a)
        if arr.ndim == 2 and arr.shape[0] > 0:
            # (n, N) rows -- or the reference's (N, n) matrix layout when the first extent is 2 or 3 and the second is not
            return int(arr.shape[1])
        if arr.ndim == 2 and arr.shape[1] in (2, 3):
            return int(arr.shape[1])
    raise ValueError("Could not infer dimension: provide a unitcell or non-empty (n, N) coordinates.")


def _as_positions(x, dim, dtype):
    a = np.asarray(x, dtype=dtype)
    if a.ndim == 1 and a.size == 0:
        a = a.reshape(0, dim)
    if a.ndim != 2 or a.shape[1] != dim:
        raise DimensionMismatch(f"Incompatible dimensions: coordinates must be (n, {dim}), got {a.shape}")
    return np.ascontiguousarray(a)


def _dtype_of(x, y):
    """the reference works in the element type of the coordinates; Float32 stays Float32, the rest is Float64."""
    ts = [np.asarray(a).dtype for a in (x, y) if a is not None]
    return np.dtype(np.float32) if ts and all(t == np.float32 for t in ts) else np.dtype(np.float64)


class Box:
    """read-only view of the Box record (src/internals/Box.jl:84-96) the library computed."""

    def __init__(self, info):
        n = info.dim
        m = lambda a: np.array(a[: n * n]).reshape(n, n).T.copy()  # column-major -> rows x cols
        self.dim, self.lcell = n, info.lcell
        self.unit_cell_type = ("OrthorhombicCell", "TriclinicCell", "NonPeriodicCell")[info.cell_type]
        self.input_unit_cell, self.aligned_unit_cell = m(info.input_unit_cell), m(info.aligned_unit_cell)
        self.rotation, self.inv_rotation = m(info.rotation), m(info.inv_rotation)
        self.nc = np.array(info.nc[:n], dtype=np.int64)
        self.cutoff, self.cutoff_sqr = info.cutoff, info.cutoff_sqr
        self.computing_box = (np.array(info.computing_box_min[:n]), np.array(info.computing_box_max[:n]))
        self.cell_size = np.array(info.cell_size[:n])
        self.origin = np.array(info.origin[:n])

    def __repr__(self):
        return (f"Box{{{self.unit_cell_type}, {self.dim}}}\n  unit cell matrix = {self.input_unit_cell.tolist()}\n"
                f"  cutoff = {self.cutoff}\n  number of computing cells on each dimension = {self.nc.tolist()}\n"
                f"  computing cell sizes = {self.cell_size.tolist()} (lcell: {self.lcell})\n"
                f"  Total number of cells = {int(np.prod(self.nc))}")


class ParticleSystem:
    """ParticleSystem1 / ParticleSystem2 (src/API/AbstractParticleSystem.jl:32-60) on one B200.

    Keyword arguments as in the reference (src/API/ParticleSystem.jl:178-202); `parallel`, `nbatches` are accepted
    and ignored (the GPU grid replaces the task batches); `device` selects the CUDA ordinal.
    """

    def __init__(self, *, positions=None, xpositions=None, ypositions=None, unitcell=None, cutoff, output,
                 output_name="default_output_name", parallel=True, nbatches=(0, 0), lcell=1, device=0):
        if (positions is None) == (xpositions is None):
            raise ValueError("Either `positions` OR `xpositions` must be defined.")
        x = positions if xpositions is None else xpositions
        self.dim = _get_dim(unitcell, x, ypositions)
        if self.dim not in (2, 3):
            raise DimensionMismatch("Dimension must be 2 or 3.")
        self.dtype = _dtype_of(x, ypositions)
        self.xpositions = ParticleSystemPositions(x, self.dim, self.dtype)
        self.ypositions = None if ypositions is None else ParticleSystemPositions(ypositions, self.dim, self.dtype)
        self.output = output
        self.output_name = output_name
        self.parallel = parallel
        self.nbatches = tuple(nbatches)
        self._lcell = int(lcell)
        self._cutoff = cutoff
        try:
            self._h = Handle(self.dim, self.dtype, device)
        except ClmError as e:
            _raise(e)
        if unitcell is None:
            self._cell_type, self._unitcell = _capi.NONPERIODIC, None
        else:
            uc = np.asarray(unitcell, dtype=self.dtype)
            self._cell_type = _capi.ORTHORHOMBIC if uc.ndim == 1 else _capi.TRICLINIC
            self._unitcell = uc.copy()
        self._box_dirty = True
        # the reference builds the cell list at construction (CellList(x, box), src/API/ParticleSystem.jl:157)
        self._sync()

    # -- properties of the reference (src/API/ParticleSystem.jl:204-221) --
    @property
    def positions(self):
        return self.xpositions

    @property
    def unitcell(self):
        if self._cell_type == _capi.NONPERIODIC:
            return self.box.input_unit_cell
        return self._unitcell if self._unitcell.ndim == 2 else np.diag(self._unitcell)

    @unitcell.setter
    def unitcell(self, uc):
        update(self, unitcell=uc)

    @property
    def cutoff(self):
        return self._cutoff

    @cutoff.setter
    def cutoff(self, rc):
        update(self, cutoff=rc)

    def __getattr__(self, name):
        if name != "output_name" and name == self.__dict__.get("output_name"):
            return self.__dict__["output"]
        raise AttributeError(name)

    @property
    def box(self):
        self._sync()
        return Box(self._h.get_box())

    def stats(self):
        return self._h.stats()

    # UpdateParticleSystem! (src/internals/ParticleSystem.jl:158-164, :209-224): rebuild only what changed
    def _sync(self):
        try:
            if self._box_dirty:
                self._h.set_box(self._cell_type, self._unitcell, self._cutoff, self._lcell)
                self._box_dirty = False
                self.xpositions.updated = True
            if self.xpositions.updated:
                self._h.set_positions(0, self.xpositions._a)
            if self.ypositions is not None and self.ypositions.updated:
                self._h.set_positions(1, self.ypositions._a)
            if self.xpositions.updated or (self.ypositions is not None and self.ypositions.updated):
                self._h.build()
                self.xpositions.updated = False
                if self.ypositions is not None:
                    self.ypositions.updated = False
        except ClmError as e:
            _raise(e)


def pairwise(f, sys, *, show_progress=False, reset=True):
    """pairwise!(f, sys; show_progress, reset) (src/API/pairwise.jl:48-63)."""
    if not isinstance(f, _Functor):
        raise TypeError("f must be one of the compiled-in catalogue functors or a CustomPairFunction (CUDA C++ source compiled "
                        "at run time); Python closures cannot run on the device (SURVEY.md §2 row 13)")
    sys._sync()
    try:
        sys.output = f.run(sys, reset)
    except ClmError as e:
        _raise(e)
    return sys.output


def update(sys, *, positions=None, xpositions=None, ypositions=None, cutoff=None, unitcell=None, parallel=None):
    """update!(sys; ...) (src/API/updating.jl:165-187): nothing is recomputed until the next pairwise call."""
    if isinstance(sys, InPlaceNeighborList):
        return sys.update(positions if xpositions is None else xpositions, ypositions, cutoff=cutoff, unitcell=unitcell,
                          parallel=parallel)
    if positions is not None and xpositions is not None:
        raise ValueError("Either `positions` OR `xpositions` must be provided, not both.")
    x = xpositions if positions is None else positions
    if ypositions is not None and sys.ypositions is None:
        raise ValueError("ypositions can only be set for a two-set particle system")
    if x is not None:
        sys.xpositions.assign(x)
    if ypositions is not None:
        sys.ypositions.assign(ypositions)
    if cutoff is not None:
        sys._cutoff = cutoff
        sys._box_dirty = True
    if unitcell is not None:
        if sys._cell_type == _capi.NONPERIODIC:
            raise ValueError("Manual updating of the unit cell of non-periodic systems is not allowed.")
        uc = np.asarray(unitcell, dtype=sys.dtype)
        if uc.shape[0] != sys.dim:
            raise DimensionMismatch("unit cell dimension does not match the system")
        sys._unitcell = uc.copy()   # the cell TYPE of the system is kept (update_box, src/internals/Box.jl:395-423)
        sys._box_dirty = True
    if parallel is not None:
        sys.parallel = parallel
    return sys


def resize_output(sys, n):
    """resize_output!(sys, n) (src/API/updating.jl:18-25): array outputs follow the particle count."""
    out = sys.output
    arr = out.forces if isinstance(out, EnergyAndForces) else (out.per_particle if isinstance(out, CustomOutput) else out)
    new = np.zeros((n,) + arr.shape[1:], arr.dtype)
    m = min(n, arr.shape[0])
    new[:m] = arr[:m]
    if isinstance(out, EnergyAndForces):
        out.forces = new
    elif isinstance(out, CustomOutput):
        out.per_particle = new
    else:
        sys.output = new
    return sys


def get_computing_box(sys):
    """(src/API/get_computing_box.jl:11)"""
    return sys.box.computing_box


def wrap_relative_to(x, xref, unitcell):
    """minimum image of x relative to xref (src/internals/CellOperations.jl:102-127); host helper, not on the hot path."""
    x, xref, uc = np.asarray(x, float), np.asarray(xref, float), np.asarray(unitcell, float)
    M = np.diag(uc) if uc.ndim == 1 else uc
    frac = lambda v: (lambda p: np.where(p - np.floor(p) == 1.0, 0.0, p - np.floor(p)))(np.linalg.solve(M, v))
    xf, rf = frac(x), frac(xref)
    w = np.mod(xf - rf, 1.0)
    w = np.where(w >= 0.5, w - 1.0, w)
    return M @ ((w + rf) - rf) + xref


# ------------------------------------------------------------------------------------------------------
class InPlaceNeighborList:
    """InPlaceNeighborList (src/API/neighborlist.jl:12-15, :84-111): reusable neighbour-list system.
    `x`/`y` as in the reference; the list is a numpy structured array with fields i, j (1-based int64) and d --
    the memory of Julia's Vector{Tuple{Int,Int,T}} -- owned by the system and overwritten by the next call."""

    def __init__(self, *, x, y=None, cutoff, unitcell=None, parallel=True, show_progress=False, nbatches=(0, 0), lcell=1,
                 device=0):
        self.sys = ParticleSystem(xpositions=x, ypositions=y, unitcell=unitcell, cutoff=cutoff, output=None,
                                  output_name="nb", parallel=parallel, nbatches=nbatches, lcell=lcell, device=device)
        self.show_progress = show_progress
        self._list = np.zeros(0, dtype=nl_dtype(self.sys.dtype))
        self._pinned, self._reused = False, 0
        self.n = 0
        self.n_cutoff_band = 0   # pairs of the last list whose d2 lies within 1 ulp of cutoff^2 (north_star: reported separately)

    def _unpin(self):
        if self._pinned:
            _capi.host_unregister(self._list)
            self._pinned = False

    def __del__(self):
        try:
            self._unpin()
        except Exception:
            pass

    def update(self, x=None, y=None, *, cutoff=None, unitcell=None, parallel=None):
        """update!(system, x, [y]; cutoff, unitcell, parallel) (src/API/neighborlist.jl:159-169)."""
        update(self.sys, xpositions=x, ypositions=y, cutoff=cutoff, unitcell=unitcell, parallel=parallel)
        return self

    def neighborlist(self):
        """neighborlist!(system) (src/API/neighborlist.jl:217-231)."""
        s = self.sys
        s._sync()
        try:
            n = s._h.neighborlist_count()
            if self._list.shape[0] < n:
                self._unpin()
                self._list = np.zeros(max(n, int(1.2 * self._list.shape[0])), dtype=nl_dtype(s.dtype))
                self._reused = 0
            elif n and not self._pinned and self._reused >= 1:
                # the record array is being reused in place: page-lock it once, so that the copy-out of every later list is a
                # direct DMA transfer (clm_host_register; a one-shot list never pays the registration)
                self._pinned = _capi.host_register(self._list)
            self._reused += 1
            if n:
                s._h.neighborlist_copy(self._list)
        except ClmError as e:
            _raise(e)
        self.n = n
        self.n_cutoff_band = int(s._h.stats().n_cutoff_band)
        s.output = self._list[:n]
        return s.output


def neighborlist_(system):
    """neighborlist!(system)"""
    return system.neighborlist()


def neighborlist(*, xpositions=None, ypositions=None, positions=None, cutoff, unitcell=None, parallel=True,
                 show_progress=False, nbatches=(0, 0), lcell=1, device=0):
    """neighborlist(; xpositions, [ypositions], cutoff, unitcell, ...) (src/API/neighborlist.jl:314-340).
    Returns a structured array of (i, j, d): every pair within the cutoff exactly once, i/j unordered for one set,
    (i in x, j in y) for two sets, order unspecified (src/API/neighborlist.jl:75-77)."""
    if (positions is None) == (xpositions is None):
        raise ValueError("Either `positions` OR `xpositions` must be defined.")
    x = positions if xpositions is None else xpositions
    nb = InPlaceNeighborList(x=x, y=ypositions, cutoff=cutoff, unitcell=unitcell, parallel=parallel,
                             show_progress=show_progress, nbatches=nbatches, lcell=lcell, device=device)
    return nb.neighborlist().copy()


class FramePipeline:
    """Independent frames of a trajectory -- the reference's `sys.xpositions .= frame; pairwise!(f, sys)` loop
    (docs/src/ParticleSystem/updating.md) -- through `handles` particle systems that take the frames in turn.

    Every handle has its own copy-in, compute and copy-out streams (clm_set_positions_async + CLM_ASYNC maps), so within a
    handle the copy-in of its next frame, the compute of the current one and the copy-out of the previous one overlap; across
    handles the cell-list build of frame k+1 (a chain of short, latency-bound kernels) runs NEXT TO the pair sweep of frame k,
    which `blocks_per_sm` keeps from filling the SMs' register files (-1: one resident sweep CTA per SM fewer than fit, i.e. 4
    instead of 5 for the Float32 force sweep on B200: the sweep alone gets ~5 % slower, the pair of frames ~15 % faster: 0.658 -> 0.558 ms per 1M-particle LJ frame, L2 flushed
    between frames, tools/diag_e2e_two.py).  Positions and outputs are PINNED host arrays that must stay untouched until
    `synchronize()`; a buffer may be reused once `handles * 2` later frames have been submitted and synchronised or -- simpler
    -- after `synchronize()`.  Self-set LJ energy + forces (the headline map); other maps run through `Handle` directly."""

    def __init__(self, dim, dtype, unitcell, cutoff, lcell=1, handles=2, device=0, blocks_per_sm=-1, streams=None):
        self.dtype = np.dtype(dtype)
        uc = None if unitcell is None else np.asarray(unitcell, dtype=self.dtype)
        cell_type = _capi.NONPERIODIC if uc is None else (_capi.TRICLINIC if uc.ndim == 2 else _capi.ORTHORHOMBIC)
        self.handles = []
        for k in range(int(handles)):
            h = Handle(dim, self.dtype, device)
            if streams is not None:
                h.set_stream(streams[k])
            h.set_box(cell_type, uc, cutoff, lcell)
            if blocks_per_sm and handles > 1:
                h.set_option("blocks_per_sm", int(blocks_per_sm))
            self.handles.append(h)
        self.frames = 0
        self._last = [None] * len(self.handles)

    def next_handle(self):
        """index of the handle (and of the stream passed in `streams`) the next submitted frame runs on."""
        return self.frames % len(self.handles)

    def submit_lj(self, c6, c12, x_pinned, energy_pinned, forces_pinned):
        """enqueue one frame: positions in, LJ energy and forces out (pinned host arrays of the pipeline's dtype)."""
        a = self.next_handle()
        h = self.handles[a]
        cur = (c6, c12, x_pinned, energy_pinned, forces_pinned)
        try:
            try:
                h.set_positions_async(0, x_pinned)
                h.map_lj(c6, c12, energy_pinned, forces_pinned, async_=True)
            except ClmError as e:
                if e.code != 6 or self._last[a] is None:      # CLM_ERR_CAPACITY: the record capacity of this handle's PREVIOUS frame was
                    raise                                     # too small (first frames only); it has been grown: repeat that frame, then this one
                for (p6, p12, px, pe, pf) in (self._last[a], cur):
                    h.set_positions_async(0, px)
                    h.map_lj(p6, p12, pe, pf, async_=True)
        except ClmError as e:
            _raise(e)
        self._last[a] = cur
        self.frames += 1

    def synchronize(self):
        """every submitted frame's outputs are in host memory on return."""
        try:
            for a, h in enumerate(self.handles):
                for attempt in range(4):
                    try:
                        h.synchronize()
                        break
                    except ClmError as e:
                        if e.code != 6 or self._last[a] is None or attempt == 3:
                            raise
                        # the record capacity of this handle's LAST frame was too small (it has been grown): repeat that frame
                        p6, p12, px, pe, pf = self._last[a]
                        h.set_positions_async(0, px)
                        h.map_lj(p6, p12, pe, pf, async_=True)
        except ClmError as e:
            _raise(e)

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

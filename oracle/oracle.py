"""ctypes binding of the CPU ORACLE (test infrastructure only -- see clm_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORTHO, TRICLINIC, NONPERIODIC = 0, 1, 2
ALGO_CELLLIST, ALGO_CELLLIST_NOPROJ, ALGO_NAIVE = 0, 1, 2


def build(force=False):
    so = os.path.join(_HERE, "libclm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("clm_oracle.cpp", "clm_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "-s"], env=env)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libclm_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.ora_create.restype = C.c_void_p
        L.ora_create.argtypes = [C.c_int, C.c_int]
        L.ora_destroy.argtypes = [C.c_void_p]
        L.ora_last_error.restype = C.c_char_p
        L.ora_last_error.argtypes = [C.c_void_p]
        vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
        L.ora_set_box.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int]
        L.ora_set_positions.argtypes = [vp, C.c_int, vp, C.c_int64]
        L.ora_build.argtypes = [vp]
        L.ora_get_box.argtypes = [vp, C.POINTER(C.c_double)]
        L.ora_get_stats.argtypes = [vp, i64p]
        L.ora_candidates.argtypes = [vp, i64p]
        L.ora_map_sum_d_d2.argtypes = [vp, C.c_int, C.c_int, vp, vp, i64p]
        L.ora_map_lj.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
        L.ora_map_coulomb.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        L.ora_map_dist_hist.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, i64p]
        L.ora_map_pairvel.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, i64p, vp]
        L.ora_map_mindist.argtypes = [vp, C.c_int, C.c_int, i64p, i64p, vp]
        L.ora_neighborlist.argtypes = [vp, C.c_int, C.c_int, i64p]
        L.ora_neighborlist_copy.argtypes = [vp, vp, C.c_int64]
        L.ora_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[oracle code {code}] {msg}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One particle system evaluated by the CPU restatement of the reference."""

    def __init__(self, x, cutoff, unitcell=None, y=None, lcell=1, dtype=np.float64, triclinic=None):
        self.L = lib()
        self.dtype = np.dtype(dtype)
        x = np.ascontiguousarray(x, dtype=self.dtype)
        self.dim = x.shape[1] if x.ndim == 2 and x.shape[0] > 0 else (
            x.shape[1] if x.ndim == 2 else (len(unitcell) if unitcell is not None else 3))
        self.x = x.reshape(-1, self.dim)
        self.y = None if y is None else np.ascontiguousarray(y, dtype=self.dtype).reshape(-1, self.dim)
        self.h = self.L.ora_create(self.dim, 0 if self.dtype == np.float32 else 1)
        if not self.h:
            raise ValueError("bad dim/dtype")
        rc = np.array([cutoff], dtype=self.dtype)
        if unitcell is None:
            self._chk(self.L.ora_set_box(self.h, NONPERIODIC, None, 0, _ptr(rc), lcell))
        else:
            uc = np.asarray(unitcell, dtype=self.dtype)
            if uc.ndim == 1:
                ct = ORTHO
                cell = np.ascontiguousarray(uc)
                is_matrix = 0
            else:
                ct = TRICLINIC if triclinic in (None, True) else ORTHO
                cell = np.asfortranarray(uc).ravel(order="F").copy()  # column major, columns = lattice vectors
                is_matrix = 1
            self._chk(self.L.ora_set_box(self.h, ct, _ptr(cell), is_matrix, _ptr(rc), lcell))
        self._chk(self.L.ora_set_positions(self.h, 0, _ptr(self.x), self.x.shape[0]))
        if self.y is not None:
            self._chk(self.L.ora_set_positions(self.h, 1, _ptr(self.y), self.y.shape[0]))
        self._chk(self.L.ora_build(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.ora_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _chk(self, code):
        if code != 0:
            raise OracleError(code, self.L.ora_last_error(self.h).decode())

    # ---- geometry / statistics ----
    def box(self):
        o = np.zeros(60, dtype=np.float64)
        self._chk(self.L.ora_get_box(self.h, o.ctypes.data_as(C.POINTER(C.c_double))))
        N = self.dim
        m = lambda k: o[k:k + N * N].reshape(N, N, order="F").copy()
        return dict(input_unit_cell=m(0), aligned_unit_cell=m(9), rotation=m(18), inv_rotation=m(27),
                    nc=o[36:36 + N].astype(np.int64), cutoff=o[39], cutoff_sqr=o[40], cb_min=o[41:41 + N].copy(),
                    cb_max=o[44:44 + N].copy(), cell_size=o[47:47 + N].copy(), origin=o[50:50 + N].copy(),
                    lcell=int(o[53]), cell_type=int(o[54]))

    def stats(self):
        o = np.zeros(8, dtype=np.int64)
        self._chk(self.L.ora_get_stats(self.h, o.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(n_real=int(o[0]), n_particles=int(o[1]), n_cells_real=int(o[2]), n_cells=int(o[3]),
                    n_real_y=int(o[4]), n_particles_y=int(o[5]), n_cells_real_y=int(o[6]), n_cells_y=int(o[7]))

    def candidates(self):
        o = np.zeros(2, dtype=np.int64)
        self._chk(self.L.ora_candidates(self.h, o.ctypes.data_as(C.POINTER(C.c_int64))))
        return int(o[0]), int(o[1])

    # ---- catalogue ----
    def sum_d_d2(self, algo=0, nbatches=0):
        sd, sd2 = np.zeros(1, self.dtype), np.zeros(1, self.dtype)
        n = C.c_int64(0)
        self._chk(self.L.ora_map_sum_d_d2(self.h, algo, nbatches, _ptr(sd), _ptr(sd2), C.byref(n)))
        return sd[0], sd2[0], n.value

    def lj(self, c6, c12, forces=False, algo=0, nbatches=0):
        p = np.array([c6, c12], dtype=self.dtype)
        e = np.zeros(1, self.dtype)
        f = np.zeros_like(self.x) if forces else None
        self._chk(self.L.ora_map_lj(self.h, algo, nbatches, _ptr(p), _ptr(e), _ptr(f)))
        return (e[0], f) if forces else e[0]

    def coulomb(self, k, wx, wy=None, forces=False, algo=0, nbatches=0):
        wx = np.ascontiguousarray(wx, dtype=self.dtype)
        wy = None if wy is None else np.ascontiguousarray(wy, dtype=self.dtype)
        kk = np.array([k], dtype=self.dtype)
        e = np.zeros(1, self.dtype)
        f = np.zeros_like(self.x) if forces else None
        self._chk(self.L.ora_map_coulomb(self.h, algo, nbatches, _ptr(wx), _ptr(wy), _ptr(kk), _ptr(e), _ptr(f)))
        return (e[0], f) if forces else e[0]

    def dist_hist(self, width, nbins, algo=0, nbatches=0):
        w = np.array([width], dtype=self.dtype)
        c = np.zeros(nbins, dtype=np.int64)
        self._chk(self.L.ora_map_dist_hist(self.h, algo, nbatches, _ptr(w), nbins, c.ctypes.data_as(C.POINTER(C.c_int64))))
        return c

    def pairvel(self, vx, rbins, vy=None, algo=0, nbatches=0):
        vx = np.ascontiguousarray(vx, dtype=self.dtype)
        vy = None if vy is None else np.ascontiguousarray(vy, dtype=self.dtype)
        rb = np.ascontiguousarray(rbins, dtype=self.dtype)
        nb = rb.shape[0] - 1
        c = np.zeros(nb, dtype=np.int64)
        s = np.zeros(nb, dtype=self.dtype)
        self._chk(self.L.ora_map_pairvel(self.h, algo, nbatches, _ptr(vx), _ptr(vy), _ptr(rb), nb,
                                         c.ctypes.data_as(C.POINTER(C.c_int64)), _ptr(s)))
        return c, s

    def mindist(self, algo=0, nbatches=0):
        i, j = C.c_int64(0), C.c_int64(0)
        d = np.zeros(1, self.dtype)
        self._chk(self.L.ora_map_mindist(self.h, algo, nbatches, C.byref(i), C.byref(j), _ptr(d)))
        return i.value, j.value, d[0]

    def neighborlist(self, algo=0, nbatches=0):
        """Returns (i, j, d) arrays, 1-based indices, order unspecified (as in the reference)."""
        n = C.c_int64(0)
        self._chk(self.L.ora_neighborlist(self.h, algo, nbatches, C.byref(n)))
        rec = np.zeros(n.value, dtype=nl_dtype(self.dtype))
        if n.value:
            self._chk(self.L.ora_neighborlist_copy(self.h, _ptr(rec), n.value))
        return rec["i"].copy(), rec["j"].copy(), rec["d"].copy()


def nl_dtype(dtype):
    """Memory layout of Julia's Tuple{Int,Int,T}: 24 bytes for Float32 and Float64."""
    dtype = np.dtype(dtype)
    return np.dtype({"names": ["i", "j", "d"], "formats": [np.int64, np.int64, dtype], "offsets": [0, 8, 16], "itemsize": 24})


def sorted_pairs(i, j, d=None, ordered=False):
    """Canonical form of a neighbour list: rows (min(i,j), max(i,j)) sorted lexicographically."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    a, b = (i, j) if ordered else (np.minimum(i, j), np.maximum(i, j))
    order = np.lexsort((b, a))
    if d is None:
        return np.stack([a[order], b[order]], axis=1)
    return np.stack([a[order], b[order]], axis=1), np.asarray(d)[order]

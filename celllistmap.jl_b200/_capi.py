"""ctypes binding of libclm_b200.so (include/clm_b200.h) -- the same entry points a Julia host binds
with `ccall`.  There is NO CPU fallback: if the library is missing or no CUDA device is
present every call fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CLM_SO", os.path.join(_HERE, "libclm_b200.so"))

F32, F64 = 0, 1
ORTHORHOMBIC, TRICLINIC, NONPERIODIC = 0, 1, 2
RESET, OUT_DEVICE, PROFILE, ASYNC = 1, 2, 4, 8

STATUS_NAMES = {
    0: "CLM_OK", 1: "CLM_ERR_INVALID_COORDINATES", 2: "CLM_ERR_UNIT_CELL", 3: "CLM_ERR_ARGUMENT", 4: "CLM_ERR_STATE",
    5: "CLM_ERR_DIMENSION", 6: "CLM_ERR_CAPACITY", 7: "CLM_ERR_CUDA", 8: "CLM_ERR_COMM", 9: "CLM_ERR_UNSUPPORTED",
}

# every symbol include/clm_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "clm_create", "clm_destroy", "clm_last_error", "clm_set_stream", "clm_synchronize", "clm_set_box", "clm_get_box",
    "clm_set_positions", "clm_set_positions_async", "clm_build", "clm_map_lj", "clm_map_coulomb", "clm_map_dist_hist", "clm_map_pairvel",
    "clm_map_mindist", "clm_map_sum_d_d2", "clm_neighborlist", "clm_neighborlist_copy", "clm_get_stats",
    "clm_set_option", "clm_version", "clm_measure_fma_peak", "clm_host_register", "clm_host_unregister", "clm_set_foreign", "clm_set_foreign_mask", "clm_read_ints", "clm_cell_coords", "clm_select_layers",
    "clm_custom_compile", "clm_custom_log", "clm_map_custom", "clm_custom_check",
    "clm_comm_unique_id", "clm_comm_init", "clm_comm_destroy", "clm_slab_range", "clm_slab_update", "clm_comm_allreduce_sum", "clm_slab_info",
]


class BoxInfo(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dtype", C.c_int32), ("cell_type", C.c_int32), ("lcell", C.c_int32),
        ("nc", C.c_int64 * 3), ("cutoff", C.c_double), ("cutoff_sqr", C.c_double),
        ("input_unit_cell", C.c_double * 9), ("aligned_unit_cell", C.c_double * 9),
        ("rotation", C.c_double * 9), ("inv_rotation", C.c_double * 9),
        ("computing_box_min", C.c_double * 3), ("computing_box_max", C.c_double * 3),
        ("cell_size", C.c_double * 3), ("origin", C.c_double * 3),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_real", C.c_int64 * 2), ("n_total", C.c_int64 * 2), ("n_cells", C.c_int64), ("n_cells_real", C.c_int64 * 2),
        ("n_tiles", C.c_int64), ("n_pairs", C.c_int64), ("n_cutoff_band", C.c_int64),
        ("build_ms", C.c_double), ("map_ms", C.c_double), ("sweep_ms", C.c_double), ("n_sm", C.c_int32), ("launches", C.c_int32),
    ]


class CustomInfo(C.Structure):
    _fields_ = [("nscalar", C.c_int32), ("npart", C.c_int32), ("naux", C.c_int32), ("hist", C.c_int32),
                ("scalar_min_mask", C.c_int32), ("scalar_max_mask", C.c_int32)]


class ClmError(RuntimeError):
    """Error class of the C ABI.  The reference-facing layer maps codes 1-3 to ValueError (Julia's
    ArgumentError), 5 to the DimensionMismatch analogue, the rest to RuntimeError (ErrorException)."""

    def __init__(self, code, msg):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.message = msg


_LIB = None


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it; every rank passes it to Handle.comm_init)"""
    buf = (C.c_char * 128)()
    rc = lib().clm_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc:
        raise ClmError(rc, lib().clm_last_error(None).decode())
    return bytes(buf)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `python celllistmap.jl_b200/build.py` "
            "(__graft_entry__.build()).  There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, i64, i64p, ci = C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_int
    L.clm_version.restype = ci
    L.clm_last_error.restype = C.c_char_p
    L.clm_last_error.argtypes = [vp]
    L.clm_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci]
    L.clm_destroy.argtypes = [vp]
    L.clm_set_stream.argtypes = [vp, vp]
    L.clm_synchronize.argtypes = [vp]
    L.clm_set_box.argtypes = [vp, ci, vp, ci, vp, ci]
    L.clm_get_box.argtypes = [vp, C.POINTER(BoxInfo)]
    L.clm_set_positions.argtypes = [vp, ci, vp, i64, ci]
    L.clm_set_positions_async.argtypes = [vp, ci, vp, i64]
    L.clm_comm_unique_id.argtypes = [vp]
    L.clm_comm_init.argtypes = [vp, vp, ci, ci]
    L.clm_comm_destroy.argtypes = [vp]
    L.clm_slab_range.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.clm_slab_update.argtypes = [vp, vp, i64, ci]
    L.clm_comm_allreduce_sum.argtypes = [vp, vp, i64, ci, ci]
    L.clm_slab_info.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.clm_build.argtypes = [vp]
    L.clm_map_lj.argtypes = [vp, vp, ci, vp, vp]
    L.clm_map_coulomb.argtypes = [vp, vp, vp, vp, ci, vp, vp]
    L.clm_map_dist_hist.argtypes = [vp, vp, ci, ci, vp]
    L.clm_map_pairvel.argtypes = [vp, vp, vp, vp, ci, ci, vp, vp]
    L.clm_map_mindist.argtypes = [vp, ci, vp, vp, vp]
    L.clm_map_sum_d_d2.argtypes = [vp, ci, vp, vp, vp]
    L.clm_neighborlist.argtypes = [vp, ci, i64p]
    L.clm_neighborlist_copy.argtypes = [vp, vp, i64, ci]
    L.clm_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.clm_set_option.argtypes = [vp, C.c_char_p, i64]
    L.clm_measure_fma_peak.argtypes = [ci, ci, C.POINTER(C.c_double)]
    L.clm_host_register.argtypes = [vp, i64]
    L.clm_host_unregister.argtypes = [vp]
    L.clm_set_foreign.argtypes = [vp, ci, vp, i64, ci]
    L.clm_set_foreign_mask.argtypes = [vp, ci, vp, i64, ci]
    L.clm_read_ints.argtypes = [vp, vp, C.c_int32, C.POINTER(C.c_int32)]
    L.clm_cell_coords.argtypes = [vp, vp, i64, ci, ci, vp]
    L.clm_select_layers.argtypes = [vp, vp, i64, ci, C.POINTER(C.c_int32), ci, vp, vp, i64, vp, vp, vp]
    L.clm_custom_compile.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(CustomInfo)]
    L.clm_custom_log.restype = C.c_char_p
    L.clm_custom_log.argtypes = [vp]
    L.clm_map_custom.argtypes = [vp, C.c_int32, vp, ci, vp, vp, ci, ci, vp, vp, vp, vp]
    L.clm_custom_check.argtypes = [C.c_char_p, C.c_char_p, ci, C.c_char_p, i64]
    for name in SYMBOLS:
        getattr(L, name)
    _LIB = L
    return L


def nl_dtype(dtype):
    """Memory layout of Julia's Tuple{Int,Int,T} (24 bytes for Float32 and Float64)."""
    return np.dtype({"names": ["i", "j", "d"], "formats": [np.int64, np.int64, np.dtype(dtype)], "offsets": [0, 8, 16],
                     "itemsize": 24})


def _is_torch(a):
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


def _addr(a):
    """(address, on_device) of a numpy array / torch tensor / None."""
    if a is None:
        return None, False
    if _is_torch(a):
        if not a.is_contiguous():
            raise ValueError("device tensors must be contiguous")
        return C.c_void_p(a.data_ptr()), bool(a.is_cuda)
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("arrays must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p), False


def host_register(a):
    """page-lock the memory of a numpy array the caller reuses across calls (clm_host_register); False when the driver refuses."""
    return lib().clm_host_register(a.ctypes.data_as(C.c_void_p), int(a.nbytes)) == 0


def host_unregister(a):
    lib().clm_host_unregister(a.ctypes.data_as(C.c_void_p))


def measure_fma_peak(dtype, device=0):
    """measured SIMT FMA peak in TFLOP/s (FP32 or FP64) of `device`."""
    t = C.c_double(0)
    code = lib().clm_measure_fma_peak(int(device), F32 if np.dtype(dtype) == np.float32 else F64, C.byref(t))
    if code != 0:
        raise ClmError(code, "FMA microbenchmark failed")
    return t.value


def custom_check(source, name, dtype=np.float64):
    """compile-only check of a user pair function (every sweep mode; needs libnvrtc, no device).  Returns the NVRTC
    log; raises ClmError if the source does not compile."""
    buf = C.create_string_buffer(1 << 16)
    code = lib().clm_custom_check(source.encode(), name.encode(), F32 if np.dtype(dtype) == np.float32 else F64, buf, len(buf))
    log = buf.value.decode(errors="replace")
    if code != 0:
        raise ClmError(code, log)
    return log


class Handle:
    """One clm_handle: a particle system resident on one B200."""

    def __init__(self, dim, dtype, device=0):
        self.L = lib()
        self.dim = int(dim)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("dtype must be float32 or float64")
        self.h = C.c_void_p()
        code = self.L.clm_create(C.byref(self.h), self.dim, F32 if self.dtype == np.float32 else F64, int(device), 1)
        if code != 0:
            msg = self.L.clm_last_error(None).decode()
            self.h = None
            raise ClmError(code, msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.clm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, code):
        if code != 0:
            raise ClmError(code, self.L.clm_last_error(self.h).decode())

    def _scalar(self, v):
        return np.array([v], dtype=self.dtype)

    # ---- box / positions / build ----
    def set_box(self, cell_type, unitcell, cutoff, lcell=1):
        rc = self._scalar(cutoff)
        if cell_type == NONPERIODIC or unitcell is None:
            self._chk(self.L.clm_set_box(self.h, NONPERIODIC, None, 0, rc.ctypes.data_as(C.c_void_p), int(lcell)))
            return
        uc = np.asarray(unitcell, dtype=self.dtype)
        if uc.ndim == 1:
            cell, is_matrix = np.ascontiguousarray(uc), 0
        else:
            cell, is_matrix = np.ascontiguousarray(uc.T).ravel().copy(), 1  # column-major, columns = lattice vectors
        self._chk(self.L.clm_set_box(self.h, int(cell_type), cell.ctypes.data_as(C.c_void_p), is_matrix,
                                     rc.ctypes.data_as(C.c_void_p), int(lcell)))

    def get_box(self):
        b = BoxInfo()
        self._chk(self.L.clm_get_box(self.h, C.byref(b)))
        return b

    def set_positions(self, which, x):
        """x: (n, dim) numpy array (host) or contiguous torch CUDA tensor of the handle's dtype; None removes set y."""
        if x is None:
            self._chk(self.L.clm_set_positions(self.h, int(which), None, 0, 0))
            return
        if not _is_torch(x):
            x = np.ascontiguousarray(x, dtype=self.dtype)
        n = int(x.shape[0])
        p, dev = _addr(x)
        if n == 0:
            p = C.c_void_p(1) if which == 1 else None  # non-NULL + n = 0: empty second set
        self._chk(self.L.clm_set_positions(self.h, int(which), p, n, 1 if dev else 0))

    def set_positions_async(self, which, x):
        """pipelined frames: x is a PINNED host array (numpy view of a pinned torch tensor) of the handle's dtype; the copy
        is enqueued on the handle's copy-in stream and the call returns at once (see clm_set_positions_async)."""
        if _is_torch(x):
            if x.is_cuda:
                raise ValueError("set_positions_async takes pinned HOST memory")
            x = x.numpy()
        if x.dtype != self.dtype or not x.flags["C_CONTIGUOUS"]:
            raise ValueError("set_positions_async needs a C-contiguous array of the handle's dtype (no implicit copy)")
        self._chk(self.L.clm_set_positions_async(self.h, int(which), x.ctypes.data_as(C.c_void_p), int(x.shape[0])))

    def set_foreign(self, which, x):
        """particles owned by other ranks that are within the stencil reach of this rank's slab (None / empty: none)."""
        if x is None or int(x.shape[0]) == 0:
            self._chk(self.L.clm_set_foreign(self.h, int(which), None, 0, 0))
            return
        if not _is_torch(x):
            x = np.ascontiguousarray(x, dtype=self.dtype)
        p, dev = _addr(x)
        self._chk(self.L.clm_set_foreign(self.h, int(which), p, int(x.shape[0]), 1 if dev else 0))

    def set_foreign_mask(self, which, mask):
        """rows of set_positions(which, ...) that belong to other ranks (uint8, one per row; None removes the mask)."""
        if mask is None:
            self._chk(self.L.clm_set_foreign_mask(self.h, int(which), None, 0, 0))
            return
        if _is_torch(mask):
            import torch
            mask = mask.to(torch.uint8).contiguous()
        else:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        p, dev = _addr(mask)
        self._chk(self.L.clm_set_foreign_mask(self.h, int(which), p, int(mask.shape[0]), 1 if dev else 0))
        self._keep_mask = mask

    def read_ints(self, t):
        """a small int32 CUDA tensor (<= 64 entries) -> Python ints, without a DMA copy (clm_read_ints)."""
        n = int(t.numel())
        out = (C.c_int32 * max(n, 1))()
        self._chk(self.L.clm_read_ints(self.h, _addr(t)[0], n, out))
        return [int(out[k]) for k in range(n)]

    def cell_coords(self, x, axis, out=None):
        """0-based reference-cell index along `axis` of every row of x (numpy in -> numpy out, torch CUDA in -> torch out)."""
        n = int(x.shape[0])
        if _is_torch(x):
            import torch
            out = torch.empty(n, dtype=torch.int32, device=x.device) if out is None else out
            self._chk(self.L.clm_cell_coords(self.h, _addr(x)[0], n, 1, int(axis), _addr(out)[0]))
            return out
        x = np.ascontiguousarray(x, dtype=self.dtype)
        out = np.empty(n, np.int32) if out is None else out
        self._chk(self.L.clm_cell_coords(self.h, _addr(x)[0], n, 0, int(axis), _addr(out)[0]))
        return out

    def select_layers(self, x, axis, ranges, merge, out_a, out_b, counts, idx_a=None, idx_b=None):
        """one-pass face selection (torch CUDA tensors; enqueue only): see clm_select_layers."""
        r = (C.c_int32 * 4)(*[int(v) for v in ranges])
        self._chk(self.L.clm_select_layers(self.h, _addr(x)[0], int(x.shape[0]), int(axis), r, 1 if merge else 0,
                                           _addr(out_a)[0], _addr(out_b)[0], int(out_a.shape[0]), _addr(counts)[0],
                                           _addr(idx_a)[0], _addr(idx_b)[0]))

    # ---- slab decomposition over NCCL behind the C ABI (clm_comm.cu) ----
    def comm_init(self, unique_id, rank, world):
        """unique_id: the 128 bytes of comm_unique_id() (called on rank 0 and handed to every rank)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._chk(self.L.clm_comm_init(self.h, C.cast(buf, C.c_void_p), int(rank), int(world)))

    def comm_destroy(self):
        self._chk(self.L.clm_comm_destroy(self.h))

    def slab_range(self):
        lo, hi = C.c_int32(0), C.c_int32(0)
        self._chk(self.L.clm_slab_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def slab_update(self, x):
        if not _is_torch(x):
            x = np.ascontiguousarray(x, dtype=self.dtype)
        p, dev = _addr(x)
        self._chk(self.L.clm_slab_update(self.h, p, int(x.shape[0]), 1 if dev else 0))

    def comm_allreduce_sum(self, buf):
        """in-place sum over the ranks of a numpy array (host) or torch CUDA tensor of the handle's real type, int64 or float64"""
        dt = np.dtype(str(buf.dtype).replace("torch.", ""))
        kind = 0 if dt == self.dtype else (1 if dt == np.int64 else (2 if dt == np.float64 else -1))
        p, dev = _addr(buf)
        n = int(buf.numel()) if _is_torch(buf) else int(buf.size)
        self._chk(self.L.clm_comm_allreduce_sum(self.h, p, n, kind, 1 if dev else 0))

    def slab_info(self):
        a, b, r, w = C.c_int64(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self._chk(self.L.clm_slab_info(self.h, C.byref(a), C.byref(b), C.byref(r), C.byref(w)))
        return a.value, b.value, r.value, w.value

    def build(self):
        self._chk(self.L.clm_build(self.h))

    def synchronize(self):
        self._chk(self.L.clm_synchronize(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.L.clm_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def set_option(self, name, value):
        self._chk(self.L.clm_set_option(self.h, name.encode(), int(value)))

    def stats(self):
        s = Stats()
        self._chk(self.L.clm_get_stats(self.h, C.byref(s)))
        return s

    # ---- maps: outputs are numpy arrays (host) or torch CUDA tensors (device, enqueue only) ----
    def _flags(self, reset, outs, profile):
        dev = [d for d in (_addr(o)[1] for o in outs if o is not None)]
        if any(dev) and not all(dev):
            raise ValueError("outputs of one call must be all host or all device arrays")
        return (RESET if reset else 0) | (OUT_DEVICE if any(dev) else 0) | (PROFILE if profile else 0)

    def map_lj(self, c6, c12, energy, forces=None, reset=True, profile=False, async_=False):
        """async_=True (pipelined frames): host outputs must be pinned; they are valid after synchronize()."""
        p = np.array([c6, c12], dtype=self.dtype)
        fl = self._flags(reset, (energy, forces), profile) | (ASYNC if async_ else 0)
        self._chk(self.L.clm_map_lj(self.h, p.ctypes.data_as(C.c_void_p), fl, _addr(energy)[0], _addr(forces)[0]))

    def map_coulomb(self, k, wx, wy, energy, forces=None, reset=True, profile=False):
        kk = self._scalar(k)
        fl = self._flags(reset, (energy, forces, wx, wy), profile)
        self._chk(self.L.clm_map_coulomb(self.h, _addr(wx)[0], _addr(wy)[0], kk.ctypes.data_as(C.c_void_p), fl,
                                         _addr(energy)[0], _addr(forces)[0]))

    def map_dist_hist(self, width, counts, reset=True, profile=False):
        w = self._scalar(width)
        fl = self._flags(reset, (counts,), profile)
        self._chk(self.L.clm_map_dist_hist(self.h, w.ctypes.data_as(C.c_void_p), int(counts.shape[0]), fl, _addr(counts)[0]))

    def map_pairvel(self, vx, vy, rbins, counts, sums, reset=True, profile=False):
        rb = np.ascontiguousarray(rbins, dtype=self.dtype)
        fl = self._flags(reset, (counts, sums, vx, vy), profile)
        self._chk(self.L.clm_map_pairvel(self.h, _addr(vx)[0], _addr(vy)[0], rb.ctypes.data_as(C.c_void_p),
                                         int(rb.shape[0]) - 1, fl, _addr(counts)[0], _addr(sums)[0]))

    def map_mindist(self, i_out, j_out, d_out, reset=True, profile=False):
        fl = self._flags(reset, (i_out, j_out, d_out), profile)
        self._chk(self.L.clm_map_mindist(self.h, fl, _addr(i_out)[0], _addr(j_out)[0], _addr(d_out)[0]))

    def map_sum_d_d2(self, sum_d, sum_d2, npairs, reset=True, profile=False):
        fl = self._flags(reset, (sum_d, sum_d2, npairs), profile)
        self._chk(self.L.clm_map_sum_d_d2(self.h, fl, _addr(sum_d)[0], _addr(sum_d2)[0], _addr(npairs)[0]))

    def custom_compile(self, source, name):
        """compile a user pair function (CUDA C++ source of a struct `name`, see include/clm_b200.h); returns
        (functor_id, CustomInfo)."""
        fid, info = C.c_int32(-1), CustomInfo()
        self._chk(self.L.clm_custom_compile(self.h, source.encode(), name.encode(), C.byref(fid), C.byref(info)))
        return fid.value, info

    def custom_log(self):
        return self.L.clm_custom_log(self.h).decode(errors="replace")

    def map_custom(self, fid, params=(), aux_x=None, aux_y=None, scalars=None, per_particle=None, hist_counts=None, hist_sums=None,
                   reset=True, profile=False):
        par = np.ascontiguousarray(params, dtype=self.dtype).ravel()
        fl = self._flags(reset, (scalars, per_particle, hist_counts, hist_sums, aux_x, aux_y), profile)
        nbins = int(hist_counts.shape[0]) if hist_counts is not None else 0
        self._chk(self.L.clm_map_custom(self.h, int(fid), par.ctypes.data_as(C.c_void_p) if par.size else None, int(par.size),
                                        _addr(aux_x)[0], _addr(aux_y)[0], nbins, fl, _addr(scalars)[0], _addr(per_particle)[0],
                                        _addr(hist_counts)[0], _addr(hist_sums)[0]))

    def neighborlist_count(self, profile=False):
        n = C.c_int64(0)
        self._chk(self.L.clm_neighborlist(self.h, PROFILE if profile else 0, C.byref(n)))
        return n.value

    def neighborlist_copy(self, records):
        """records: numpy structured array of nl_dtype (host) or a torch CUDA uint8/int64 tensor of >= 24*n bytes."""
        p, dev = _addr(records)
        cap = (records.numel() * records.element_size() // 24) if _is_torch(records) else records.shape[0]
        self._chk(self.L.clm_neighborlist_copy(self.h, p, int(cap), 1 if dev else 0))

"""CPU-side checks (no GPU, no compute calls): the C-ABI library builds, loads and exports every symbol
include/clm_b200.h declares; the product path fails loudly without a CUDA device; the product never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clm():
    import __graft_entry__ as ge
    ge.build()
    import celllistmap_b200 as c
    return c


def test_library_exports_every_declared_symbol(clm):
    hdr = open(os.path.join(ROOT, "include", "clm_b200.h")).read()
    declared = sorted(set(re.findall(r"CLM_API\s+[\w\s\*]+?\b(clm_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(clm.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/clm_b200.h but not exported"
    assert sorted(clm._capi.SYMBOLS) == declared, "the ctypes binding must cover exactly the declared ABI"
    assert lib.clm_version() == 100


def test_struct_layouts_match_header(clm):
    # sizes the C compiler gives the two ABI structs (computed from the header's field list)
    assert ctypes.sizeof(clm._capi.BoxInfo) == 4 * 4 + 3 * 8 + 2 * 8 + 4 * 9 * 8 + 4 * 3 * 8
    assert ctypes.sizeof(clm._capi.Stats) == (2 + 2 + 1 + 2 + 1 + 1 + 1) * 8 + 3 * 8 + 2 * 4


def test_no_cpu_fallback(clm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        clm.ParticleSystem(xpositions=np.random.rand(10, 3), unitcell=[1, 1, 1], cutoff=0.1, output=0.0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        clm.neighborlist(xpositions=np.random.rand(10, 3), cutoff=0.1)


def test_argument_errors_before_any_device_work(clm):
    with pytest.raises(ValueError, match="positions` OR `xpositions"):
        clm.ParticleSystem(unitcell=[1, 1, 1], cutoff=0.1, output=0.0)
    with pytest.raises(ValueError, match="positions` OR `xpositions"):
        clm.ParticleSystem(positions=np.zeros((2, 3)), xpositions=np.zeros((2, 3)), unitcell=[1, 1, 1], cutoff=0.1, output=0.0)
    with pytest.raises(ValueError, match="Could not infer dimension"):
        clm.ParticleSystem(xpositions=np.zeros((0, 3))[:, :0], cutoff=0.1, output=0.0)
    with pytest.raises(clm.DimensionMismatch, match="Incompatible dimensions"):
        clm.ParticleSystem(xpositions=np.zeros((4, 2)), unitcell=[1, 1, 1], cutoff=0.1, output=0.0)
    with pytest.raises(ValueError, match="square"):
        clm.ParticleSystem(xpositions=np.zeros((4, 3)), unitcell=np.zeros((3, 2)), cutoff=0.1, output=0.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "celllistmap.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"libclm_oracle", r"#include\s+[\"<][^\">]*oracle", r"oracle[/\\]", r"ora_\w+\("):
                    assert not re.search(pat, src, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def test_workload_generators_are_reproducible():
    import workloads as W
    a, b = W.c1_neighborlist(100), W.c1_neighborlist(100)
    assert np.array_equal(a["x"], b["x"]) and a["x"].min() >= 0 and a["x"].max() < 1
    assert int(W.splitmix64(321, 1)[0]) == int(W.splitmix64(321, 3)[0])
    w = W.c2_argon(10, np.float32)
    assert w["x"].shape == (1000, 3) and abs(w["L"] - 10 * W.ARGON_RHO ** (-1 / 3)) < 1e-12
    # counter-based stream: the multi-GPU slab generator reproduces the single-array lattice (as a set)
    import bench_multi
    parts = [bench_multi.slab_lattice(r, 2, 8, 4, np.float64)[0] for r in range(2)]
    full = W.c2_argon(8, np.float64)["x"]
    key = lambda x: sorted(map(tuple, np.round(x, 9).tolist()))
    assert key(np.concatenate(parts)) == key(full)


USER_SRC = """
struct Coordination {   // per-particle neighbour count + sum of 1/d
    static constexpr int NSCALAR = 1, NPART = 1, NAUX = 0, HIST = 0;
    template <class T, class Out>
    __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
        out.add_scalar(0, par[0] / p.d());
        out.add_i(0, T(1));
    }
};
"""


def test_user_pair_function_compiles_without_a_device(clm):
    """clm_custom_check: NVRTC compiles a user functor against the embedded device headers for every sweep mode and
    both precisions (compile only -- no compute, no device)."""
    f = clm.CustomPairFunction(USER_SRC, "Coordination", params=(1.0,))
    try:
        f.check(np.float32)
    except RuntimeError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("libnvrtc is not installed here")
        raise
    f.check(np.float64)
    with pytest.raises(ValueError, match=r"user_pair_function\(\d+\): error"):
        clm.CustomPairFunction(USER_SRC.replace("p.d()", "p.dist()"), "Coordination").check()
    with pytest.raises(ValueError, match="NSCALAR must be in 0..8"):
        clm.CustomPairFunction(USER_SRC.replace("NSCALAR = 1", "NSCALAR = 9"), "Coordination").check()
    with pytest.raises(ValueError, match="identifier"):
        clm.CustomPairFunction(USER_SRC, "not an identifier").check()


def test_header_is_plain_c_and_library_loads_from_c(clm):
    """include/clm_b200.h compiles as strict C99 and a C program can dlopen the library and call it."""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = os.path.join(tempfile.mkdtemp(prefix="clm_abi_"), "abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe, "-ldl"])
    out = subprocess.run([exe, clm.SO_PATH], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "version 100" in out.stdout
    assert f"sizeof clm_box_info {ctypes.sizeof(clm._capi.BoxInfo)} clm_stats {ctypes.sizeof(clm._capi.Stats)} clm_custom_info {ctypes.sizeof(clm._capi.CustomInfo)}" in out.stdout
    assert "Dimension must be 2 or 3" in out.stdout


def test_wrap_relative_to_kats(clm):
    """public helper wrap_relative_to (src/internals/CellOperations.jl:102-127) against the reference's own vectors,
    with matrix and side-vector cells of either sign (test/internals/CellOperations.jl:7-26)"""
    from golden import kats as K
    for x, y, xy, yx in K.WRAP_RELATIVE_KATS:
        for cell in (np.diag([10.0, 10.0]), np.diag([-10.0, -10.0]), [10.0, 10.0], [-10.0, -10.0]):
            assert np.allclose(clm.wrap_relative_to(x, y, cell), xy, atol=1e-12)
            assert np.allclose(clm.wrap_relative_to(y, x, cell), yx, atol=1e-12)

"""In-tree build of libclm_b200.so (hand-written sm_100a CUDA + the C ABI of include/clm_b200.h).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
Usage: python celllistmap.jl_b200/build.py [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.environ.get("CLM_SO", os.path.join(HERE, "libclm_b200.so"))
UNITS = ["clm_api", "clm_map_lj", "clm_map_hist", "clm_map_misc"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("CLM_NVCC_EXTRA", "").split()
FLAGS = EXTRA + ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "clm_b200.h"))
    out.append(os.path.abspath(__file__))
    return out


def is_stale():
    return (not os.path.exists(SO)) or _newest(_sources()) > os.path.getmtime(SO)


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link the shared library.  Returns its path."""
    if not force and not is_stale():
        return SO
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}: cannot build libclm_b200.so (there is no CPU fallback)")
    os.makedirs(OBJ, exist_ok=True)

    def cc(unit):
        src, obj = os.path.join(CSRC, unit + ".cu"), os.path.join(OBJ, unit + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(cc, UNITS))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    cmd = [NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

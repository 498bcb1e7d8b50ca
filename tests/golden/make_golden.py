"""Generates the committed golden fixtures from the reference's own test data.

Run once in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Outputs (committed):
    tests/golden/namd_frames.npz   float32 coordinates of the 8 NAMD DCD frames used by
                                   test/applications/namd/compare_with_namd.jl (9 999 Ne atoms each)
    tests/golden/argon_cubic.npy   float64 coordinates of test/applications/gromacs/argon/cubic.pdb
                                   (= CellListMap.argon_pdb_file, 100 Ar atoms)
The expected VALUES live in tests/golden/kats.py, each with the reference file:line it was
transcribed from.  Nothing here copies reference source code; only test data is converted.
"""
import os
import struct

import numpy as np

REF = "/root/reference/test/applications"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_dcd_first_frame(path):
    with open(path, "rb") as fh:
        b = fh.read()
    pos = 0

    def rec():
        nonlocal pos
        (n,) = struct.unpack_from("<i", b, pos)
        data = b[pos + 4:pos + 4 + n]
        (n2,) = struct.unpack_from("<i", b, pos + 4 + n)
        assert n == n2, "bad Fortran record"
        pos += 8 + n
        return data

    hdr = rec()
    assert hdr[:4] == b"CORD"
    icntrl = struct.unpack_from("<20i", hdr, 4)
    has_cell = icntrl[10] != 0
    rec()  # titles
    (natoms,) = struct.unpack("<i", rec())
    if has_cell:
        rec()
    xyz = np.stack([np.frombuffer(rec(), dtype="<f4") for _ in range(3)], axis=1)
    assert xyz.shape == (natoms, 3)
    return np.ascontiguousarray(xyz)


def read_pdb(path):
    out = []
    with open(path) as fh:
        for line in fh:
            if line.startswith(("ATOM", "HETATM")):
                out.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.array(out, dtype=np.float64)


if __name__ == "__main__":
    frames = {n: read_dcd_first_frame(f"{REF}/namd/{n}.dcd") for n in ("o1", "o2", "o3", "o4", "o5", "o6", "t1", "t2")}
    np.savez_compressed(os.path.join(OUT, "namd_frames.npz"), **frames)
    ar = read_pdb(f"{REF}/gromacs/argon/cubic.pdb")
    assert ar.shape == (100, 3)
    np.save(os.path.join(OUT, "argon_cubic.npy"), ar)
    print({k: v.shape for k, v in frames.items()}, ar.shape)

"""Shared helpers of the GPU parity tests: the north_star tolerances, and error reports for force maps.

Bars (BASELINE.json north_star): energies / forces / sums within 1e-10 relative in Float64 and 1e-5 in Float32, "results
must match the reference's own implementation on the same inputs".  For a Float32 input the reference computes in
Float32, so the comparison partner is the oracle run in the SAME precision: device and oracle then work on bit-identical
wrapped coordinates (the wrap is repeated operation for operation) and differ only in the pair arithmetic (FMA
contraction of d2, reciprocal, summation order).  The distance of either to exact (Float64) arithmetic on the same
Float32 inputs is a property of Float32 COORDINATES, not of the kernel: a pair at distance r whose coordinates carry an
absolute rounding error delta ~ ulp(L)/2 from the wrap has a force off by (m + 1) * delta / r relative (m = 12 for
the LJ repulsion), e.g. 13 * 1.5e-5 / 1.8 = 1.1e-4 for the 1M-particle C2 system (L = 360 A, closest pair 1.8 A).
Every force test REPORTS both numbers and asserts the same-precision one at the north_star bar and the Float64 one at
that conditioning bound.
"""
import numpy as np

RTOL = {np.dtype(np.float64): 1e-10, np.dtype(np.float32): 1e-5}


def conditioning_bound(dtype, coord_scale, rmin, power):
    """relative force error that rounding the wrapped COORDINATES to `dtype` can cause: (power + 1) * 4 ulp(coord) / rmin"""
    return 4.0 * (power + 1) * np.spacing(dtype(coord_scale)) / rmin


def force_report(name, f_gpu, f_same, f_64):
    """max-norm errors of a force array relative to the largest force component: against the oracle in the same
    precision (the bar) and against the Float64 oracle on the same inputs (conditioning)."""
    scale = np.abs(f_64).max()
    err_same = np.abs(f_gpu.astype(np.float64) - f_same.astype(np.float64)).max() / scale
    err_64 = np.abs(f_gpu.astype(np.float64) - f_64).max() / scale
    ref_64 = np.abs(f_same.astype(np.float64) - f_64).max() / scale
    print(f"[parity] {name}: max|F - F_oracle(same precision)| / max|F| = {err_same:.3e}; vs Float64 oracle = {err_64:.3e} "
          f"(the oracle's own same-precision arithmetic vs Float64: {ref_64:.3e})")
    return err_same, err_64, ref_64

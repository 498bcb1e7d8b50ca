#!/bin/bash
# CTA size / register cap of the Newton's-third-law sweep: times every variants/libclm_*.so (built here, on the build box, with
#   CLM_NVCC_EXTRA="-DCLM_SWEEP_THREADS=64 -DCLM_N3_MAXNREG=88" CLM_SO=variants/libclm_t64r88.so python celllistmap.jl_b200/build.py --force
# ) on the C2 workload next to the shipped library.  The warps of a CTA are independent, so the CTA size only sets the
# granularity at which registers limit the resident warps: 128 x 96 regs -> 20 warps/SM, 64 x 88 -> 22, 32 x 88 -> 23.
# usage (on the GPU box): bash tools/tune_n3_cta.sh
python tools/time_n3.py 100 f32 2>&1 | grep "n3=1\|n3=2"
for so in variants/libclm_*.so; do
  CLM_SO=$PWD/$so timeout 300 python tools/time_n3.py 100 f32 2>&1 | grep "n3=1\|n3=2"
done

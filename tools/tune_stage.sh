#!/bin/bash
# Rebuild the library with different per-warp staging sizes and time the LJ force sweep (tools/tune_sweep.py).
# usage (on the GPU box): bash tools/tune_stage.sh
for sb in 4096 6144 8192 12288; do
  so=/tmp/libclm_b200_s$sb.so
  CLM_NVCC_EXTRA="-DCLM_STAGE_BYTES_F32=$sb -DCLM_STAGE_BYTES_F64=$sb" CLM_SO=$so python celllistmap.jl_b200/build.py --force > /dev/null
  echo "STAGE_BYTES=$sb"
  CLM_SO=$so timeout 300 python tools/tune_sweep.py 100 2>&1 | grep -E "sub=2|sub=3"
done

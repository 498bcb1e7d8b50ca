#!/bin/bash
# Rebuild the library with different per-warp staging sizes and time (a) the LJ force sweep (tools/tune_sweep.py:
# CLM_STAGE_BYTES_F32 / _F64, functors without a side array) and (b) the pair-velocity map (tools/prof_c4.py:
# CLM_STAGE_BYTES_F32_AUX / _F64_AUX, functors that stage a per-record side array next to the records).
# usage (on the GPU box): bash tools/tune_stage.sh
for sb in 4096 6144 8192 12288; do
  so=/tmp/libclm_b200_s$sb.so
  CLM_NVCC_EXTRA="-DCLM_STAGE_BYTES_F32=$sb -DCLM_STAGE_BYTES_F64=$sb -DCLM_STAGE_BYTES_F32_AUX=$sb -DCLM_STAGE_BYTES_F64_AUX=$sb" CLM_SO=$so python celllistmap.jl_b200/build.py --force > /dev/null
  echo "STAGE_BYTES=$sb"
  CLM_SO=$so timeout 300 python tools/tune_sweep.py 100 2>&1 | grep -E "sub=2|sub=3"
  CLM_SO=$so timeout 300 python tools/prof_c4.py 1000000 3 2 2>&1 | tail -1
done

#!/bin/bash
# Resident CTAs / staging size of the energy-only lean sweep: times the shipped library and every variants/libclm_e*.so (built on
# the build box with CLM_NVCC_EXTRA="-DCLM_N3E_MINB_F32=.. -DCLM_N3E_STAGE_BYTES_F32=.." CLM_SO=variants/libclm_eN.so) on the C2 workload.
python tools/time_lj_energy.py 100 both 2>&1 | grep -v "^$"
for so in variants/libclm_e*.so; do
  CLM_SO=$PWD/$so timeout 300 python tools/time_lj_energy.py 100 both 2>&1 | grep "n3=1"
done

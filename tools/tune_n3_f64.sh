python tools/time_n3.py 100 f64 2>&1 | grep "n3=1\|n3=0"
for so in variants/libclm_*.so; do CLM_SO=$PWD/$so timeout 300 python tools/time_n3.py 100 f64 2>&1 | grep "n3=1"; done

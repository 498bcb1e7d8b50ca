#!/usr/bin/env python
"""profiles/r2_traffic.json from one `ncu --set full` report: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum) of the kernels bench.py quotes a roofline for, tied to the CUDA sources by bench.src_hash().

  python tools/traffic_from_ncu.py gpurun_out/prof_xxx.ncu-rep [more reports ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = {"k_sweep_n3<float": "k_sweep_n3_f32", "k_sweep_n3<double": "k_sweep_n3_f64", "k_sweep<float, 2, FLJ": "k_sweep_all_lj_f32",
        "k_sweep<double, 2, FLJ": "k_sweep_all_lj_f64", "k_bin<float": "k_bin_f32", "k_place<float": "k_place_f32", "k_force_finish<float": "k_force_finish_f32"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"report": [], "src_hash": bench.src_hash(), "kernels": {}, "unit": "bytes per launch"}
for rep in sys.argv[1:]:
    out["report"].append(os.path.basename(rep))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in rows[2:]:
        for pat, key in KEYS.items():
            if pat in r[kn].replace("clm::", ""):
                v = float(r[rd].replace(",", "")) * UNIT[units[rd]] + float(r[wr].replace(",", "")) * UNIT[units[wr]]
                out["kernels"].setdefault(key, v)
out["report"] = ", ".join(out["report"])
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

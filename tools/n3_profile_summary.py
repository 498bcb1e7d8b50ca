#!/usr/bin/env python
"""profiles/r2_n3_lean_sweep_lj_f32.txt: source-level split of k_sweep_n3<float, MODE_HALF, N3LJ> from one ncu --set full
--import-source on capture of tools/prof_c2.py (joined with nvdisasm line info by tools/sass_lines.py).
  python tools/n3_profile_summary.py gpurun_out/prof_r2_step.ncu-rep"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rep = sys.argv[1]
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools/sass_lines.py"), rep, os.path.join(ROOT, "celllistmap.jl_b200/libclm_b200.so"),
                      "k_sweep_n3IfLi0ENS_4N3LJIfLb1ELb1", "4000"], capture_output=True, text=True).stdout.splitlines()
src = open(os.path.join(ROOT, "celllistmap.jl_b200/csrc/clm_sweep_n3.cuh")).read().splitlines()
def line_of(marker):
    return next(i + 1 for i, l in enumerate(src) if marker in l)
L_CHUNK, L_CULL, L_ROWS, L_PASS = line_of("auto chunk = [&]"), line_of("auto cull_step = [&]"), line_of("for (int rb = 0; rb < nrows_st"), line_of("for (int c0 = 0; c0 < total;)")
L_PAIR0, L_PAIR1 = line_of("for (int ii = 0; ii < IH; ++ii)"), line_of("if (any_nonzero(fjx, fjy, fjz))")
L_CULLLOOP, L_SWEEP = line_of("far-away dummies round the staged records"), line_of("auto sweep_chunks = [&]")
L_TILE, L_REDUCE, L_RBASE = line_of("const Tile tl = a.tiles[t];"), line_of("auto reduce_fi = [&]"), line_of("for (int rbase = rfa;")
L_END = line_of("if constexpr (NH == 1) reduce_fi(0);")
# helper blocks in front of the kernel: reduction / slot helpers, the pair-force functors, the vector load of particle i
L_RED0, L_FUNC0, L_FUNC1 = line_of("__device__ __forceinline__ void red_add3(float* p"), line_of("// ---- pair-force functors"), line_of("template <class T> struct Vec4S;")
L_VEC0, L_VEC1 = L_FUNC1, line_of("enum { N3_PLAIN = 0")
keys = ["pair loop (8 i-steps per 32-partner chunk)", "cull + in-place compaction", "staging: bulk-copy issue + mbarrier", "chunk loop: partner load, key, RED flush, carry",
        "row classification, prefixes, pass control", "tile setup (tile fetch, cells of the tile, keys, bounding box)", "f_i reduce-scatter + RED, energy fold",
        "other (shuffle / vote intrinsics, address arithmetic)"]
def cat(f, l):
    if f == "clm_sweep.cuh":
        txt = open(os.path.join(ROOT, "celllistmap.jl_b200/csrc/clm_sweep.cuh")).read().splitlines()[l - 1]
        if "xfma<float>" in txt or "fast_rcp<float>" in txt: return 0
        if "cp.async.bulk" in txt or "mbarrier" in txt: return 2
        if "tile_min" in txt or "tile_max" in txt: return 5
        if "float4 v = *reinterpret_cast<const float4*>" in txt: return 3
        return 7
    if f == "clm_common.cuh": return 3
    if f == "cmath": return 1
    if f == "device_atomic_functions.hpp": return 5
    if f != "clm_sweep_n3.cuh": return 7
    if L_FUNC0 <= l < L_FUNC1 or L_VEC0 <= l < L_VEC1 or L_PAIR0 <= l < L_PAIR1: return 0
    if L_CULL <= l < L_ROWS or L_CULLLOOP <= l < L_SWEEP - 12: return 1
    if L_PASS <= l < L_CULLLOOP: return 2
    if L_CHUNK <= l < L_PAIR0 or L_PAIR1 <= l < L_CULL or L_SWEEP - 12 <= l < L_END or L_RED0 <= l < L_FUNC0: return 3
    if L_ROWS <= l < L_PASS: return 4
    if L_REDUCE <= l < L_RBASE or l >= L_END: return 6
    if l < L_REDUCE or L_RBASE <= l < L_CHUNK: return 5
    return 7
cats = {k: 0 for k in keys}
tot = 0
for r in out[3:]:
    m = re.match(r"(\S+):(\d+)\s+(\d+)\s+(\d+)", r)
    if not m: continue
    n = int(m.group(4)); cats[keys[cat(m.group(1), int(m.group(2)))]] += n; tot += n
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
import csv, io
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
kr = next(r for r in rows[2:] if "k_sweep_n3" in r[h.index("Kernel Name")])
g = lambda k: kr[h.index(k)]
ntiles = 126384
L = ["# Newton's-third-law force sweep k_sweep_n3<float, MODE_HALF, N3LJ<float,true,true>> on the C2 workload (1 M argon-density",
     "# particles, cutoff 12 A), FINAL round-2 kernel (sources %s).  Source: %s" % (bench.src_hash(), rep),
     "# (ncu --set full --clock-control none --import-source on -k regex:... python tools/prof_c2.py 100 f32 4), joined with",
     "# nvdisasm -g line info by tools/sass_lines.py; written by tools/n3_profile_summary.py.  Warm CUDA-event time of the same kernel: 0.449-0.452 ms.",
     "#",
     "# gpu__time_duration %s us (cold, under ncu); smsp__inst_executed %.4g; issue active %.1f %%; %s registers, 20 warps / SM;" % (g("gpu__time_duration.sum"), float(g("smsp__inst_executed.sum")), float(g("smsp__issue_active.avg.pct_of_peak_sustained_active")), g("launch__registers_per_thread")),
     "# dram__bytes_read %s + write %s (units of the report) per launch (profiles/r2_traffic.json); thread instructions per warp instruction %s." % (g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), g("smsp__thread_inst_executed_per_inst_executed.ratio")),
     "#", "# Where the instructions go (%.4g warp instructions, %d tiles):" % (tot, ntiles)]
for k, v in sorted(cats.items(), key=lambda kv: -kv[1]):
    L.append("  %-62s %11d  %5.2f%%  %7.1f / tile" % (k, v, 100 * v / tot, v / ntiles))
L += ["#", "# Reading: the pair loop (23 instructions per warp step: LDS.128 of particle i, 3 FADD, FMUL + 2 FFMA, FSETP, zero + predicated",
      "# MUFU.RCP, 7 for the sigma-normalised energy + force scalar, 6 FFMA for f_i and f_j) is %.0f %% of the kernel; the largest single" % (100 * cats[keys[0]] / tot),
      "# item outside it is the divergent UBLKCP issue loop (the cp.async.bulk line of clm_sweep.cuh: ELECT, 5 R2UR.BROADCAST, 2 PLOP3, UBLKCP,",
      "# BRA per copy, ~21 copies per tile).  CTA size / register cap variants: profiles/r2_tune_n3_cta.txt (more resident warps at the same",
      "# register count change nothing: the kernel is bound by the instructions it issues).", "#"]
open(os.path.join(ROOT, "profiles/r2_n3_lean_sweep_lj_f32.txt"), "w").write("\n".join(L + out[:63]) + "\n")
print("\n".join(L))

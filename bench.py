#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the cutoff-pair hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference arm: CPU restatement on the host cores)

A "step" is one pass of the hot path over one batch of synthetic input: new positions -> UpdateCellList!
(cell-list build) -> pairwise!(LJ energy + forces) -> outputs.  Workload at N = 1: BASELINE.json configs[1]
(1M argon-density particles, cubic PBC, cutoff 12 A, Float32; the Float64 figure rides along in `f64`).
metric = in-cutoff pair evaluations per second (pairs with d2 <= cutoff^2, each counted once, as in the reference).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import workloads as W  # noqa: E402


def host_threads():
    """threads the CPU arms use: every core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its children;
    the OpenMP runtime of the oracle reads the variable when the library is loaded, so it is set here, before the import."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def src_hash():
    """hash of the CUDA sources the shipped library was built from: ties profiles/r2_traffic.json to a binary"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "celllistmap.jl_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]


def measured_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_key` from the committed ncu capture
    (profiles/r2_traffic.json, written by tools/traffic_from_ncu.py from one `ncu --set full` run of tools/prof_c2.py)"""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None, "no profiles/r2_traffic.json"
    t = json.load(open(p))
    v = t.get("kernels", {}).get(kernel_key)
    note = f"profiles/r2_traffic.json ({t.get('report')}; sources {t.get('src_hash')}" + ("" if t.get("src_hash") == src_hash() else f", CURRENT sources {src_hash()}: capture is from an older build") + ")"
    return v, note

METRIC = "cutoff pair-evals/s (1M LJ forces)"
UNIT = "pair-evals/s"
# 8 flops per stencil candidate (distance test) + 20 per in-cutoff pair (LJ energy + forces)
FLOPS_PER_CANDIDATE, FLOPS_PER_PAIR_LJ = 8.0, 20.0


def reference_candidates(x, sides, cutoff):
    """C_st: candidate pairs of the REFERENCE's own stencil on the REFERENCE's own grid for an orthorhombic self-set
    system in 2-D or 3-D (forward half of the 8 / 26 neighbour cells + same-cell upper triangle; ghost cells hold the
    periodic images of the real cells), from the cell histogram."""
    sides = np.atleast_1d(np.asarray(sides, np.float64))
    dim = x.shape[1]
    if sides.size == 1:
        sides = np.full(dim, float(sides[0]))
    m = np.floor(sides / cutoff).astype(np.int64)
    cs = sides / m
    c = np.floor(np.mod(x.astype(np.float64), sides) / cs).astype(np.int64) % m
    lin = c[:, 0]
    for k in range(1, dim):
        lin = lin * m[k] + c[:, k]
    h = np.bincount(lin, minlength=int(np.prod(m))).reshape(tuple(m)).astype(np.float64)
    total = (h * (h - 1) / 2).sum()
    import itertools
    for d in itertools.product((-1, 0, 1), repeat=dim):
        if d > (0,) * dim:   # forward half of the neighbours
            total += (h * np.roll(h, tuple(-v for v in d), axis=tuple(range(dim)))).sum()
    return total


class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line):
    one streaming `nvidia-smi -lms 20` child, started before and killed after the region."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device=0):
        self.device, self.proc = device, None

    def start(self):
        fields = "clocks.sm,clocks.max.sm," + ",".join("clocks_event_reasons." + n for n in self.NAMES)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={fields}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        time.sleep(0.25)   # let the first samples arrive before the timed region starts

    def stop(self):
        samples, reasons, mx = [], set(), None
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=10)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.splitlines():
                r = [t.strip() for t in line.split(",")]
                try:
                    samples.append(float(r[0]))
                    mx = float(r[1])
                except (ValueError, IndexError):
                    continue
                for n, v in zip(self.NAMES, r[2:]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(samples) if samples else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(samples)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6553.6, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------
def cpu_port_run(w, dtype, reps, threads):
    """the oracle (CPU restatement of the reference's algorithm incl. projection filter and batch-private outputs)
    on the host cores: full workload, `reps` repetitions of build + map; returns (median seconds, pairs, threads, energy,
    forces of the last repetition)."""
    from oracle import oracle as om
    nt = threads
    x = np.ascontiguousarray(w["x"], dtype=dtype)
    times, e, f = [], None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        o = om.Oracle(x, w["cutoff"], unitcell=w["unitcell"].astype(dtype), dtype=dtype)   # UpdateCellList!
        e, f = o.lj(w["c6"], w["c12"], forces=True, nbatches=nt)                              # pairwise!
        times.append(time.perf_counter() - t0)
        del o
    o = om.Oracle(x, w["cutoff"], unitcell=w["unitcell"].astype(dtype), dtype=dtype)
    npairs = o.sum_d_d2(nbatches=nt)[2]
    return statistics.median(times), npairs, nt, e, f


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Julia and no Julia
    toolchain exists here or on the GPU box, so this is the oracle port with all host threads (the thread count is
    taken from the CPU affinity mask: torchrun's OMP_NUM_THREADS=1 is overridden before the OpenMP runtime loads).
    N = 1: the full 1M-particle C2 workload.  N > 1 (the GPU arm runs the slab-decomposed C5 system, 8M particles per
    GPU): a bounded sample of that system -- the same density, cutoff and generator at 2M particles -- because the metric
    is a throughput (in-cutoff pair evaluations per second) and one 64M-particle CPU step takes about a minute."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nt = host_threads()
    dtype = np.float32
    from oracle import oracle as om
    nt = om.lib().ora_num_threads()
    if args.gpus > 1:
        import bench_multi
        nside = 126
        x, uc = bench_multi.slab_lattice(0, 1, nside, nside, dtype)
        w = dict(x=x, unitcell=uc, cutoff=12.0, c6=W.ARGON_C6, c12=W.ARGON_C12)
        workload = ("C5 sample: LJ energy+forces, argon-density particles, cubic PBC, cutoff 12 A (BASELINE.json configs[4]): "
                    f"{nside}^3 = {nside ** 3} particles of the generator the GPU arm shards over {args.gpus} GPUs")
    else:
        nside = args.cpu_nside
        w = W.c2_argon(nside, dtype)
        workload = "C2: LJ energy+forces, 1M argon-density particles, cubic PBC, cutoff 12 A (BASELINE.json configs[1])"
    x, uc = w["x"], w["unitcell"]

    def step():
        o = om.Oracle(x, w["cutoff"], unitcell=uc, dtype=dtype)
        o.lj(w["c6"], w["c12"], forces=True, nbatches=nt)
        return o

    # bounded: the whole --steps K --warmup W run should end within a few minutes.  One probe step on the full sample; when the
    # projected total exceeds ~3 minutes the sample shrinks (same generator, density and cutoff; the metric is a throughput)
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    budget = 180.0
    total_steps = args.steps + max(args.warmup - 1, 0)
    if t1 * total_steps > budget:
        shrink = (budget / (t1 * total_steps)) ** (1.0 / 3.0)
        nside = max(48, int(nside * shrink))
        if args.gpus > 1:
            x, uc = bench_multi.slab_lattice(0, 1, nside, nside, dtype)
        else:
            wz = W.c2_argon(nside, dtype)
            x, uc = wz["x"], wz["unitcell"]
            workload += f" -- CPU arm bounded to {nside}^3 particles of the same generator (K + W = {args.steps + args.warmup} steps within ~3 minutes)"
    for _ in range(max(args.warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = step()
    dt = time.perf_counter() - t0
    npairs = o.sum_d_d2(nbatches=nt)[2]
    value = npairs * args.steps / dt
    sample = (f"{nside}^3 = {nside ** 3} argon-density particles (same generator/density/cutoff as the GPU arm), build + LJ energy+forces per step, "
              f"{nt} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "n_particles": nside ** 3,
                   "note": "CPU restatement of CellListMap.jl (C++/OpenMP oracle port, not Julia: no Julia toolchain on the box)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nt, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line, default=lambda o: o.item() if hasattr(o, "item") else str(o)))


# ----------------------------------------------------------------------------------------------------------
def c5_single_gpu(clm, torch, local, stream, flush_buf, nside, nx, steps=5):
    """the C5 particle system (argon-density lattice, nx x nside x nside sites, Float32) on ONE GPU through a plain handle:
    the denominators of the weak-scaling (8M particles per GPU) and strong-scaling (64M particles) figures of the N > 1 runs"""
    import bench_multi
    dev = torch.device("cuda", local)
    x_dev, uc = bench_multi.slab_lattice_torch(0, 1, nside, nx, np.float32, dev)
    torch.cuda.synchronize()   # generated on torch's default stream, consumed on the bench stream
    n = x_dev.shape[0]
    h = clm.Handle(3, np.float32, device=local)
    h.set_stream(stream.cuda_stream)
    h.set_box(clm._capi.ORTHORHOMBIC, uc, 12.0, 1)
    e_dev = torch.zeros(1, dtype=torch.float32, device=dev)
    f_dev = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    with torch.cuda.stream(stream):
        h.set_positions(0, x_dev)
        h.build()
        sd, sd2, npairs = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.int64)
        h.map_sum_d_d2(sd, sd2, npairs)
        for _ in range(3):
            h.set_positions(0, x_dev)
            h.map_lj(W.ARGON_C6, W.ARGON_C12, e_dev, f_dev)
        evs = []
        for _ in range(steps):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            h.set_positions(0, x_dev)
            h.map_lj(W.ARGON_C6, W.ARGON_C12, e_dev, f_dev)
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
        sw, bd = [], []
        for _ in range(3):
            h.set_positions(0, x_dev)
            h.map_lj(W.ARGON_C6, W.ARGON_C12, e_dev, f_dev, profile=True)
            st = h.stats()
            sw.append(st.sweep_ms)
            bd.append(st.build_ms)
    out = {"n_particles": int(n), "in_cutoff_pairs": int(npairs[0]), "ms_per_step": ms, "sweep_kernel_ms": statistics.mean(sw), "build_ms": statistics.mean(bd),
           "value": int(npairs[0]) / (ms * 1e-3), "energy": float(e_dev[0]), "step": "update positions (D2D) + UpdateCellList! + pairwise!(LJ energy+forces), device-resident"}
    h.close()
    del x_dev, f_dev
    torch.cuda.empty_cache()
    return out


def other_configs(clm, torch, local, fp64_peak_tf):
    """BASELINE.json configs 3 and 4 at full size (parity-test configurations: extra keys, not the metric): device time of
    the map (CUDA events around the sweep, build separately) and the end-to-end call with HOST arrays in and out; the work
    model of the roofline is SURVEY.md's 8 flops per reference-stencil candidate + F_f per in-cutoff pair, candidates from
    the uniform density (27 (9) reference cells around every particle) against the FP64 FMA peak measured live."""
    out = {}

    def timed(fn, h, reps):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        st = h.stats()
        return 1e3 * statistics.median(ts), st.build_ms, st.sweep_ms

    w = W.c3_triclinic_cross(1_000_000, 1_000_000)
    h = clm.Handle(3, np.float64, device=local)
    h.set_box(clm._capi.TRICLINIC, w["unitcell"], w["cutoff"], 1)
    i, j, d = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1)
    sd, sd2, npr = np.zeros(1), np.zeros(1), np.zeros(1, np.int64)

    def c3_e2e():
        h.set_positions(0, w["x"])
        h.set_positions(1, w["y"])
        h.map_mindist(i, j, d, profile=True)

    c3_e2e()
    h.map_sum_d_d2(sd, sd2, npr)
    t, b, s = timed(c3_e2e, h, 5)
    box = h.get_box()
    vcell = float(np.prod([box.cell_size[k] for k in range(3)]))
    cand = 1_000_000 * 27.0 * vcell * (1_000_000 / abs(np.linalg.det(w["unitcell"].astype(np.float64))))   # x particles x (27 cells x density of y)
    flops = 8.0 * cand + 2.0 * int(npr[0])
    out["c3_triclinic_cross_mindist_f64"] = {
        "config": "configs[2] at 1M x 1M particles (the 2M-particle system), triclinic cell, cutoff 12 A, Float64", "pairs": int(npr[0]),
        "sweep_kernel_ms": s, "build_ms": b, "e2e_ms": t, "value": int(npr[0]) / ((s + b) * 1e-3), "e2e_value": int(npr[0]) / (t * 1e-3),
        "roofline": {"bound": "fp64", "achieved": flops / (s * 1e-3) / 1e12, "peak": fp64_peak_tf, "unit": "TFLOP/s", "frac": flops / (s * 1e-3) / 1e12 / fp64_peak_tf,
                     "algorithmic_flops_per_launch": flops, "work_model": "8 flops x 27 reference cells of candidates per x particle + 2 per in-cutoff pair"}}
    h.close()
    for dim in (3, 2):
        w = W.c4_galaxies(4_000_000, dim)
        h = clm.Handle(dim, np.float64, device=local)
        h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
        c, sm = np.zeros(5, np.int64), np.zeros(5)

        def c4_e2e():
            h.set_positions(0, w["x"])
            h.map_pairvel(w["v"], None, w["rbins"], c, sm, profile=True)

        c4_e2e()
        t, b, s = timed(c4_e2e, h, 3)
        npairs = int(c.sum())
        box = h.get_box()
        vcell = float(np.prod([box.cell_size[k] for k in range(dim)]))
        n = w["x"].shape[0]
        cand = 0.5 * n * (3 ** dim) * vcell * (n / float(w["L"]) ** dim)
        flops = 8.0 * cand + 12.0 * npairs
        out[f"c4_pairwise_velocities_{dim}d_f64"] = {
            "config": f"configs[3]: mean pairwise velocity histogram, 4M galaxies, {dim}-D, cutoff 5, 5 bins, Float64", "pairs": npairs,
            "sweep_kernel_ms": s, "build_ms": b, "e2e_ms": t, "value": npairs / ((s + b) * 1e-3), "e2e_value": npairs / (t * 1e-3),
            "roofline": {"bound": "fp64", "achieved": flops / (s * 1e-3) / 1e12, "peak": fp64_peak_tf, "unit": "TFLOP/s", "frac": flops / (s * 1e-3) / 1e12 / fp64_peak_tf,
                         "algorithmic_flops_per_launch": flops, "work_model": f"8 flops x half of {3 ** dim} reference cells of candidates per particle + 12 per in-cutoff pair"}}
        h.close()
    return out


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nside", type=int, default=100, help="particles = nside^3 (100 -> the 1M-particle C2 config)")
    ap.add_argument("--cpu-nside", type=int, default=100, help="size of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f64", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the extra keys (C5 on one GPU, configs 3 and 4)")
    ap.add_argument("--no-64m", action="store_true", help="skip the 64M-particle single-GPU run (strong-scaling denominator)")
    ap.add_argument("--no-nl", action="store_true", help="N > 1: skip the neighbour-list build timing")
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--multi-nside", type=int, default=400, help="N > 1: lattice sites along y and z")
    ap.add_argument("--multi-nx-per-rank", type=int, default=50, help="N > 1: lattice planes along x per rank (50 x 400 x 400 = 8M particles per GPU)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    nthreads = host_threads()   # before anything loads an OpenMP runtime

    import torch
    import torch.distributed as dist
    import celllistmap_b200 as clm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1 or args.workload == "c5":
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
        import bench_multi
        return bench_multi.run(args, rank, world, local)

    dev = torch.device("cuda", local)
    peaks, peaks_src = load_peaks()
    stream = torch.cuda.Stream(device=dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def measure(dtype, steps, warmup, sample_clocks):
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        w = W.c2_argon(args.nside, dtype)
        n = w["x"].shape[0]
        h = clm.Handle(3, dtype, device=local)
        h.set_stream(stream.cuda_stream)
        h.set_box(clm._capi.ORTHORHOMBIC, w["unitcell"], w["cutoff"], 1)
        # ---- device-resident arm: inputs already in HBM when the timed region starts ----
        x_dev = torch.from_numpy(w["x"]).to(dev)
        e_dev = torch.zeros(1, dtype=tdt, device=dev)
        f_dev = torch.zeros((n, 3), dtype=tdt, device=dev)
        torch.cuda.synchronize()

        def step_dev(profile=False):
            h.set_positions(0, x_dev)                       # update!(sys; xpositions): owning copy, D2D
            h.map_lj(w["c6"], w["c12"], e_dev, f_dev, reset=True, profile=profile)   # pairwise!: UpdateCellList! + map

        with torch.cuda.stream(stream):
            # pair count, at-cutoff band and reference-stencil candidates (work model), outside the timed region
            h.set_positions(0, x_dev)
            h.build()
            sd, sd2, npairs = np.zeros(1, dtype), np.zeros(1, dtype), np.zeros(1, np.int64)
            h.map_sum_d_d2(sd, sd2, npairs)
            P_in = int(npairs[0])
            band = int(h.stats().n_cutoff_band)
            for _ in range(warmup):
                step_dev()
            stream.synchronize()
            sampler = ClockSampler(local) if sample_clocks else None
            if sampler:
                sampler.start()
            l0 = h.stats().launches
            evs = []
            torch.cuda.synchronize()
            for _ in range(steps):
                flush_buf.zero_()                           # L2 flush between timed iterations (untimed)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                step_dev()
                b.record(stream)
                evs.append((a, b))
            torch.cuda.synchronize()
            evs = [a.elapsed_time(b) for a, b in evs]
            launches = (h.stats().launches - l0) / steps
            clocks = sampler.stop() if sampler else None
            # kernel-level durations (CUDA events around the sweep launch / the build) from a separate short profiled loop
            sweep_ms, build_ms, map_ms = [], [], []
            for _ in range(10):
                flush_buf.zero_()
                step_dev(profile=True)
                st = h.stats()
                sweep_ms.append(st.sweep_ms)
                build_ms.append(st.build_ms)
                map_ms.append(st.map_ms)
            dev_ms = sum(evs) / steps
            f_gpu = f_dev.cpu().numpy()
            e_gpu = float(e_dev[0])
            # ---- end-to-end arms: HOST buffers through the C ABI, H2D + D2H inside the timed region ----
            # (a) synchronous calls: clm_set_positions + clm_map_lj return with the outputs in host memory
            x_pin = [torch.from_numpy(w["x"]).pin_memory() for _ in range(2)]
            f_pin = [torch.zeros((n, 3), dtype=tdt).pin_memory() for _ in range(2)]
            e_pin = [torch.zeros(1, dtype=tdt).pin_memory() for _ in range(2)]

            def step_sync(k):
                h.set_positions(0, x_pin[k & 1].numpy())    # H2D from pinned host memory
                h.map_lj(w["c6"], w["c12"], e_pin[k & 1].numpy(), f_pin[k & 1].numpy(), reset=True)   # forces + energy D2H, synchronous on return

            for k in range(3):
                step_sync(k)
            nsync = max(5, steps // 4)
            e2e = []
            for k in range(nsync):
                flush_buf.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                step_sync(k)
                torch.cuda.synchronize()
                e2e.append(time.perf_counter() - t0)
            e2e_sync_ms = 1e3 * sum(e2e) / len(e2e)
            # The pipelined arms take their frames from a RING of distinct host frames whose total size exceeds the L2 (inputs
            # larger than L2: frame r holds the same particles in a rotated order, so every frame has the same pair set and
            # in-cutoff pair count; by the time a frame is reused, 200 MB of other frames have gone through the GPU).  Every
            # frame's positions are copied from pinned host memory and its forces + energy copied back inside the timed
            # region.  Each arm is timed twice: as it is (`fl` False) and with a 256 MiB memset (L2 flush) on the compute
            # stream in front of every frame, INSIDE the timed region (the round-2 figure: it pays 0.07 ms of memset per
            # frame and evicts the positions the copy-in has just delivered).
            nh = 2
            xbytes = int(w["x"].nbytes)
            nring = -(-int(200e6) // xbytes)
            nring = -(-nring // (2 * nh)) * (2 * nh)
            roll = n // nring
            xring = [torch.from_numpy(np.ascontiguousarray(np.roll(w["x"], r * roll, axis=0))).pin_memory() for r in range(nring)]
            # (b) pipelined frames through ONE handle (clm_set_positions_async + CLM_ASYNC): the copy-in of frame k+1, the
            # compute of frame k and the copy-out of frame k-1 overlap on three streams.
            def step_pipe(k, fl):
                if fl:
                    flush_buf.zero_()
                h.set_positions_async(0, xring[k % nring].numpy())
                h.map_lj(w["c6"], w["c12"], e_pin[k & 1].numpy(), f_pin[k & 1].numpy(), async_=True)

            def time_pipe(fl):
                for k in range(4):
                    step_pipe(k, False)
                h.synchronize()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(steps):
                    step_pipe(k, fl)
                h.synchronize()
                torch.cuda.synchronize()
                return 1e3 * (time.perf_counter() - t0) / steps

            e2e_one_ms, e2e_one_flushed_ms = time_pipe(False), time_pipe(True)
            # (c) the same frames through clm.FramePipeline: TWO handles take the frames in turn, each with its own three streams,
            # so that the cell-list build of frame k+1 (short latency-bound kernels) runs next to the pair sweep of frame k (4 resident
            # sweep CTAs per SM instead of 5 leave the room).
            pstreams = [torch.cuda.Stream(device=dev) for _ in range(nh)]
            pipe = clm.FramePipeline(3, dtype, w["unitcell"], w["cutoff"], handles=nh, device=local, streams=[st.cuda_stream for st in pstreams])
            fq = [torch.zeros((n, 3), dtype=tdt).pin_memory() for _ in range(2 * nh)]
            eq = [torch.zeros(1, dtype=tdt).pin_memory() for _ in range(2 * nh)]

            def frame(k, fl):
                a = pipe.next_handle()
                q = a * 2 + ((k // nh) & 1)
                with torch.cuda.stream(pstreams[a]):
                    if fl:
                        flush_buf.zero_()
                    pipe.submit_lj(w["c6"], w["c12"], xring[k % nring].numpy(), eq[q].numpy(), fq[q].numpy())
                return q

            def time_frames(fl):
                for k in range(4 * nh):
                    frame(k, False)
                pipe.synchronize()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(steps):
                    q = frame(k, fl)
                pipe.synchronize()
                torch.cuda.synchronize()
                return 1e3 * (time.perf_counter() - t0) / steps, q

            e2e_flushed_ms, _ = time_frames(True)
            e2e_ms, qlast = time_frames(False)
            e_pipe = float(eq[qlast][0])
            # (d) device-resident throughput of independent steps over the same two handles (positions and outputs stay in HBM;
            # the L2 flush in front of every step is INSIDE the timed region here, since steps overlap)
            fdd = [torch.zeros((n, 3), dtype=tdt, device=dev) for _ in range(nh)]
            edd = [torch.zeros(1, dtype=tdt, device=dev) for _ in range(nh)]

            def step_two(k, fl):
                a = k % nh
                with torch.cuda.stream(pstreams[a]):
                    if fl:
                        flush_buf.zero_()
                    pipe.handles[a].set_positions(0, x_dev)
                    pipe.handles[a].map_lj(w["c6"], w["c12"], edd[a], fdd[a], reset=True)

            for k in range(2 * nh):
                step_two(k, False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(steps):
                step_two(k, True)
            torch.cuda.synchronize()
            dev_two_ms = 1e3 * (time.perf_counter() - t0) / steps
            # the last frame holds the particles of the device-resident arm in a rotated order: its forces, rotated back, agree
            # with that arm's result to the rounding of the order-free reductions
            fa = np.roll(fq[qlast].numpy().astype(np.float64), -(((steps - 1) % nring) * roll), axis=0)
            fb = f_gpu.astype(np.float64)
            f_pipe_diff = float(np.abs(fa - fb).max() / np.abs(fb).max())
            pipe.close()
        res = dict(P_in=P_in, band=band, n=n, dev_ms=dev_ms, sweep_ms=statistics.mean(sweep_ms), build_ms=statistics.mean(build_ms), map_ms=statistics.mean(map_ms),
                   e2e_ms=e2e_ms, e2e_flushed_ms=e2e_flushed_ms, e2e_one_ms=e2e_one_ms, e2e_one_flushed_ms=e2e_one_flushed_ms, nring=nring, dev_two_ms=dev_two_ms, e2e_sync_ms=e2e_sync_ms, launches=launches, clocks=clocks, w=w, energy=e_gpu, forces=f_gpu, energy_pipe=e_pipe,
                   h2d=int(x_pin[0].numpy().nbytes), d2h=int(f_pin[0].numpy().nbytes + e_pin[0].numpy().nbytes), stats=h.stats(), frames_diff=f_pipe_diff)
        h.close()
        return res

    def neighborlist_ms():
        """BASELINE metric 2: neighbour-list build time of configs[0] (10k particles, unit cube, cutoff 0.1, Float64):
        end to end through the reference-facing API (host positions in, host records out) and device-resident."""
        w = W.c1_neighborlist()
        nb = clm.InPlaceNeighborList(x=w["x"], cutoff=w["cutoff"], unitcell=w["unitcell"], device=local)
        hh = nb.sys._h
        for _ in range(3):
            clm.update(nb, xpositions=w["x"])
            lst = nb.neighborlist()
        te = []
        for _ in range(20):
            t0 = time.perf_counter()
            clm.update(nb, xpositions=w["x"])
            lst = nb.neighborlist()
            te.append(time.perf_counter() - t0)
        xd = torch.from_numpy(w["x"]).to(dev)
        td = []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hh.set_positions(0, xd)
            hh.neighborlist_count()
            td.append(time.perf_counter() - t0)
        return {"config": "configs[0]: 10k random 3-D particles, orthorhombic unit cube, cutoff 0.1, Float64", "pairs": int(len(lst)),
                "at_cutoff_band_pairs": int(nb.n_cutoff_band),
                "e2e_ms": 1e3 * statistics.median(te), "device_resident_ms": 1e3 * statistics.median(td),
                "reference_published_ms": 7.978, "reference_source": "src/API/neighborlist.jl:204-207 (serial Julia, unstated CPU)"}

    r32 = measure(np.float32, args.steps, args.warmup, True)
    r64 = None if args.no_f64 else measure(np.float64, max(3, args.steps // 4), 3, False)

    C_st = reference_candidates(r32["w"]["x"], r32["w"]["L"], r32["w"]["cutoff"])
    F_alg = FLOPS_PER_CANDIDATE * C_st + FLOPS_PER_PAIR_LJ * r32["P_in"]
    # FP32 SIMT peak: 148 SMs x 128 lanes x 2 flop x max SM clock (no FP32-pipe figure in MEASURED_PEAKS.json:
    # nominal from the measured max clock; B200_PROFILING.md fallback)
    n_sm = r32["stats"].n_sm
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_nominal_tf = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12
    fp32_peak_tf = clm._capi.measure_fma_peak(np.float32, local)      # measured live: register-resident FMA loop
    fp64_peak_tf = clm._capi.measure_fma_peak(np.float64, local)
    achieved_tf = F_alg / (r32["sweep_ms"] * 1e-3) / 1e12
    # algorithmic HBM bytes of the same launch: records in (16 B / particle incl. images) + forces out (12 B / particle)
    B_alg = 16.0 * r32["stats"].n_total[0] + 12.0 * r32["n"]
    kernel_name = "k_sweep_n3<float, MODE_HALF, N3LJ<float,true,true>>"
    traffic, traffic_note = measured_traffic("k_sweep_n3_f32")
    line = {
        "metric": METRIC, "value": r32["P_in"] / (r32["dev_ms"] * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r32["dev_ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: LJ energy+forces, 1M argon-density particles, cubic PBC, cutoff 12 A (BASELINE.json configs[1])",
                   "n_particles": r32["n"], "in_cutoff_pairs": r32["P_in"], "at_cutoff_band_pairs": r32["band"], "reference_stencil_candidates": C_st,
                   "step": "update positions (D2D) + UpdateCellList! + pairwise!(LJ energy+forces)",
                   "l2": "`value`: L2 flushed between timed steps (256 MiB memset, untimed), per-step CUDA events on the launching stream; `e2e`: inputs larger "
                         f"than L2 -- a ring of {r32['nring']} distinct host frames ({r32['nring'] * r32['h2d'] / 1e6:.0f} MB > 126 MB L2), every frame copied in from pinned host "
                         "memory inside the timed region, no flush; `e2e_flushed`: the same with a 256 MiB memset in front of every frame INSIDE the timed region"},
        "clocks": r32["clocks"],
        "e2e": {"value": r32["P_in"] / (r32["e2e_ms"] * 1e-3), "unit": UNIT, "ms_per_step": r32["e2e_ms"],
                "h2d_bytes_per_step": r32["h2d"], "d2h_bytes_per_step": r32["d2h"],
                "mode": "independent frames through clm.FramePipeline: two handles (particle systems) take the frames in turn, each through the C ABI's pipelined "
                        "calls (clm_set_positions_async + clm_map_lj with CLM_ASYNC); every frame's positions are copied from pinned host memory and its forces + "
                        "energy copied back inside the timed region; within a handle copy-in / compute / copy-out of consecutive frames overlap, across handles "
                        "the cell-list build of frame k+1 runs next to the sweep of frame k; the frames come from a ring of distinct host frames larger than the "
                        "L2 (same particles in a rotated order: same pair set); wall clock over all frames",
                "input_ring_frames": r32["nring"], "input_ring_bytes": r32["nring"] * r32["h2d"],
                "last_frame_vs_device_resident_force_max_rel_diff": r32["frames_diff"], "energy": r32["energy_pipe"]},
        "e2e_flushed": {"value": r32["P_in"] / (r32["e2e_flushed_ms"] * 1e-3), "unit": UNIT, "ms_per_step": r32["e2e_flushed_ms"],
                        "mode": "`e2e` with a 256 MiB memset (L2 flush) on the frame's compute stream in front of every frame, inside the timed region (0.07 ms of memset "
                                "per frame; it also evicts the positions the copy-in has just delivered): the round-2 headline figure, kept for comparison"},
        "e2e_one_handle": {"value": r32["P_in"] / (r32["e2e_one_ms"] * 1e-3), "unit": UNIT, "ms_per_step": r32["e2e_one_ms"],
                           "ms_per_step_flushed": r32["e2e_one_flushed_ms"],
                           "mode": "the same ring of frames pipelined through ONE handle (copy-in of frame k+1, compute of frame k, copy-out of frame k-1 overlap; builds and sweeps "
                                   "of consecutive frames do not); ms_per_step_flushed: with the L2 flush in front of every frame inside the timed region"},
        "device_resident_two_handles": {"value": r32["P_in"] / (r32["dev_two_ms"] * 1e-3), "unit": UNIT, "ms_per_step": r32["dev_two_ms"],
                                        "mode": "throughput of independent device-resident steps alternating between two handles (blocks_per_sm = -1): the cell-list build "
                                                "of one step runs next to the sweep of the other; L2 flush in front of every step INSIDE the timed region (0.07 ms of "
                                                "memset per step); wall clock over all steps.  `value` above is the single-handle step with the flush untimed"},
        "e2e_sync": {"value": r32["P_in"] / (r32["e2e_sync_ms"] * 1e-3), "unit": UNIT, "ms_per_step": r32["e2e_sync_ms"],
                     "mode": "one synchronous clm_set_positions + clm_map_lj per step (outputs in host memory on return): the latency of a dependent step"},
        "gpu_launches": r32["launches"] * args.steps,
        "roofline": {"bound": "fp32", "kernel": kernel_name, "achieved": achieved_tf, "peak": fp32_peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak_tf,
                     "traffic": traffic, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload: " + traffic_note,
                     "algorithmic_flops_per_launch": F_alg, "kernel_ms": r32["sweep_ms"], "build_ms": r32["build_ms"],
                     "peak_source": "FP32 FMA peak measured live by clm_measure_fma_peak (register-resident FMA loop, all SMs); "
                                    f"nominal {fp32_nominal_tf:.1f} TFLOP/s = 148 SM x 128 lanes x 2 x sm_max_mhz from {peaks_src}; no tensor cores on this path",
                     "fp64_peak_measured": fp64_peak_tf,
                     "hbm": {"algorithmic_bytes_per_launch": B_alg, "achieved_gbs": B_alg / (r32["sweep_ms"] * 1e-3) / 1e9,
                             "peak_gbs": peaks.get("hbm_gbs")}},
        "breakdown_ms": {"step": r32["dev_ms"], "build": r32["build_ms"], "sweep_kernel": r32["sweep_ms"], "map_call_device": r32["map_ms"]},
        "source_hash": src_hash(),
    }
    if r64:
        F64_alg = FLOPS_PER_CANDIDATE * C_st + FLOPS_PER_PAIR_LJ * r64["P_in"]
        line["f64"] = {"value": r64["P_in"] / (r64["dev_ms"] * 1e-3), "unit": UNIT, "ms_per_step": r64["dev_ms"],
                       "sweep_kernel_ms": r64["sweep_ms"], "build_ms": r64["build_ms"],
                       "e2e_value": r64["P_in"] / (r64["e2e_ms"] * 1e-3), "e2e_sync_value": r64["P_in"] / (r64["e2e_sync_ms"] * 1e-3), "in_cutoff_pairs": r64["P_in"],
                       "roofline_frac_fp64": F64_alg / (r64["sweep_ms"] * 1e-3) / 1e12 / fp64_peak_tf}
    line["neighborlist_build"] = neighborlist_ms()
    if not args.no_extras:
        line["c5_1gpu"] = {"8M": c5_single_gpu(clm, torch, local, stream, flush_buf, args.multi_nside, args.multi_nx_per_rank)}
        if not args.no_64m:
            line["c5_1gpu"]["64M"] = c5_single_gpu(clm, torch, local, stream, flush_buf, args.multi_nside, args.multi_nx_per_rank * 8, steps=3)
        line["other_configs"] = other_configs(clm, torch, local, fp64_peak_tf)
    if not args.no_cpu_baseline:
        from oracle import oracle as om
        nt = om.lib().ora_num_threads()
        wc = W.c2_argon(args.cpu_nside, np.float32)
        t, npairs, nt, e_o, f_o = cpu_port_run(wc, np.float32, 3, nt)
        line["cpu_baseline"] = {"value": npairs / t, "unit": UNIT, "cores": nt, "kind": "port",
                                "sample": f"full workload ({args.cpu_nside}^3 particles), median of 3 x (build + LJ energy+forces), "
                                          "C++/OpenMP restatement of the reference (projection filter, batch-private outputs)",
                                "seconds_per_step": t, "host_threads_available": nthreads}
        if args.cpu_nside == args.nside:
            # parity of the timed GPU result against the CPU restatement on the same input: forces against the oracle run in
            # the same precision (Float32 input -> Float32 arithmetic in the reference), the energy against the oracle in
            # Float64 (the reference's Float32 sum of 7.7e7 signed terms is itself only good to ~1e-2)
            fo = np.asarray(f_o, np.float64)
            o64 = om.Oracle(wc["x"].astype(np.float64), wc["cutoff"], unitcell=wc["unitcell"].astype(np.float64))
            e64 = float(o64.lj(wc["c6"], wc["c12"], forces=False, nbatches=nt))
            del o64
            line["parity"] = {"pairs_equal": bool(npairs == r32["P_in"]), "energy_rel_err": abs(r32["energy"] - e64) / abs(e64),
                              "force_max_rel_err": float(np.abs(r32["forces"].astype(np.float64) - fo).max() / np.abs(fo).max()),
                              "against": "forces: oracle (C++ restatement of the reference) in Float32 on the same input; energy: the oracle in Float64; "
                                         "north_star tolerance 1e-5", "oracle_f32_energy_rel_err": abs(float(e_o) - e64) / abs(e64),
                              "at_cutoff_band_pairs": r32["band"]}
    print(json.dumps(line, default=lambda o: o.item() if hasattr(o, "item") else str(o)))


if __name__ == "__main__":
    main()

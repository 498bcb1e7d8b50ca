"""bench.py's N > 1 arm: the slab-decomposed LJ force map (BASELINE.json configs[4]) over NCCL, one rank per GPU.

Workload (weak scaling): every rank generates ITS slab of the argon-density jittered lattice on its own GPU's host
(nside_x_per_rank x nside x nside sites; 8 ranks x 50 x 400 x 400 = the 64M-particle config).  A step = halo exchange
(NCCL send/recv of the face cell layers) + UpdateCellList! + pairwise!(LJ energy+forces) + all_reduce of the energy;
forces stay sharded with their owners.  Timed per rank with CUDA events on the launching stream, max over ranks."""
import json
import os
import statistics

import numpy as np
import torch
import torch.distributed as dist

import workloads as W


def slab_lattice(rank, world, nside, nx_per_rank, dtype):
    """this rank's planes of the nside_x x nside x nside jittered lattice (nside_x = world * nx_per_rank)."""
    a = W.ARGON_RHO ** (-1.0 / 3.0)
    nx_tot = world * nx_per_rank
    ix = np.arange(rank * nx_per_rank, (rank + 1) * nx_per_rank, dtype=np.int64)
    g = np.arange(nside, dtype=np.int64)
    I, J, K = np.meshgrid(ix, g, g, indexing="ij")
    site = ((I * nside + J) * nside + K).reshape(-1)
    n = site.shape[0]
    sites = np.stack([I.reshape(-1), J.reshape(-1), K.reshape(-1)], 1).astype(np.float64) * a
    # counter-based splitmix64: component c of site s is stream element 3 s + c (same values whatever the rank layout)
    with np.errstate(over="ignore"):
        k = (site[:, None] * 3 + np.arange(3)[None, :] + 1).astype(np.uint64)
        z = np.uint64(W.SEED) + np.uint64(0x9E3779B97F4A7C15) * k
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    x = sites + (u - 0.5) * (0.5 * a) + 0.25 * a
    x = x[W.shuffle_perm(W.SEED + 1 + rank, n)]
    unitcell = np.array([a * nx_tot, a * nside, a * nside], dtype)
    return np.ascontiguousarray(x.astype(dtype)), unitcell


def _splitmix64_torch(seed, k):
    """splitmix64 output for the 1-based counters k (int64 tensor, two's-complement arithmetic = the uint64 arithmetic of
    workloads.splitmix64; logical right shifts emulated by masking the sign extension)."""
    def shr(z, s):
        return (z >> s) & ((1 << (64 - s)) - 1)

    def i64(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v
    z = i64(seed) + i64(0x9E3779B97F4A7C15) * k
    z = (z ^ shr(z, 30)) * i64(0xBF58476D1CE4E5B9)
    z = (z ^ shr(z, 27)) * i64(0x94D049BB133111EB)
    return z ^ shr(z, 31)


def slab_lattice_torch(rank, world, nside, nx_per_rank, dtype, device):
    """slab_lattice generated on `device` with torch (bit-identical values and order): the 64M-particle single-GPU case
    takes a second instead of a minute of numpy."""
    a = W.ARGON_RHO ** (-1.0 / 3.0)
    nx_tot = world * nx_per_rank
    ix = torch.arange(rank * nx_per_rank, (rank + 1) * nx_per_rank, dtype=torch.int64, device=device)
    g = torch.arange(nside, dtype=torch.int64, device=device)
    I, J, K = torch.meshgrid(ix, g, g, indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    site = (I * nside + J) * nside + K
    n = site.shape[0]
    x = torch.empty((n, 3), dtype=torch.float64, device=device)
    for c, idx in enumerate((I, J, K)):
        z = _splitmix64_torch(W.SEED, site * 3 + (c + 1))
        u = ((z >> 11) & ((1 << 53) - 1)).to(torch.float64) * 2.0 ** -53
        x[:, c] = idx.to(torch.float64) * a + (u - 0.5) * (0.5 * a) + 0.25 * a
    del I, J, K, site
    # workloads.shuffle_perm: stable argsort of the uint64 stream (sign bit flipped for the signed sort)
    keys = _splitmix64_torch(W.SEED + 1 + rank, torch.arange(1, n + 1, dtype=torch.int64, device=device)) ^ (-(1 << 63))
    perm = torch.sort(keys, stable=True).indices
    del keys
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    x = x[perm].to(tdt).contiguous()
    unitcell = np.array([a * nx_tot, a * nside, a * nside], dtype)
    return x, unitcell


def run(args, rank, world, local):
    import celllistmap_b200 as clm  # noqa: F401
    from celllistmap_b200 import slab
    import bench
    bench.host_threads()   # torchrun exports OMP_NUM_THREADS=1: the CPU baseline leg uses every core of the affinity mask
    dev = torch.device("cuda", local)
    dtype = np.float32
    nside = args.multi_nside
    nx_per_rank = args.multi_nx_per_rank
    x_host, unitcell = slab_lattice(rank, world, nside, nx_per_rank, dtype)
    cutoff = 12.0
    stream = torch.cuda.Stream(device=dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    s = slab.SlabSystem(unitcell, cutoff, dtype=dtype, device=dev)
    s.set_stream(stream)
    with torch.cuda.stream(stream):
        x_dev = torch.from_numpy(x_host).to(dev)
        # the generator hands every rank its own planes, which is (up to jitter across a face) its slab; migrate strays
        c = s.cell_layers(x_dev).to(torch.int64)
        owner = s.plan.owner_of(c)
        stray = int((owner != rank).sum())
        n_own = x_dev.shape[0]
        f_dev = torch.zeros((n_own, 3), dtype=torch.float32, device=dev)

        def step(profile=False):
            s.update(x_dev)                                  # halo exchange + positions to the engine
            return s.map_lj(W.ARGON_C6, W.ARGON_C12, f_dev, profile=profile)   # build + sweep + energy all_reduce

        if stray:
            raise SystemExit(f"rank {rank}: {stray} generated particles fall outside the rank's slab (jitter across a face)")
        for _ in range(args.warmup):
            step()
        s.update(x_dev)
        sd = s.sum_d_d2()
        P_in = sd[2]                                         # global in-cutoff pairs (all_reduced)
        sampler = bench.ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        l0 = s.h.stats().launches
        torch.cuda.synchronize()
        dist.barrier()
        evs = []
        for _ in range(args.steps):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            e = step()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        dist.barrier()
        times = [a.elapsed_time(b) for a, b in evs]
        launches = s.h.stats().launches - l0
        clocks = sampler.stop() if sampler else None
        sweep, build = [], []
        for _ in range(5):                                   # kernel-level durations from a separate profiled loop
            step(profile=True)
            st = s.h.stats()
            sweep.append(st.sweep_ms)
            build.append(st.build_ms)
        if os.environ.get("CLM_BENCH_VERBOSE"):
            st = s.h.stats()
            print(f"[rank {rank}] step {sum(times) / args.steps:.3f} ms  sweep {statistics.mean(sweep):.3f}  build {statistics.mean(build):.3f}  "
                  f"owned {n_own} foreign {s.n_foreign} records {st.n_total[0]} tiles {st.n_tiles} cells {st.n_cells}", flush=True)
        t = torch.tensor([sum(times) / args.steps, statistics.mean(sweep), float(s.n_foreign), float(n_own), statistics.mean(build)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        # end-to-end arm: host buffers in (pinned), forces + energy back to the host, inside the timed region
        x_pin = torch.from_numpy(x_host).pin_memory()
        f_pin = torch.zeros((n_own, 3), dtype=torch.float32).pin_memory()
        # (a) synchronous: copy in, step, copy out, one frame at a time (the latency of a dependent step)
        e2e = []
        for it in range(2 + max(3, args.steps // 4)):
            flush_buf.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            x_dev.copy_(x_pin, non_blocking=True)
            e = step()
            f_pin.copy_(f_dev, non_blocking=True)
            e_host = e.cpu()
            b.record(stream)
            b.synchronize()
            if it >= 2:
                e2e.append(a.elapsed_time(b))
        te_sync = torch.tensor([sum(e2e) / len(e2e)], dtype=torch.float64, device=dev)
        dist.all_reduce(te_sync, op=dist.ReduceOp.MAX)
        # (b) pipelined frames (independent frames of a trajectory): copy-in of frame k+1, exchange + build + sweep of frame k
        # and copy-out of frame k-1 overlap on three streams, double-buffered by frame parity; EVERY frame's positions come
        # from pinned host memory and its forces + energy go back to it inside the timed region; the L2 flush runs on the
        # compute stream between frames (timed).  Wall clock over all frames between two barriers, max over ranks.
        cin, cout = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        xb = [x_dev, torch.empty_like(x_dev)]
        fb = [f_dev, torch.zeros_like(f_dev)]
        xp = [x_pin, torch.from_numpy(x_host).pin_memory()]
        fp = [f_pin, torch.zeros((n_own, 3), dtype=torch.float32).pin_memory()]
        ep = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        ev_h2d = [torch.cuda.Event() for _ in range(2)]
        ev_xfree = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        used = [False, False]
        keep = []
        trace = []      # CLM_BENCH_VERBOSE: (frame, label, timing event) of the last frames, printed as a timeline
        verbose = bool(os.environ.get("CLM_BENCH_VERBOSE"))

        def mark(k, label, st):
            if verbose and k >= args.steps + 1:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(st)
                trace.append((k, label, ev))

        issued = set()

        def h2d(k):
            p = k & 1
            with torch.cuda.stream(cin):
                if used[p]:
                    cin.wait_event(ev_xfree[p])              # the step that read this buffer two frames ago is done with it
                mark(k, "h2d start", cin)
                xb[p].copy_(xp[p], non_blocking=True)
                ev_h2d[p].record(cin)
                mark(k, "h2d end", cin)
            issued.add(k)

        def frame(k, prefetch_next=True):
            p = k & 1
            if k not in issued:
                h2d(k)
            if prefetch_next:
                h2d(k + 1)                                   # next frame's positions travel next to this frame's exchange + build + sweep
            stream.wait_event(ev_h2d[p])
            if used[p]:
                stream.wait_event(ev_out[p])                 # the copy-out of frame k-2 has drained this force buffer
            mark(k, "compute start", stream)
            flush_buf.zero_()
            s.update(xb[p])
            mark(k, "exchange end", stream)
            copy_out_pending()                               # frame k-1's outputs travel next to this frame's build + sweep
            e = s.map_lj(W.ARGON_C6, W.ARGON_C12, fb[p])
            ev_xfree[p].record(stream)
            ev_done[p].record(stream)
            mark(k, "compute end", stream)
            keep.append(e)
            pending.append((k, p, e))
            used[p] = True
            if len(keep) > 4:
                keep.pop(0)

        pending = []

        def copy_out_pending():
            while pending:
                k, p, e = pending.pop(0)
                with torch.cuda.stream(cout):
                    cout.wait_event(ev_done[p])
                    mark(k, "d2h start", cout)
                    fp[p].copy_(fb[p], non_blocking=True)
                    ep[p].copy_(e, non_blocking=True)
                    ev_out[p].record(cout)
                    mark(k, "d2h end", cout)

        for k in range(4):
            frame(k, prefetch_next=(k < 3))
        copy_out_pending()
        torch.cuda.synchronize()
        dist.barrier()
        import time as _time
        t0 = _time.perf_counter()
        for k in range(4, 4 + args.steps):                   # every timed frame's copy-in is issued inside the timed region
            frame(k, prefetch_next=(k < 3 + args.steps))
        copy_out_pending()
        torch.cuda.synchronize()
        te = torch.tensor([1e3 * (_time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_pipe = float(ep[(args.steps + 3) & 1][0])
        x_dev, f_dev = xb[0], fb[0]
        if trace:
            base = trace[0][2]
            print(f"[rank {rank}] pipelined e2e timeline (ms): " + "; ".join(f"f{k} {lab} {base.elapsed_time(ev):.2f}" for k, lab, ev in trace), flush=True)
        # BASELINE metric 2 at N GPUs: neighbour-list build (halo exchange + cell-list build + emission, per-rank lists left
        # on the device), max over ranks
        nl_ms, nl_pairs = [], 0
        # ~77.4 in-cutoff pairs per particle at this density and cutoff, 24-byte records: skip when the per-rank list would
        # not comfortably fit next to the particle arrays (the 64M-particle single-GPU run: 119 GB)
        big = torch.tensor([1.0 if n_own * 77.4 * 24 * 1.2 > 60e9 else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(big, op=dist.ReduceOp.MAX)           # one decision for every rank (the timing below has collectives)
        if float(big) > 0:
            args.no_nl = True
        if not args.no_nl:
            for it in range(4):
                flush_buf.zero_()
                torch.cuda.synchronize()
                dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                s.update(x_dev)
                nl_pairs = s.h.neighborlist_count()          # synchronous: the record count comes back to the host
                b.record(stream)
                b.synchronize()
                if it >= 1:
                    nl_ms.append(a.elapsed_time(b))
            tn = torch.tensor([statistics.mean(nl_ms), float(nl_pairs)], dtype=torch.float64, device=dev)
            tn_max, tn_sum = tn.clone(), tn.clone()
            dist.all_reduce(tn_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(tn_sum, op=dist.ReduceOp.SUM)
        # reference-stencil candidates of the GLOBAL system (roofline work model): cell histogram summed over ranks
        m = [int(np.floor(float(unitcell[k]) / cutoff)) for k in range(3)]
        cs = [float(unitcell[k]) / m[k] for k in range(3)]
        ci = [torch.clamp((torch.remainder(x_dev[:, k].double(), float(unitcell[k])) / cs[k]).floor().long(), 0, m[k] - 1) for k in range(3)]
        hist = torch.bincount((ci[0] * m[1] + ci[1]) * m[2] + ci[2], minlength=m[0] * m[1] * m[2]).double()
        dist.all_reduce(hist)
    if rank == 0:
        ms = float(tmax[0])
        n_total = int(tsum[3])
        line = {
            "metric": bench.METRIC, "value": P_in / (ms * 1e-3), "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: LJ energy+forces, argon-density particles, cubic-lattice PBC box slab-decomposed along x, cutoff 12 A "
                                   "(BASELINE.json configs[4]; 8 ranks = the 64M-particle case)",
                       "n_particles": n_total, "particles_per_gpu": n_total // world, "in_cutoff_pairs": P_in,
                       "halo_particles_per_gpu_max": int(tmax[2]),
                       "step": "halo exchange (NCCL send/recv) + UpdateCellList! + pairwise!(LJ energy+forces) + energy all_reduce",
                       "l2": "flushed between timed steps (256 MiB memset, untimed); per-step CUDA events, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": P_in / (float(te[0]) * 1e-3), "unit": bench.UNIT, "ms_per_step": float(te[0]),
                    "h2d_bytes_per_step": int(x_host.nbytes) * world, "d2h_bytes_per_step": (int(f_pin.numel()) * 4 + 4) * world,
                    "mode": "pipelined frames: every frame's positions are copied from pinned host memory and its forces + energy copied back inside the "
                            "timed region; copy-in of frame k+1, halo exchange + build + sweep of frame k and copy-out of frame k-1 overlap on three streams; "
                            "L2 flush between frames on the compute stream, inside the timed region; wall clock over all frames, max over ranks",
                    "energy": e_pipe},
            "e2e_sync": {"value": P_in / (float(te_sync[0]) * 1e-3), "unit": bench.UNIT, "ms_per_step": float(te_sync[0]),
                         "mode": "copy in, step, copy out, one frame at a time (barrier between frames): the latency of a dependent step"},
            "gpu_launches": launches * world,
            "breakdown_ms": {"step_max": ms, "sweep_kernel_max": float(tmax[1]), "build_max": float(tmax[4])},
            "energy": float(e_host),
        }
        # roofline of the dominant kernel (k_sweep), per rank: algorithmic flops of the rank's share of the pairs over the
        # slowest rank's kernel time, against the FP32 FMA peak measured live on this GPU
        h = hist.cpu().numpy().reshape(m)
        C_st = (h * (h - 1) / 2).sum()
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    if (dx, dy, dz) > (0, 0, 0):
                        C_st += (h * np.roll(h, (-dx, -dy, -dz), axis=(0, 1, 2))).sum()
        F_alg = (bench.FLOPS_PER_CANDIDATE * C_st + bench.FLOPS_PER_PAIR_LJ * P_in) / world
        peak = clm._capi.measure_fma_peak(np.float32, local)
        ach = F_alg / (float(tmax[1]) * 1e-3) / 1e12
        line["roofline"] = {"bound": "fp32", "kernel": "k_sweep<float, MODE_ALL, FLJ<float,true,true>> (per rank, slowest rank)", "achieved": ach,
                            "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                            "algorithmic_flops_per_launch": F_alg, "kernel_ms": float(tmax[1]),
                            "peak_source": "FP32 FMA peak measured live by clm_measure_fma_peak on rank 0's GPU; no tensor cores on this path"}
        if not args.no_nl:
            line["neighborlist_build"] = {"config": "the same slab-decomposed system, cutoff 12 A: halo exchange + cell-list build + emission, per-rank lists left on the device",
                                          "pairs": int(tn_sum[1]), "ms_max_over_ranks": float(tn_max[0])}
        # strong-scaling denominator: the SAME global system (all ranks' planes) on rank 0's GPU alone, through a plain handle
        # (the other ranks wait at the barrier below)
        if not args.no_64m:
            one = bench.c5_single_gpu(clm, torch, local, stream, flush_buf, nside, nx_per_rank * world, steps=3)
            key = "strong_scaling_64M" if one["n_particles"] == 64_000_000 else "strong_scaling"
            line[key] = {"n_particles": one["n_particles"], "ms_per_step_1gpu": one["ms_per_step"], "ms_per_step_ngpu": ms, "n_gpus": world,
                         "speedup": one["ms_per_step"] / ms, "sweep_kernel_ms_1gpu": one["sweep_kernel_ms"], "build_ms_1gpu": one["build_ms"],
                         "energy_1gpu": one["energy"], "energy_ngpu": float(e_host), "energy_rel_diff": abs(one["energy"] - float(e_host)) / abs(float(e_host)),
                         "note": "same particle system, device-resident steps; the 1-GPU step has no halo exchange and no all_reduce"}
        if not args.no_cpu_baseline:
            # the CPU restatement on a bounded sample of the same system (2M particles of the same generator), all host threads
            from oracle import oracle as om
            import time
            nt = om.lib().ora_num_threads()
            xs, ucs = slab_lattice(0, 1, 126, 126, dtype)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                o = om.Oracle(xs, cutoff, unitcell=ucs, dtype=dtype)
                o.lj(W.ARGON_C6, W.ARGON_C12, forces=True, nbatches=nt)
                ts.append(time.perf_counter() - t0)
            npc = o.sum_d_d2(nbatches=nt)[2]
            line["cpu_baseline"] = {"value": npc / min(ts), "unit": bench.UNIT, "cores": nt, "kind": "port",
                                    "sample": "126^3 = 2 000 376 particles of the same generator / density / cutoff (bounded sample of the sharded system), "
                                              "best of 2 x (build + LJ energy+forces), C++/OpenMP restatement of the reference", "seconds_per_step": min(ts)}
        print(json.dumps(line, default=lambda o: o.item() if hasattr(o, "item") else str(o)))
    s.close()
    dist.barrier()
    dist.destroy_process_group()

"""Multi-GPU slab decomposition of the cutoff-pair path (the reference is single-process).

One process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink on the B200 box, gloo in the CPU tests).
The box is cut into slabs of whole REFERENCE cells along reference dimension 1.  A rank owns the particles whose
(wrapped, nudged) cell index along that dimension falls in its slab; before every cell-list build it receives from
its two neighbours (periodically) the particles of their `lcell` outermost cell layers -- the halo -- and hands them
to the engine as FOREIGN particles (clm_set_foreign): partners j like any other record, never particle i.  Every pair
evaluation of the single-GPU sweep therefore happens on exactly one rank, on bit-identical coordinates:

  * per-particle outputs (forces) stay sharded with their owners -- the full-shell sweep needs no reverse exchange;
  * per-particle INPUTS of a pair function (weights, velocities, user side arrays) travel with the halo (`aux`);
  * scalars and histograms are summed with all_reduce; minimum distances with an all_gather + min;
  * neighbour lists stay per rank (indices mapped to the caller's global ids), concatenation is the caller's choice.

The path has ONE exchange step (the halo); it is a neighbour send/recv, not a collective reduction, so it is issued
as batched P2P ops.

Triclinic cells (unitcell = matrix): lattice shifts move images across slabs, so the halo of a rank is every particle
with ANY periodic image inside the rank's slab +- lcell layers (computed here from the engine's own box record, with one
layer of slack; extra halo particles are harmless -- they only ever act as partners), sent point to point to whichever
ranks need it.  The reference's exactly-once rule for triclinic cells compares particle indices (index_i < index_j,
src/internals/self.jl:164-184), so a rank hands owned + halo particles to the engine as ONE array sorted by GLOBAL id with
the halo rows flagged (clm_set_foreign_mask): local order = global order, every pair is evaluated once, by the owner of
its smaller-index particle, from the same periodic image as on one GPU.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import Handle


class SlabPlan:
    """Pure host logic: which cell layers each rank owns, who needs what.  `n_inner` = number of reference cells along
    dimension 1 that hold real particles (Box.nc[0] - 2*lcell - 1); real cells are lcell .. lcell + n_inner - 1."""

    def __init__(self, n_inner, lcell, world):
        if world > 1 and n_inner // world < lcell:
            raise ValueError(f"slabs would be thinner than the stencil reach: {n_inner} cell layers over {world} ranks, lcell = {lcell}")
        self.n_inner, self.lcell, self.world = int(n_inner), int(lcell), int(world)
        self.bounds = [lcell + (n_inner * r) // world for r in range(world + 1)]   # rank r owns [bounds[r], bounds[r+1])

    def owner_of(self, c):
        """rank owning cell layer(s) c (numpy or torch integer array)."""
        b = self.bounds
        if isinstance(c, torch.Tensor):
            edges = torch.as_tensor(b[1:-1], device=c.device, dtype=c.dtype)
            return torch.bucketize(c, edges, right=True)
        return np.searchsorted(np.asarray(b[1:-1]), c, side="right")

    def face_masks(self, c, rank):
        """(to_lower, to_upper): which owned particles (cell layers c) the lower / upper neighbour needs."""
        lo, hi = self.face_ranges(rank)[0], self.face_ranges(rank)[3]
        return (c >= lo) & (c < lo + self.lcell), (c >= self.face_ranges(rank)[2]) & (c < hi)

    def face_ranges(self, rank):
        """(lo, lo + lcell, hi - lcell, hi'): the cell layers of the lower and the upper face.  The last rank's upper face
        is closed at bounds[world]: a real particle can round into layer lcell + n_inner (the build's border nudge keeps it
        there), it belongs to the last rank and rank 0 needs it as a periodic partner."""
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        return lo, lo + self.lcell, hi - self.lcell, hi + (1 if rank == self.world - 1 else 0)

    def neighbours(self, rank):
        return (rank - 1) % self.world, (rank + 1) % self.world


def exchange_halo(payloads, to_lower, to_upper, plan, rank, group=None):
    """Send the rows selected by the face masks to the two neighbours and return what they sent us.

    payloads: list of tensors with the same first extent (positions, ids, weights, ...), on any device the process
    group can move (CUDA for nccl, CPU for gloo; CUDA tensors are staged through the host when the backend is gloo).
    Returns a list of tensors (rows received from the upper neighbour first, then from the lower one)."""
    world = plan.world
    if world == 1:
        return [p[:0] for p in payloads]
    lower, upper = plan.neighbours(rank)
    backend = dist.get_backend(group)
    stage = backend == "gloo"
    if world == 2:                      # both faces go to the same peer: one message, no duplicates
        sends = {lower: to_lower | to_upper}
    else:
        sends = {lower: to_lower, upper: to_upper}
    counts = torch.zeros(world, dtype=torch.int64)
    for peer, m in sends.items():
        counts[peer] = int(m.sum())
    dev = payloads[0].device
    comm_dev = torch.device("cpu") if stage else dev
    all_counts = [torch.zeros(world, dtype=torch.int64, device=comm_dev) for _ in range(world)]
    dist.all_gather(all_counts, counts.to(comm_dev), group=group)
    out = []
    for p in payloads:
        ops, recv = [], {}
        keep = []
        for peer, m in sends.items():
            buf = p[m].contiguous()
            buf = buf.cpu() if stage else buf
            keep.append(buf)
            if buf.shape[0]:
                ops.append(dist.P2POp(dist.isend, buf, peer, group))
        for peer in sends:
            n_in = int(all_counts[peer][rank])
            r = torch.empty((n_in,) + tuple(p.shape[1:]), dtype=p.dtype, device=comm_dev)
            recv[peer] = r
            if n_in:
                ops.append(dist.P2POp(dist.irecv, r, peer, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        order = [upper, lower] if world > 2 else [upper]
        got = torch.cat([recv[q] for q in order], dim=0)
        out.append(got.to(dev) if stage else got)
    return out


def exchange_rows(payloads, sends, world, rank, group=None):
    """General point-to-point row exchange: `sends` maps peer rank -> boolean mask over the rows of every payload.
    Returns the received rows of every payload, concatenated in ascending peer order."""
    if world == 1:
        return [p[:0] for p in payloads]
    stage = dist.get_backend(group) == "gloo"
    dev = payloads[0].device
    comm_dev = torch.device("cpu") if stage else dev
    counts = torch.zeros(world, dtype=torch.int64)
    for peer, m in sends.items():
        counts[peer] = int(m.sum())
    all_counts = [torch.zeros(world, dtype=torch.int64, device=comm_dev) for _ in range(world)]
    dist.all_gather(all_counts, counts.to(comm_dev), group=group)
    n_in = [int(all_counts[q][rank]) if q != rank else 0 for q in range(world)]
    out = []
    for p in payloads:
        ops, keep, recv = [], [], {}
        for peer, m in sends.items():
            if int(counts[peer]) == 0:
                continue
            buf = p[m].contiguous()
            buf = buf.cpu() if stage else buf
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, peer, group))
        for q in range(world):
            if n_in[q]:
                recv[q] = torch.empty((n_in[q],) + tuple(p.shape[1:]), dtype=p.dtype, device=comm_dev)
                ops.append(dist.P2POp(dist.irecv, recv[q], q, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        got = torch.cat([recv[q] for q in sorted(recv)], dim=0) if recv else torch.empty((0,) + tuple(p.shape[1:]), dtype=p.dtype, device=comm_dev)
        out.append(got.to(dev) if stage else got)
    return out


class SlabSystem:
    """One rank's share of a slab-decomposed self-set particle system on its own B200.

    x_owned: (n, N) torch tensor (CUDA) or numpy array with the particles this rank owns -- use `SlabSystem.partition`
    to split a global array.  `ids` are the caller's global 1-based particle ids of the owned particles (for neighbour
    lists)."""

    def __init__(self, unitcell, cutoff, dtype=np.float32, lcell=1, dim=3, device=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = np.dtype(dtype)
        self.tdtype = torch.float32 if self.dtype == np.float32 else torch.float64
        self.dim = dim
        uc = np.asarray(unitcell, dtype=self.dtype)
        self.triclinic = uc.ndim == 2
        self.h = Handle(dim, self.dtype, self.device.index or 0)
        # the engine enqueues on torch's current stream, so torch ops and collectives are ordered with its kernels
        # (0 is the legacy default stream: CUDA's cudaStreamLegacy handle is 1)
        self.h.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)
        self.h.set_box(_capi.TRICLINIC if self.triclinic else _capi.ORTHORHOMBIC, uc, cutoff, lcell)
        b = self.h.get_box()
        if self.triclinic:
            # what the halo selection needs from the engine's box record (float64 copies; column-major, stride = dim)
            g = lambda a: torch.tensor(np.array(a[:dim * dim], dtype=np.float64).reshape(dim, dim).T.copy(), dtype=torch.float64, device=self.device)
            self._M, self._A = g(b.input_unit_cell), g(b.aligned_unit_cell)
            self._cb0, self._cs0 = float(b.computing_box_min[0]), float(b.cell_size[0])
            idx = torch.cartesian_prod(*([torch.tensor([-1.0, 0.0, 1.0], dtype=torch.float64, device=self.device)] * dim)).reshape(-1, dim)
            self._shift0 = torch.unique(idx @ self._A[0])        # distinct displacements of the periodic images along dimension 1
        self._own_rows = None         # triclinic: rows of the owned particles in the engine's (global-id ordered) array
        self.plan = SlabPlan(int(b.nc[0]) - 2 * lcell - 1, lcell, self.world)
        self._inner_cells = float(np.prod([max(1, int(b.nc[k]) - 2 * lcell - 1) for k in range(dim)]))
        self._lcell = lcell
        self._n_global = None
        self._cap = None              # capacity of the fixed-size halo messages (set by the first, exact exchange)
        self.force_fast = False       # tests: take the single-sync exchange over gloo too (messages staged through the host)
        self.fast_exchanges = 0       # how many updates went through the single-sync exchange
        self.n_owned = 0
        self.n_foreign = 0
        self.ids = None
        self.foreign_ids = None

    def set_stream(self, stream):
        self.h.set_stream(stream.cuda_stream or 1)

    def cell_layers(self, x):
        """reference-cell index along dimension 1, computed by the engine (same arithmetic as the build)."""
        return self.h.cell_coords(x, 0)

    def partition(self, x_global, ids=None):
        """rows of a global coordinate array (every rank passes the same array) this rank owns."""
        x = torch.as_tensor(x_global).to(self.device, self.tdtype).contiguous()
        c = self.cell_layers(x)
        mine = self.plan.owner_of(c.to(torch.int64)) == self.rank
        gid = torch.arange(1, x.shape[0] + 1, device=self.device, dtype=torch.int64) if ids is None else torch.as_tensor(ids).to(self.device)
        return x[mine].contiguous(), gid[mine].contiguous()

    def update(self, x_owned, ids=None, aux=None):
        """halo exchange + hand owned / foreign particles to the engine (nothing is built until the next map).
        `aux`: (n_owned, k) per-particle side data (weights, velocities, ...) that the halo particles carry along."""
        x = torch.as_tensor(x_owned).to(self.device, self.tdtype).contiguous()
        self.n_owned = int(x.shape[0])
        if aux is not None:
            aux = torch.as_tensor(aux).to(self.device, self.tdtype).reshape(self.n_owned, -1).contiguous()
        if self.triclinic:
            return self._update_triclinic(x, ids, aux)
        if self._n_global is None:
            # the engine sizes its device grid from the particle density; a rank only sees its slab, so tell it the
            # global density (same rule as Engine::build: ~4 particles per device cell)
            n = torch.tensor([self.n_owned], dtype=torch.int64, device=self.device)
            if self.world > 1:
                dist.all_reduce(n, group=self.group)
            self._n_global = int(n)
            per_cell = self._n_global / self._inner_cells
            sub = int(np.floor(max(per_cell / 4.0, 1.0) ** (1.0 / self.dim) + 0.35))
            self.h.set_option("sub", max(1, min(sub, 7 // self._lcell)))
        got = None
        ids_t = None if ids is None else torch.as_tensor(ids).to(self.device)
        if self.world > 1 and self._cap is not None and (self.force_fast or dist.get_backend(self.group) != "gloo"):
            got = self._exchange_fast(x, ids_t, aux)
        if got is not None:
            self.fast_exchanges += 1
            got, self.foreign_ids, aux_f = got
            self.ids = ids_t
            self.aux = None if aux is None else torch.cat([aux, aux_f], dim=0).contiguous()
        else:
            c = self.cell_layers(x).to(torch.int64)
            to_lower, to_upper = self.plan.face_masks(c, self.rank)
            payloads = [x] + ([] if ids is None else [ids_t]) + ([] if aux is None else [aux])
            res = exchange_halo(payloads, to_lower, to_upper, self.plan, self.rank, self.group)
            got = res[0].contiguous()
            self.ids = None if ids is None else payloads[1]
            self.foreign_ids = None if ids is None else res[1]
            # side data of owned + foreign particles, in the engine's index order (owned first)
            self.aux = None if aux is None else torch.cat([aux, res[-1]], dim=0).contiguous()
            if self.world > 1:
                # size the fixed-capacity messages of the fast path from what this exchange moved (50 % slack)
                m = torch.tensor([max(int(to_lower.sum()), int(to_upper.sum()), int((to_lower | to_upper).sum()) if self.world == 2 else 0)],
                                 dtype=torch.int64, device=self.device if dist.get_backend(self.group) != "gloo" else "cpu")
                dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
                self._alloc_fast(int(int(m) * 1.5) + 1024)
        self.x_owned, self.x_foreign = x, got
        self.n_foreign = int(got.shape[0])
        self.h.set_positions(0, x)
        self.h.set_foreign(0, self.x_foreign)
        return self

    def _update_triclinic(self, x, ids, aux):
        if ids is None:
            raise ValueError("triclinic slabs need the global particle ids (partition() returns them): the index_i < index_j rule follows the global numbering")
        ids_t = torch.as_tensor(ids).to(self.device, torch.int64)
        if self._n_global is None:
            n = torch.tensor([self.n_owned], dtype=torch.int64, device=self.device)
            if self.world > 1:
                n = n.cpu() if dist.get_backend(self.group) == "gloo" else n
                dist.all_reduce(n, group=self.group)
            self._n_global = int(n)
            sub = int(np.floor(max(self._n_global / self._inner_cells / 4.0, 1.0) ** (1.0 / self.dim) + 0.35))
            self.h.set_option("sub", max(1, min(sub, 7 // self._lcell)))
        sends = {}
        if self.world > 1 and self.n_owned:
            # cell layer along dimension 1 of every periodic image of every owned particle (float64 restatement of the build's
            # wrap; one layer of slack on either side absorbs the rounding differences to the engine's arithmetic)
            xd = x.to(torch.float64)
            frac = torch.linalg.solve(self._M, xd.T).T
            frac = frac - torch.floor(frac)
            p0 = frac @ self._A[0]
            layer = torch.floor((p0[:, None] + self._shift0[None, :] - self._cb0) / self._cs0)
            b, lc = self.plan.bounds, self._lcell
            for r in range(self.world):
                if r == self.rank:
                    continue
                lo, hi = b[r] - lc - 1, b[r + 1] + lc + 1 + (1 if r == self.world - 1 else 0)
                m = ((layer >= lo) & (layer < hi)).any(dim=1)
                if bool(m.any()):
                    sends[r] = m
        payloads = [x, ids_t] + ([] if aux is None else [aux])
        res = exchange_rows(payloads, sends, self.world, self.rank, self.group)
        xs, gs = torch.cat([x, res[0]], dim=0), torch.cat([ids_t, res[1]])
        foreign = torch.cat([torch.zeros(self.n_owned, dtype=torch.uint8, device=self.device), torch.ones(res[0].shape[0], dtype=torch.uint8, device=self.device)])
        order = torch.argsort(gs)
        self.x_local = xs[order].contiguous()
        self.ids_local = gs[order].contiguous()
        self._mask = foreign[order].contiguous()
        self._own_rows = torch.nonzero(self._mask == 0).flatten()        # owned particles, by ascending global id
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.shape[0], device=self.device)
        self._own_pos = inv[:self.n_owned]                              # row of the k-th owned particle (caller's order)
        self.aux = None if aux is None else torch.cat([aux, res[2]], dim=0)[order].contiguous()
        self.ids, self.foreign_ids = ids_t, res[1]
        self.n_foreign = int(res[0].shape[0])
        self.h.set_positions(0, self.x_local)
        self.h.set_foreign(0, None)
        self.h.set_foreign_mask(0, self._mask)
        return self

    def _id_table(self):
        """global id of every engine-side particle index (0-based rows)."""
        return self.ids_local if self.triclinic else torch.cat([self.ids, self.foreign_ids])

    def _forces_arg(self, forces):
        """the per-particle output buffer the engine writes: the caller's (owned rows) or, for triclinic slabs, a local one
        covering owned + halo rows in global-id order."""
        if forces is None or not self.triclinic:
            return forces
        return torch.zeros((self.x_local.shape[0], forces.shape[1]), dtype=forces.dtype, device=self.device)

    def _forces_back(self, forces, buf):
        if forces is not None and self.triclinic:
            forces.copy_(buf[self._own_pos])

    def _alloc_fast(self, cap):
        if self._cap is not None and cap <= self._cap:
            return
        d, t = self.device, self.tdtype
        self._cap = cap
        self._send = [torch.empty((cap, self.dim), dtype=t, device=d) for _ in range(2)]
        self._recv = [torch.empty((cap, self.dim), dtype=t, device=d) for _ in range(2)]
        self._idx = [torch.zeros(cap, dtype=torch.int32, device=d) for _ in range(2)]
        self._cnt = torch.zeros(2, dtype=torch.int32, device=d)
        self._rcnt = torch.zeros(2, dtype=torch.int32, device=d)
        self._side = {}               # receive buffers of the side payloads (ids, aux), keyed by (dtype, columns)

    def _side_recv(self, p, k):
        key = (k, p.dtype, tuple(p.shape[1:]))
        if key not in self._side:
            self._side[key] = [torch.empty((self._cap,) + tuple(p.shape[1:]), dtype=p.dtype, device=self.device) for _ in range(2)]
        return self._side[key]

    def _exchange_fast(self, x, ids=None, aux=None):
        """halo exchange with ONE host synchronisation: the engine selects both faces in one pass into fixed-capacity
        buffers (and returns the source rows, with which the side payloads -- global ids, weights, velocities -- are
        gathered into messages of the same order), the buffers and their fill counts travel in one batch of send/recv, an
        all-reduced overflow flag makes the fallback decision COLLECTIVE, and only then the counts are read back.  Returns
        (foreign positions, foreign ids, foreign aux), or None on EVERY rank when a face of ANY rank outgrew the capacity
        (the caller then takes the exact path, which re-sizes the messages)."""
        lower, upper = self.plan.neighbours(self.rank)
        merge = self.world == 2
        cap = self._cap
        self._cnt.zero_()
        self.h.select_layers(x, 0, self.plan.face_ranges(self.rank), merge, self._send[0], self._send[1], self._cnt, self._idx[0], self._idx[1])
        g = self.group
        stage = dist.get_backend(g) == "gloo"        # CPU tests force this path over gloo: messages staged through the host
        side = [p for p in (ids, aux) if p is not None]
        # rows beyond the fill count gather row 0 (the index buffers start zeroed and only ever hold valid rows): harmless
        side_send = [[p.index_select(0, self._idx[f].long().clamp_(0, max(p.shape[0] - 1, 0))) if p.shape[0] else p.new_zeros((cap,) + tuple(p.shape[1:])) for f in range(2)] for p in side]
        side_recv = [self._side_recv(p, k) for k, p in enumerate(side)]
        faces = [(0, upper, 0)] if merge else [(0, lower, 1), (1, upper, 0)]   # (my face f, peer, the peer's slot in my receive buffers)
        msgs = []      # (send tensor, recv tensor, peer)
        for f, peer, rslot in faces:
            msgs.append((self._send[f], self._recv[rslot], peer))
            msgs.append((self._cnt[f:f + 1], self._rcnt[rslot:rslot + 1], peer))
            for ss, rr in zip(side_send, side_recv):
                msgs.append((ss[f], rr[rslot], peer))
        if stage:
            host = [(a.cpu(), torch.empty(b.shape, dtype=b.dtype), b, peer) for a, b, peer in msgs]
            ops = [dist.P2POp(dist.isend, a, peer, g) for a, _, _, peer in host] + [dist.P2POp(dist.irecv, r, peer, g) for _, r, _, peer in host]
        else:
            ops = [dist.P2POp(dist.isend, a, peer, g) for a, _, peer in msgs] + [dist.P2POp(dist.irecv, b, peer, g) for _, b, peer in msgs]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if stage:
            for _, r, b, _ in host:
                b.copy_(r)
        # the overflow decision is collective: every rank learns whether ANY face of ANY rank outgrew its message
        allc = torch.cat([self._cnt, self._rcnt])
        flag = (allc.max() > cap).to(torch.int32).reshape(1)
        flag = flag.cpu() if stage else flag
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=g)
        if stage:
            counts = torch.cat([allc, flag.to(allc.device)]).cpu().tolist()  # the one synchronisation
        else:
            # the one synchronisation -- through mapped pinned memory, not a DMA copy: a 20-byte cudaMemcpy would queue behind
            # the force copy-out of the previous pipelined frame on the device->host engine (clm_read_ints)
            counts = self.h.read_ints(torch.cat([allc, flag.to(allc.device)]).to(torch.int32).contiguous())
        if counts[4]:
            return None
        n_up, n_lo = counts[2], (0 if merge else counts[3])
        cat = (lambda r: torch.cat([r[0][:n_up], r[1][:n_lo]], dim=0) if n_lo else r[0][:n_up])
        out = [cat(self._recv)] + [cat(r) for r in side_recv]
        k = 1
        f_ids = f_aux = None
        if ids is not None:
            f_ids = out[k]; k += 1
        if aux is not None:
            f_aux = out[k]
        return out[0], f_ids, f_aux

    def reset_halo_capacity(self):
        """the next update() takes the exact (collective) exchange again and re-sizes the fixed-capacity messages."""
        self._cap = None

    # ---- catalogue entry points: local map + the reduction the output type needs ----
    def map_lj(self, c6, c12, forces=None, profile=False):
        """returns the GLOBAL energy (all_reduce) and fills `forces` (n_owned x N, device) for the owned particles."""
        e = torch.zeros(1, dtype=self.tdtype, device=self.device)
        buf = self._forces_arg(forces)
        self.h.map_lj(c6, c12, e, buf, reset=True, profile=profile)
        self._forces_back(forces, buf)
        if self.world > 1:
            dist.all_reduce(e, group=self.group)
        return e

    def _need_aux(self, k):
        if getattr(self, "aux", None) is None or self.aux.shape[1] != k:
            raise ValueError(f"this map reads {k} side value(s) per particle: pass them as update(x, aux=...) so that the halo carries them")
        return self.aux

    def map_coulomb(self, k, forces=None, profile=False):
        """k w_i w_j / d with the weights given as update(..., aux=w): GLOBAL energy, forces of the owned particles."""
        w = self._need_aux(1)
        e = torch.zeros(1, dtype=self.tdtype, device=self.device)
        buf = self._forces_arg(forces)
        self.h.map_coulomb(k, w, None, e, buf, reset=True, profile=profile)
        self._forces_back(forces, buf)
        if self.world > 1:
            dist.all_reduce(e, group=self.group)
        return e

    def pairvel(self, rbins):
        """mean pairwise velocity histogram with the velocities given as update(..., aux=v): GLOBAL (counts, sums)."""
        v = self._need_aux(self.dim)
        nb = len(rbins) - 1
        counts = torch.zeros(nb, dtype=torch.int64, device=self.device)
        sums = torch.zeros(nb, dtype=self.tdtype, device=self.device)
        self.h.map_pairvel(v, None, np.asarray(rbins, dtype=self.dtype), counts, sums, reset=True)
        if self.world > 1:
            dist.all_reduce(counts, group=self.group)
            dist.all_reduce(sums, group=self.group)
        return counts, sums

    def mindist(self):
        """(i, j, d) of the closest pair of the GLOBAL system (global ids when update() was given `ids`, else rank-local
        indices of the winning rank); ties are broken by the smaller (i, j) as on one GPU."""
        i, j = np.zeros(1, np.int64), np.zeros(1, np.int64)
        d = np.full(1, np.inf, self.dtype)
        self.h.map_mindist(i, j, d, reset=True)
        if self.ids is not None and i[0] > 0:
            table = self._id_table().cpu().numpy()
            i[0], j[0] = table[i[0] - 1], table[j[0] - 1]
        cand = [(float(d[0]), int(i[0]), int(j[0]))]
        if self.world > 1:
            allc = [None] * self.world
            dist.all_gather_object(allc, cand[0], group=self.group)
            cand = [c for c in allc if c[1] > 0] or [cand[0]]
        best = min(cand, key=lambda c: (c[0], min(c[1], c[2]), max(c[1], c[2])))
        return best[1], best[2], best[0]

    def map_custom(self, source, name, params=(), scalars=0, per_particle=0, nbins=0):
        """a run-time compiled user pair function (clm_map_custom); side arrays come from update(..., aux=...).
        Returns (scalars, per_particle, hist_counts, hist_sums): scalars and histograms are GLOBAL (all_reduce), the
        per-particle output covers the owned particles."""
        key = (source, name)
        if not hasattr(self, "_custom"):
            self._custom = {}
        if key not in self._custom:
            self._custom[key] = self.h.custom_compile(source, name)
        fid, info = self._custom[key]
        aux = self._need_aux(info.naux) if info.naux else None
        sc = torch.zeros(info.nscalar, dtype=self.tdtype, device=self.device) if info.nscalar else None
        pp = torch.zeros((self.x_local.shape[0] if self.triclinic else self.n_owned, info.npart), dtype=self.tdtype, device=self.device) if info.npart else None
        hc = torch.zeros(nbins, dtype=torch.int64, device=self.device) if info.hist else None
        hs = torch.zeros(nbins, dtype=self.tdtype, device=self.device) if info.hist else None
        self.h.map_custom(fid, params, aux, None, sc, pp, hc, hs, reset=True)
        if self.world > 1:
            for t in (sc, hc, hs):
                if t is not None:
                    dist.all_reduce(t, group=self.group)
        if pp is not None and self.triclinic:
            pp = pp[self._own_pos].contiguous()
        return sc, pp, hc, hs

    def sum_d_d2(self):
        sd, sd2, n = (torch.zeros(1, dtype=self.tdtype, device=self.device), torch.zeros(1, dtype=self.tdtype, device=self.device),
                      torch.zeros(1, dtype=torch.int64, device=self.device))
        self.h.map_sum_d_d2(sd, sd2, n, reset=True)
        if self.world > 1:
            for t in (sd, sd2, n):
                dist.all_reduce(t, group=self.group)
        return float(sd), float(sd2), int(n)

    def dist_hist(self, width, nbins):
        counts = torch.zeros(nbins, dtype=torch.int64, device=self.device)
        self.h.map_dist_hist(width, counts, reset=True)
        if self.world > 1:
            dist.all_reduce(counts, group=self.group)
        return counts

    def neighborlist(self):
        """this rank's part of the neighbour list as a structured numpy array with GLOBAL ids (needs `ids` in update)."""
        n = self.h.neighborlist_count()
        rec = np.zeros(n, dtype=_capi.nl_dtype(self.dtype))
        if n:
            self.h.neighborlist_copy(rec)
        if self.ids is not None:
            table = self._id_table().cpu().numpy()
            rec["i"] = table[rec["i"] - 1]
            rec["j"] = table[rec["j"] - 1]
        return rec

    def close(self):
        self.h.close()


class SlabSystem2:
    """One rank's share of a slab-decomposed TWO-set system (cross.jl:8-25: every pair (x_i, y_j) once, i over the first set).

    A rank owns the x particles of its slab -- they are the particles i of its sweep -- and holds the y particles of its slab
    plus the `lcell` outermost layers of both neighbours (the halo of the y set).  The y set is the PARTNER set: none of its
    particles ever acts as particle i, so the halo needs no foreign flag and every pair is evaluated exactly once, by the
    owner of x_i, on bit-identical coordinates.  Per-x outputs stay with their owners; scalars, histograms and minima are
    reduced over the ranks.  Orthorhombic cells."""

    def __init__(self, unitcell, cutoff, dtype=np.float64, lcell=1, dim=3, device=None, group=None):
        self.s = SlabSystem(unitcell, cutoff, dtype=dtype, lcell=lcell, dim=dim, device=device, group=group)
        self.h, self.plan, self.rank, self.world, self.group = self.s.h, self.s.plan, self.s.rank, self.s.world, self.s.group
        self.device, self.dtype, self.tdtype, self.dim = self.s.device, self.s.dtype, self.s.tdtype, dim
        self._inner_cells, self._lcell, self._n_global = self.s._inner_cells, lcell, None

    def partition(self, x_global, ids=None):
        return self.s.partition(x_global, ids)

    def update(self, x_owned, y_owned, x_ids=None, y_ids=None, y_aux=None):
        """x_owned / y_owned: the particles of either set whose cell layer is in this rank's slab (partition()).  The y halo is
        exchanged here; `y_ids` / `y_aux` (global ids, per-particle side data of the y set) travel with it."""
        x = torch.as_tensor(x_owned).to(self.device, self.tdtype).contiguous()
        y = torch.as_tensor(y_owned).to(self.device, self.tdtype).contiguous()
        if self._n_global is None:
            n = torch.tensor([max(x.shape[0], y.shape[0])], dtype=torch.int64, device=self.device)
            if self.world > 1:
                n = n.cpu() if dist.get_backend(self.group) == "gloo" else n
                dist.all_reduce(n, group=self.group)
            self._n_global = int(n)
            sub = int(np.floor(max(self._n_global / self._inner_cells / 4.0, 1.0) ** (1.0 / self.dim) + 0.35))
            self.h.set_option("sub", max(1, min(sub, 7 // self._lcell)))
        c = self.s.cell_layers(y).to(torch.int64)
        to_lower, to_upper = self.plan.face_masks(c, self.rank)
        payloads = [y] + ([] if y_ids is None else [torch.as_tensor(y_ids).to(self.device)]) + \
                   ([] if y_aux is None else [torch.as_tensor(y_aux).to(self.device, self.tdtype).reshape(y.shape[0], -1).contiguous()])
        res = exchange_halo(payloads, to_lower, to_upper, self.plan, self.rank, self.group)
        self.n_x, self.n_y_owned, self.n_y_halo = int(x.shape[0]), int(y.shape[0]), int(res[0].shape[0])
        self.x_ids = None if x_ids is None else torch.as_tensor(x_ids).to(self.device)
        self.y_ids = None if y_ids is None else torch.cat([payloads[1], res[1]])
        self.y_aux = None if y_aux is None else torch.cat([payloads[-1], res[-1]], dim=0).contiguous()
        self.h.set_positions(0, x)
        self.h.set_positions(1, torch.cat([y, res[0]], dim=0).contiguous())
        return self

    def mindist(self):
        """(i, j, d) of the closest cross pair of the GLOBAL system (global ids when update() was given ids)."""
        i, j = np.zeros(1, np.int64), np.zeros(1, np.int64)
        d = np.full(1, np.inf, self.dtype)
        self.h.map_mindist(i, j, d, reset=True)
        if i[0] > 0 and self.x_ids is not None and self.y_ids is not None:
            i[0], j[0] = int(self.x_ids[i[0] - 1]), int(self.y_ids[j[0] - 1])
        cand = [(float(d[0]), int(i[0]), int(j[0]))]
        if self.world > 1:
            allc = [None] * self.world
            dist.all_gather_object(allc, cand[0], group=self.group)
            cand = [c for c in allc if c[1] > 0] or [cand[0]]
        best = min(cand, key=lambda c: (c[0], c[1], c[2]))
        return best[1], best[2], best[0]

    def neighborlist(self):
        """this rank's part of the cross neighbour list (i: x ids, j: y ids; global when update() was given ids)."""
        n = self.h.neighborlist_count()
        rec = np.zeros(n, dtype=_capi.nl_dtype(self.dtype))
        if n:
            self.h.neighborlist_copy(rec)
        if self.x_ids is not None and self.y_ids is not None:
            xi, yi = self.x_ids.cpu().numpy(), self.y_ids.cpu().numpy()
            rec["i"] = xi[rec["i"] - 1]
            rec["j"] = yi[rec["j"] - 1]
        return rec

    def sum_d_d2(self):
        return self.s.sum_d_d2()

    def close(self):
        self.h.close()

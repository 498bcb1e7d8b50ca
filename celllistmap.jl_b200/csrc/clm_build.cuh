// Cell-list construction kernels (the B200 replacement of UpdateCellList!,
// src/internals/CellLists.jl:727-927 / add_particles! :940-950 / add_particle_to_celllist! :983-1057,
// non-periodic src/internals/NonPeriodicCells.jl:63-70, :93-230).
//
// The reference builds an array of heap-allocated per-cell vectors; here the list is ONE
// counting sort into a cell-sorted packed record array:
//   k_bin<count>  : wrap + rotate each particle, enumerate its 3^N-1 lattice images, keep those inside
//                   the computing box, histogram real+image particles per cell (global atomics)
//   k_rows        : one warp per ROW of device cells: prefix of the row's counts; the row's base comes from one
//                   atomicAdd on the record counter, so rows are contiguous but placed in arbitrary order -- the
//                   sweep only ever reads record ranges inside one row.  Replaces a 3-kernel global scan.  The same
//                   warp cuts the row's active record range into tiles, appended to the tile array with one atomicAdd
//                   per row (tile order is irrelevant: tiles are dealt out by a work counter)
//   k_bin<scatter>: same traversal, records scattered to the per-cell atomic cursor; slot_of[particle] = its slot
//   k_twin        : (on request) the slot-tagged twin of the records for the Newton's-third-law force sweep, one
//                   coalesced pass in record order
// Round 2 also built the alternative the round-1 review asked for -- ONE wrapping pass that caches (position, cell) per
// particle, then a placement without divisions (first scattered like the records, then as a 4-byte order[] scatter + a
// coalesced gather) -- and measured it slower at every size: the wrap's instructions are not what
// bounds the build, its random memory streams are, and the cached variant adds streams.
// Every per-cell array is laid out with a row pitch of nx + 1 entries: cell_start[row * (nx + 1) + x] is the first
// record of cell x of the row and entry nx is the end of the row.
// All of it is HBM/latency bound integer + a few dozen flops per particle: no tensor cores.
// Algorithmic bytes per particle: read N*sizeof(T) twice, write (1+g)*sizeof(Rec),
// plus 8*n_cells for the histogram/prefix arrays, g = image fraction.
#pragma once
#include "clm_common.cuh"

namespace clm {

constexpr int IDX_NONE = 0x7fffffff;
// device scalar block (ints)
enum { DS_NAN = 0, DS_OOB = 1, DS_NTOT = 2, DS_NTILES = 3, DS_WORK = 4, DS_NCELLS_REAL = 5, DS_NGHOST = 12 /* + set */, DS_NOFIT = 14, DS_COUNT = 16 };
constexpr int DS_SET_STRIDE_DEV = 6;   // the scalar block of the second set starts at dscal + 6

// The device scalars are initialised and published by two one-warp kernels, not by cudaMemcpyAsync: a small copy on the
// compute stream queues on the same DMA engines as the megabyte copies of pipelined frames (positions in, forces out)
// and stalled the stream behind them (0.16 ms per step on the 1M-particle system).  h_pub is MAPPED pinned host memory.
static __global__ void k_dscal_init(int* __restrict__ dscal) {
    const int k = threadIdx.x;
    if (k < DS_COUNT) dscal[k] = (k == DS_NAN || k == DS_OOB || k == DS_SET_STRIDE_DEV + DS_NAN || k == DS_SET_STRIDE_DEV + DS_OOB) ? IDX_NONE : 0;
}
// zero fills as kernels for the same reason (cudaMemsetAsync may be served by a DMA engine)
// (dscal != nullptr: block 0 also initialises the device scalar block, which saves the k_dscal_init launch of a build)
static __global__ void __launch_bounds__(256) k_zero_ints(int* __restrict__ p, long long n, int* __restrict__ dscal = nullptr) {
    if (dscal && blockIdx.x == 0 && threadIdx.x < DS_COUNT) {
        const int k = threadIdx.x;
        dscal[k] = (k == DS_NAN || k == DS_OOB || k == DS_SET_STRIDE_DEV + DS_NAN || k == DS_SET_STRIDE_DEV + DS_OOB) ? IDX_NONE : 0;
    }
    const long long n4 = n >> 2;
    int4* p4 = reinterpret_cast<int4*>(p);
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n4; k += (long long)gridDim.x * blockDim.x) p4[k] = make_int4(0, 0, 0, 0);
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) p[(n4 << 2) + threadIdx.x] = 0;
}
// out[0] = max(cnt[0], cnt[1]) (halo face counts: the value the ranks all-reduce, clm_comm.cu)
static __global__ void k_max2(const int* __restrict__ cnt, int* __restrict__ out) {
    if (threadIdx.x == 0) out[0] = max(cnt[0], cnt[1]);
}
static __global__ void k_map_begin(unsigned long long* __restrict__ res_words, int nwords, int* __restrict__ work) {
    if (threadIdx.x < nwords) res_words[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) *work = 0;
}
// n device ints -> mapped pinned host memory (clm_read_ints: a read-back that does not queue on the DMA engines)
static __global__ void k_publish_ints(const int* __restrict__ src, int n, volatile int* __restrict__ h_dst) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) h_dst[k] = src[k];
    __threadfence_system();
}
static __global__ void k_dscal_publish(const int* __restrict__ dscal, volatile int* __restrict__ h_pub) {
    const int k = threadIdx.x;
    if (k < DS_COUNT) h_pub[k] = dscal[k];
    __threadfence_system();
}

// p = rotation * (M * frac(M \ x)), every operation rounded separately in T
// (CellLists.jl:944-945; CellOperations.jl:56-66, :91-94).  Non-periodic: coordinates are used as given.
template <class T, int DIM> __device__ __forceinline__ void place_particle(const GeomT<T>& g, const T* x, T p[3]) {
    p[2] = T(0);
    if (g.cell_type == CLM_NONPERIODIC_CT) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) p[k] = x[k];
        return;
    }
    T f[DIM];
    if (DIM == 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            f[k] = xdiv(xadd(xadd(xmul(g.cof[3 * k], x[0]), xmul(g.cof[3 * k + 1], x[1])), xmul(g.cof[3 * k + 2], x[2])), g.det);
    } else {
        f[0] = xdiv(xsub(xmul(g.cof[0], x[0]), xmul(g.cof[1], x[1])), g.det);
        f[1] = xdiv(xsub(xmul(g.cof[3], x[1]), xmul(g.cof[4], x[0])), g.det);
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        f[k] = xsub(f[k], floor(f[k]));
        if (f[k] == T(1)) f[k] = T(0);
    }
    T w[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        T s = xmul(g.m[3 * k], f[0]);
#pragma unroll
        for (int c = 1; c < DIM; ++c) s = xadd(s, xmul(g.m[3 * k + c], f[c]));
        w[k] = s;
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        T s = xmul(g.rot[3 * k], w[0]);
#pragma unroll
        for (int c = 1; c < DIM; ++c) s = xadd(s, xmul(g.rot[3 * k + c], w[c]));
        p[k] = s;
    }
}

// Cell of a point on the reference grid and on the device grid.  Reference cell: floor((p - cb_min)/cs) per
// dimension, real particles nudged off the outermost layers (particle_cell Box.jl:499-508,
// real_particle_border_case CellLists.jl:956-967).  Device cell: reference cell * sub + sub-cell, the sub-cell
// from the fractional part of the same quotient, so the device grid is EXACTLY nested in the reference grid.
//
// DEVICE linearisation: the LAST reference dimension runs fastest (lin = c[N-1] + n[N-1]*(c[N-2] + ...)),
// the transpose of the reference's column-major cell_linear_index (CellOperations.jl:256-257).  With it
// the reference's forward stencil (Box.jl:436-457: d1 > 0 | d1 == 0, d2 > 0 | d1 == d2 == 0, d3 > 0)
// is "later cells of the own row + every row with a larger row index", so the sweep can visit each
// reference-cell pair from the SAME home cell as the reference and therefore evaluates the same periodic
// image of every pair (bit-identical d2).  Returns false when the point is outside the grid (or NaN/Inf).
template <class T, int DIM>
__device__ __forceinline__ bool cell_of(const GeomT<T>& g, const T p[3], bool real, int& dev_lin, int& ref_lin, int* ref_fast = nullptr) {
    dev_lin = 0; ref_lin = 0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        const T u = xdiv(xsub(p[k], g.cb_min[k]), g.cs[k]);
        const T q = floor(u);
        if (!(q >= T(-1) && q <= T(g.nc[k]))) return false;  // also rejects NaN / Inf
        int c = (int)q, sc = 0;
        if (g.sub > 1) sc = min(max((int)((u - q) * T(g.sub)), 0), g.sub - 1);
        if (real) {
            if (c == g.lcell - 1) { c += 1; sc = 0; }
            if (c == g.nc[k] - g.lcell) { c -= 1; sc = g.sub - 1; }
        }
        if (c < 0 || c >= g.nc[k]) return false;
        ref_lin = ref_lin * g.nc[k] + c;
        if (k == DIM - 1 && ref_fast) *ref_fast = c;   // reference cell along the row of the device grid
        dev_lin = dev_lin * (g.nc[k] * g.sub + ((k == DIM - 1) ? 1 : 0)) + c * g.sub + sc;   // row pitch nx + 1
    }
    return true;
}

// Count pass (SCATTER = false): per-cell histogram of real + image particles.  Scatter pass: records to the
// atomic per-cell cursor.  cell_nact[c] flags cells holding a record of a particle this rank owns (real or image);
// ref_real[] flags reference cells with a real particle.  A record can act as particle i of a pair when it is owned and
// lives in a REFERENCE cell that holds a real particle (the reference sweeps exactly those cells, self.jl:56-57):
// k_rows combines the two flags once the count pass is complete.
__device__ __forceinline__ float shfl_t(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_t(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <class T, int DIM, bool SCATTER>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 8 : 4)
k_bin(const __grid_constant__ GeomT<T> g, const T* __restrict__ pos, const T* __restrict__ fpos, int n, int n_own, int* __restrict__ cell_cursor,
      int* __restrict__ cell_nact, int* __restrict__ ref_real, RecT<T>* __restrict__ rec, int* __restrict__ slot_of, int rec_cap, int* __restrict__ dscal, const unsigned char* __restrict__ fmask) {
    typedef TagT<T> TG;
    const int lane = threadIdx.x & 31;
    constexpr int CENTER = (DIM == 3) ? 13 : 4;
    // grid-stride over blocks of 256 particles: the grid can be capped (clm_set_option "bin_blocks_per_sm"); the default is one
    // block per 256 particles, which the hardware block scheduler balances better than a capped persistent grid (measured)
    for (int ip0 = blockIdx.x * blockDim.x; ip0 < n; ip0 += gridDim.x * blockDim.x) {
    const int ip = ip0 + threadIdx.x;
    T p[3] = {T(0), T(0), T(0)};
    unsigned okmask = 0u;      // candidate images of this lane's particle (none for invalid / interior / non-periodic)
    bool is_foreign = false;   // owned by another rank: the rows behind the owned ones, or flagged by clm_set_foreign_mask
    if (ip < n) {
        T x[DIM];
        bool bad = false;
        const T* src = (ip < n_own) ? pos + (size_t)ip * DIM : fpos + (size_t)(ip - n_own) * DIM;   // owned particles, then foreign ones
#pragma unroll
        for (int k = 0; k < DIM; ++k) { x[k] = src[k]; bad |= (x[k] != x[k]); }
        int lin = 0, rlin = 0, cfast = 0;
        if (!SCATTER && g.np_check && !bad) {
            bool fit = true;
#pragma unroll
            for (int k = 0; k < DIM; ++k) fit = fit && (x[k] >= g.np_lo[k]) && (x[k] <= g.np_hi[k]);
            if (!fit) dscal[DS_NOFIT] = 1;     // the reused box does not hold this particle: the host recomputes the limits
        }
        if (bad) {
            if (!SCATTER) atomicMin(&dscal[DS_NAN], ip);     // _validate_coordinates, CellOperations.jl:6-21
        } else {
            place_particle<T, DIM>(g, x, p);
            if (!cell_of<T, DIM>(g, p, true, lin, rlin, &cfast)) { if (!SCATTER) atomicMin(&dscal[DS_OOB], ip); bad = true; }
        }
        if (!bad) {
            // cell_cursor: the per-cell histogram in the count pass; the per-cell write cursor (pre-loaded with the
            // exclusive starts) in the scatter pass
            is_foreign = (ip >= n_own) || (fmask && fmask[ip] != 0);
            const typename TG::type foreign = is_foreign ? TG::FOREIGN : (typename TG::type)0;
            const int slot = atomicAdd(&cell_cursor[lin], 1);
            if (!SCATTER) {
                if (!foreign) cell_nact[lin] = 1;   // flags: plain stores, every writer stores the same value
                ref_real[rlin] = 1;
            } else if (slot < rec_cap) {
                strec(&rec[slot], p[0], p[1], p[2], (typename TG::type)ip | TG::HOME | foreign);
            }
            if (SCATTER) slot_of[ip] = slot;     // slot of the particle's real record (force gather of the N3 sweep, k_twin)
            // replicate_particle! (Box.jl:556-566): images x + aligned_cell*idx, idx in {-1,0,1}^N \ {0}, kept iff inside the
            // computing box [cb_min, cb_max).  Orthorhombic cells: the shift of image index (i1,i2,i3) is (i1*L1, i2*L2, i3*L3)
            // exactly, so which indices can land inside the computing box is decided per dimension.
            if (g.cell_type != CLM_NONPERIODIC_CT) {
                okmask = (DIM == 3) ? 0x7ffffffu : 0x1ffu;
                if (!g.rotated && g.cell_type == CLM_ORTHO_CT) {
                    unsigned dimok[3] = {2u, 2u, 2u};   // bit (idx+1): idx allowed
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const T lo = xadd(p[k], g.shift[(k == 0) ? CENTER - 1 : (k == 1 ? CENTER - 3 : CENTER - 9)][k]);
                        const T hi = xadd(p[k], g.shift[(k == 0) ? CENTER + 1 : (k == 1 ? CENTER + 3 : CENTER + 9)][k]);
                        if (g.cb_min[k] <= lo && lo < g.cb_max[k]) dimok[k] |= 1u;
                        if (g.cb_min[k] <= hi && hi < g.cb_max[k]) dimok[k] |= 4u;
                    }
                    // okmask bit (i0 + 3 i1 + 9 i2) = dimok0[i0] & dimok1[i1] & dimok2[i2], built by replication
                    const unsigned m0 = dimok[0] & 7u;
                    const unsigned m01 = ((dimok[1] & 1u) ? m0 : 0u) | ((dimok[1] & 2u) ? (m0 << 3) : 0u) | ((dimok[1] & 4u) ? (m0 << 6) : 0u);
                    okmask = (DIM == 2) ? m01 : (((dimok[2] & 1u) ? m01 : 0u) | ((dimok[2] & 2u) ? (m01 << 9) : 0u) | ((dimok[2] & 4u) ? (m01 << 18) : 0u));
                } else {
                    // triclinic / rotated cells: every lane tests its own 3^N - 1 shifts against the computing box (the same
                    // additions and comparisons as below, so the same images survive): straight-line code, and the dealing
                    // loop below then runs once per ~32 ACTUAL images instead of once per 32 candidates (26 trips per warp)
                    unsigned m = 0u;
#pragma unroll
                    for (int img = 0; img < ((DIM == 3) ? 27 : 9); ++img) {
                        bool in = true;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            const T qk = xadd(p[k], g.shift[img][k]);
                            in = in && (g.cb_min[k] <= qk) && (qk < g.cb_max[k]);
                        }
                        if (in) m |= 1u << img;
                    }
                    okmask = m;
                }
                okmask &= ~(1u << CENTER);
            }
        }
    }
    // The (particle, image) candidates of the WARP are dealt evenly to its lanes: a warp next to a cell face holds a
    // handful of candidates in a few lanes, and would otherwise run as many divergent iterations as its busiest lane.
    const int cnt = __popc(okmask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - cnt;
    for (int w0 = 0; w0 < total; w0 += 32) {
        const int w = w0 + lane;
        int s = 0;   // source lane: the first lane whose inclusive count exceeds w
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, s + step - 1); if (v <= w) s += step; }
        s = min(s, 31);
        const int k_th = w - __shfl_sync(0xffffffffu, excl, s);
        const unsigned m = __shfl_sync(0xffffffffu, okmask, s);
        const int ips = __shfl_sync(0xffffffffu, ip, s);
        const bool fsrc = __shfl_sync(0xffffffffu, is_foreign ? 1 : 0, s) != 0;
        const T px = shfl_t(p[0], s), py = shfl_t(p[1], s), pz = shfl_t(p[2], s);
        if (w >= total) continue;
        const int img = (int)__fns(m, 0u, k_th + 1);
        const T ps[3] = {px, py, pz};
        T q[3] = {T(0), T(0), T(0)};
        bool in = true;
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            q[k] = xadd(ps[k], g.shift[img][k]);
            in = in && (g.cb_min[k] <= q[k]) && (q[k] < g.cb_max[k]);
        }
        if (!in) continue;
        int lq, rq;
        if (!cell_of<T, DIM>(g, q, false, lq, rq)) continue;
        const typename TG::type foreign = fsrc ? TG::FOREIGN : (typename TG::type)0;
        const int qslot = atomicAdd(&cell_cursor[lq], 1);
        if (!SCATTER && !foreign) cell_nact[lq] = 1;
        if (SCATTER && qslot < rec_cap) {
            const bool home = ref_real[rq] != 0;
            strec(&rec[qslot], q[0], q[1], q[2], (typename TG::type)ips | TG::GHOST | foreign | (home ? TG::HOME : (typename TG::type)0));
        }
    }
    }
}

// ---- rows: starts + tiles ----------------------------------------------------------------------------------
// One warp per row of device cells, between the count pass and the scatter pass.
//  * starts: cs = cell_start + 1 is the cursor array of the scatter pass: the cursor of cell x of a row is
//    cs[row * px + x], pre-loaded with the cell's first record; once every record is placed it holds the cell's END, i.e.
//    cell_start[row * px + x + 1] = start of cell x + 1, and cell_start[row * px] (never incremented) stays the start
//    of the row: no second counter array.  The row's base comes from one atomicAdd on the record counter.
//  * tiles (make_tiles): first / last cell holding a record that can act as particle i (owned record in a reference
//    cell with a real particle) -> the record range [start(first), end(last)) cut into tiles of tile_i records, appended
//    to the tile array with one atomicAdd per row.  Tiles only need the cell starts, not the records: building them here
//    saves a launch and a pass over the per-cell arrays after the scatter.
//  * the threads also count the reference cells holding a real particle (CellList.n_cells_with_real_particles).
template <int DIM>
static __global__ void __launch_bounds__(256)
k_rows(const int* __restrict__ cell_count, const int* __restrict__ cell_own, const int* __restrict__ ref_real, int* __restrict__ cell_start,
       int nx, int ny, int nrows, int sub, int nc1, int nc2, int nref, int make_tiles, int tile_i, Tile* __restrict__ tiles, int tiles_cap,
       int* __restrict__ dscal) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = gtid >> 5, lane = threadIdx.x & 31;
    {
        int c = 0;
        for (int i = gtid; i < nref; i += gridDim.x * blockDim.x) c += (ref_real[i] != 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(&dscal[DS_NCELLS_REAL], c);
    }
    if (row >= nrows) return;
    const int px = nx + 1;
    const int* cnt = cell_count + (size_t)row * px;
    const int* own = cell_own + (size_t)row * px;
    int* cs = cell_start + (size_t)row * px;
    // reference cells of this row: reference linear index = (c0 * nc1 + c1) * nc2 + x / sub (cell_of), row = (c0, c1) sub-cells
    const int rrow = (DIM == 3) ? ((row / ny) / sub) * nc1 + (row % ny) / sub : row / sub;
    const int* rr = ref_real + (size_t)rrow * nc2;
    int total = 0;
    for (int c = lane; c < nx; c += 32) total += cnt[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    int base = 0;
    if (lane == 0) base = (total > 0) ? atomicAdd(&dscal[DS_NTOT], total) : 0;
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane == 0) cs[0] = base;
    int first = IDX_NONE, last = -1, a = 0, b = 0, run = base;
    for (int c0 = 0; c0 < nx; c0 += 32) {
        const int c = c0 + lane;
        const int v = (c < nx) ? cnt[c] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const int start = run + inc - v;
        if (c < nx) {
            cs[c + 1] = start;    // cursor of cell c = its first record
            if (make_tiles && v > 0 && own[c] != 0 && rr[c / sub] != 0) {
                if (c < first) { first = c; a = start; }
                if (c > last) { last = c; b = start + v; }
            }
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (!make_tiles) return;
    // lanes hold ascending cells: the smallest first / largest last win, with their record bounds
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int f2 = __shfl_xor_sync(0xffffffffu, first, o), a2 = __shfl_xor_sync(0xffffffffu, a, o);
        const int l2 = __shfl_xor_sync(0xffffffffu, last, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
        if (f2 < first) { first = f2; a = a2; }
        if (l2 > last) { last = l2; b = b2; }
    }
    if (last < 0) return;
    const int ntiles = (b - a + tile_i - 1) / tile_i;
    int tbase = 0;
    if (lane == 0 && ntiles > 0) tbase = atomicAdd(&dscal[DS_NTILES], ntiles);
    tbase = __shfl_sync(0xffffffffu, tbase, 0);
    __syncwarp();   // the starts this warp wrote are read back below
    // cell of record k: the last cell c in [first, last] with start(c) = cs[c + 1] <= k
    auto cell_x = [&](int k) { int lo = first, hi = last; while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (cs[mid + 1] <= k) lo = mid; else hi = mid - 1; } return lo; };
    for (int t = lane; t < ntiles; t += 32) {
        const int k0 = a + t * tile_i;
        const int n = min(tile_i, b - k0);
        Tile tl;
        tl.k0 = k0; tl.cnt = n; tl.yz = (row % ny) | ((row / ny) << 16);
        tl.cx = cell_x(k0) | (cell_x(k0 + n - 1) << 16);
        if (tbase + t < tiles_cap) tiles[tbase + t] = tl;   // beyond the capacity only when the record capacity overflowed: the build is repeated
    }
}

// ---- slot-tagged twin of the records (on request: Newton's-third-law force sweep, clm_sweep_n3.cuh) ---------------------
// One coalesced pass in record order: rec_n3[k] = (position of rec[k], slot of the particle's REAL record -- an image
// points at its original through slot_of[]; by_index: the particle index, triclinic cells -- | GHOST | HOME | parity of
// the record's reference cell along the row << 29), and the force accumulator row of the slot is zeroed in passing.  The
// reference cell along the row is recomputed from the stored coordinate with the arithmetic of cell_of (one IEEE
// division): cheaper than a second scattered store stream in k_bin, which is a DRAM read-modify-write per record once the
// arrays outgrow the L2.
template <class T, int DIM>
__global__ void __launch_bounds__(256)
k_twin(const __grid_constant__ GeomT<T> g, const RecT<T>* __restrict__ rec, const int* __restrict__ slot_of, const int* __restrict__ ntot, int rec_cap,
       RecT<T>* __restrict__ rec_n3, T* __restrict__ facc, int by_index) {
    typedef TagT<T> TG;
    typedef typename TG::type tag_t;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (*ntot > rec_cap || k >= *ntot) return;   // overflowed build: repeated by the host
    const RecT<T> r = ldrec(rec + k);
    const bool ghost = (r.tag & TG::GHOST) != 0;
    const int idx = (int)(r.tag & TG::MASK);
    constexpr int kf = DIM - 1;                  // the row of the device grid runs along the last reference dimension
    const T pf = (kf == 0) ? r.x : ((kf == 1) ? r.y : r.z);
    int c = (int)floor(xdiv(xsub(pf, g.cb_min[kf]), g.cs[kf]));
    if (!ghost) {                                // border nudge of real particles (cell_of)
        if (c == g.lcell - 1) c += 1;
        if (c == g.nc[kf] - g.lcell) c -= 1;
    }
    const unsigned word = (unsigned)(by_index ? idx : (ghost ? slot_of[idx] : k)) | (ghost ? 0x80000000u : 0u) | ((r.tag & TG::HOME) ? 0x40000000u : 0u) |
                          ((unsigned)(c & 1) << 29);
    strec(&rec_n3[k], r.x, r.y, r.z, (tag_t)word);
    strec(reinterpret_cast<RecT<T>*>(facc) + k, T(0), T(0), T(0), (tag_t)0);
}

// reference-cell index of every particle along reference dimension `axis` (0-based, after wrapping and the
// border nudge) -- what a slab decomposition uses to decide ownership and halo membership with exactly the
// arithmetic of the cell-list build.  -1: invalid coordinate.
template <class T, int DIM>
__global__ void __launch_bounds__(256)
k_cell_coord(const __grid_constant__ GeomT<T> g, const T* __restrict__ pos, int n, int axis, int* __restrict__ out) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= n) return;
    T x[DIM], p[3];
#pragma unroll
    for (int k = 0; k < DIM; ++k) x[k] = pos[(size_t)ip * DIM + k];
    place_particle<T, DIM>(g, x, p);
    const T q = floor(xdiv(xsub(p[axis], g.cb_min[axis]), g.cs[axis]));
    int c = (q >= T(-1) && q <= T(g.nc[axis])) ? (int)q : -1;
    if (c == g.lcell - 1) c += 1;
    if (c == g.nc[axis] - g.lcell) c -= 1;
    out[ip] = c;
}

// face selection of a slab decomposition in ONE pass: particles whose reference-cell layer along `axis` lies in
// [r.x, r.y) are appended to list A, those in [r.z, r.w) to list B (merge != 0: either range -> list A, once).  Warp-aggregated
// atomics; rows beyond `capacity` are counted but not written (the caller checks the counts).
template <class T, int DIM>
__global__ void __launch_bounds__(256)
k_select_layers(const __grid_constant__ GeomT<T> g, const T* __restrict__ pos, int n, int axis, int4 r, int merge,
                T* __restrict__ out_a, T* __restrict__ out_b, int capacity, int* __restrict__ counts, int* __restrict__ idx_a, int* __restrict__ idx_b) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    bool in_a = false, in_b = false;
    T x[DIM];
    if (ip < n) {
        T p[3];
#pragma unroll
        for (int k = 0; k < DIM; ++k) x[k] = pos[(size_t)ip * DIM + k];
        place_particle<T, DIM>(g, x, p);
        const T q = floor(xdiv(xsub(p[axis], g.cb_min[axis]), g.cs[axis]));
        int c = (q >= T(-1) && q <= T(g.nc[axis])) ? (int)q : -1;
        if (c == g.lcell - 1) c += 1;
        if (c == g.nc[axis] - g.lcell) c -= 1;
        in_a = (c >= r.x && c < r.y);
        in_b = (c >= r.z && c < r.w);
        if (merge) { in_a = in_a || in_b; in_b = false; }
    }
    auto append = [&](bool take, T* out, int* counter, int* idx_out) {
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m == 0u) return;
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(counter, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (take && slot < capacity) {
#pragma unroll
            for (int k = 0; k < DIM; ++k) out[(size_t)slot * DIM + k] = x[k];
            if (idx_out) idx_out[slot] = ip;   // source row: the caller gathers the particles' side data (ids, weights, ...) with it
        }
    };
    append(in_a, out_a, counts, idx_a);
    append(in_b, out_b, counts + 1, idx_b);
}

// per-block min/max of the coordinates (limits(), CellOperations.jl:262-324): out[b][0..2] = min, [3..5] = max
template <class T, int DIM>
__global__ void __launch_bounds__(256) k_minmax(const T* __restrict__ pos, int n, T* __restrict__ out, int* __restrict__ dscal) {
    __shared__ T smin[8][3], smax[8][3];
    T lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = CUDART_INF_T<T>(); hi[k] = -CUDART_INF_T<T>(); }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const T v = pos[(size_t)i * DIM + k];
            if (v != v) atomicMin(&dscal[DS_NAN], i);
            lo[k] = fmin(lo[k], v); hi[k] = fmax(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) for (int k = 0; k < 3; ++k) { smin[w][k] = lo[k]; smax[w][k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        T a = smin[0][k], b = smax[0][k];
        for (int q = 1; q < 8; ++q) { a = fmin(a, smin[q][k]); b = fmax(b, smax[q][k]); }
        out[blockIdx.x * 6 + k] = a; out[blockIdx.x * 6 + 3 + k] = b;
    }
}

// gather per-particle auxiliary data (weights, velocities) into record order, one record-sized slot (4 x T) per
// record, so that the sweep stages it with the same bulk copies as the records; velocities are rotated into the
// aligned frame (dot(v, R^-1 d) == dot(R v, d)).
template <class T>
static __global__ void __launch_bounds__(256)
k_gather_aux(const RecT<T>* __restrict__ rec, int n_tot, const T* __restrict__ aux, int ncomp, const __grid_constant__ GeomT<T> g,
             int rotate, T* __restrict__ out) {
    typedef TagT<T> TG;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_tot) return;
    const size_t idx = (size_t)(rec[k].tag & TG::MASK);
    if (ncomp == 1) { out[(size_t)k * 4 + 0] = aux[idx]; out[(size_t)k * 4 + 1] = T(0); out[(size_t)k * 4 + 2] = T(0); out[(size_t)k * 4 + 3] = T(0); return; }
    T v[3] = {T(0), T(0), T(0)};
    for (int c = 0; c < ncomp; ++c) v[c] = aux[idx * ncomp + c];
    if (rotate) {
        T r[3];
        for (int a = 0; a < 3; ++a) r[a] = g.rot[3 * a] * v[0] + g.rot[3 * a + 1] * v[1] + g.rot[3 * a + 2] * v[2];
        for (int a = 0; a < 3; ++a) v[a] = r[a];
    }
    // stored as 4 components per record (x, y, z, 0): one 128-bit (F32) / two 128-bit (F64) loads
    out[(size_t)k * 4 + 0] = v[0]; out[(size_t)k * 4 + 1] = v[1]; out[(size_t)k * 4 + 2] = v[2]; out[(size_t)k * 4 + 3] = T(0);
}

}  // namespace clm

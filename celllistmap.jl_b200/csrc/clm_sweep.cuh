// Pair sweep: the B200 replacement of _pairwise! (src/internals/self.jl:28-184, cross.jl:8-129,
// vicinal_cells.jl:4-75, NonPeriodicCells.jl:281-352) for the compiled-in functor catalogue.
//
// Work decomposition (B200-first, not the reference's cell-by-cell task batches):
//   * the device grid splits every reference cell into sub^N sub-cells (clm_build.cuh) so that the candidate
//     volume around a group of particles approaches the cutoff sphere instead of 27 cutoff-sized cells.
//   * a work item ("tile") is TI consecutive records of one row of the cell-sorted array; one warp owns a
//     tile: lane -> (i-slot = lane % TI, j-slice = lane / TI).  Particle i lives in registers.
//   * for every stencil row the candidate partners are ONE contiguous record range
//     [cell_start[first - w], cell_start[last + w + 1]) because cells are linearised along the row; the
//     half-width w per row offset comes from a host table (rows farther than the cutoff are skipped).
//   * the row ranges of a tile are bulk-copied (TMA 1-D copies, one per row, issued by the lane that classified the
//     row) into the warp's shared-memory staging buffer, culled against the bounding box of the tile's particles and
//     compacted in place; one flat loop then evaluates 32 pairs per warp step, every lane of a j-slice reading the
//     same staged record (128-bit broadcast LDS).
//   * warps fetch tiles from a global atomic counter, one tile ahead (persistent grid = SMs x resident CTAs).
//   * exactly-once rules are the reference's, expressed on REFERENCE cells (device cell / sub):
//     MODE_HALF (orthorhombic / non-periodic self): partner's reference cell is lexicographically after the
//     home reference cell -- the reference's forward stencil, Box.jl:436-457 -- or the same cell and a later
//     record; real_i | real_j (self.jl:143-161, vicinal_cells.jl:33); home cells hold a real particle.
//     MODE_TRI (triclinic self): full stencil, i real, index_i < index_j (self.jl:164-184, vicinal_cells.jl:53-65).
//     MODE_ALL (two-set: full stencil, i real, cross.jl:111-129; also the full-shell force sweep, where each
//     ordered pair contributes to f_i only so that per-particle outputs need no atomics and no per-batch
//     output copies, the reference's A16).
//   * the distance test of the exact functors is bit-identical to the oracle's:
//     d2 = (dx*dx + dy*dy) + dz*dz, unfused, <=.  Force functors (tolerance parity) contract it to FMAs.
#pragma once
#include "clm_common.cuh"
#include "clm_build.cuh"

namespace clm {

enum { MODE_HALF = 0, MODE_TRI = 1, MODE_ALL = 2 };
#ifndef CLM_SWEEP_THREADS
#define CLM_SWEEP_THREADS 128
#endif
constexpr int SWEEP_THREADS = CLM_SWEEP_THREADS;   // warps of a CTA work independently (own staging buffer, own mbarrier): the CTA size only sets the register / occupancy granularity
static_assert(SWEEP_THREADS % 32 == 0 && SWEEP_THREADS >= 32 && SWEEP_THREADS <= 128, "block_sum scratch holds 4 warps");
constexpr int NB_PRIV_MAX = 16;  // histograms with <= this many bins use per-thread private shared-memory bins
constexpr int LF_MAX = 15;       // largest stencil reach in device cells (lcell * sub): the stencil tables of SweepArgs hold (2 LF_MAX + 1)^2 rows
constexpr int SUB_MAX = 7;       // largest sub-cell split chosen for a reference cell (lcell * sub <= SUB_MAX unless lcell itself is larger)
constexpr int TILE_I = 8, LOG2_TILE_I = 3, NSLICE = 32 / TILE_I;   // particles i per warp tile; the other lanes split the partners into j-slices
// per-warp staging buffer of partner records: small buffers keep 8 CTAs (32 warps) resident per SM, which hides the
// latency of the tile fetch / cell_start loads / bulk copies better than fewer, larger chunks (tools/tune_stage.sh)
#ifndef CLM_STAGE_BYTES_F32
#define CLM_STAGE_BYTES_F32 6144
#endif
#ifndef CLM_STAGE_BYTES_F64
#define CLM_STAGE_BYTES_F64 8192
#endif
// functors that stage a per-record side array next to the records (AUX) hold two buffers per warp: smaller ones keep
// more CTAs resident (measured, pair-velocity map F64 1M galaxies: 8 KB 25.9 ms, 6 KB 21.8 ms, 4 KB 22.6 ms)
#ifndef CLM_STAGE_BYTES_F32_AUX
#define CLM_STAGE_BYTES_F32_AUX 6144
#endif
#ifndef CLM_STAGE_BYTES_F64_AUX
#define CLM_STAGE_BYTES_F64_AUX 6144
#endif
template <class T, bool AUX = false> struct StageBytes {
    static constexpr int value = (sizeof(T) == 4) ? (AUX ? CLM_STAGE_BYTES_F32_AUX : CLM_STAGE_BYTES_F32) : (AUX ? CLM_STAGE_BYTES_F64_AUX : CLM_STAGE_BYTES_F64);
};
template <class T, bool AUX = false> struct StageTotal { static constexpr int value = (SWEEP_THREADS / 32) * StageBytes<T, AUX>::value + 64; };   // + one mbarrier per warp; functor shared memory follows

// result block: accumulators every map kernel adds into (zeroed before the launch)
enum { RB_ENERGY = 0, RB_SUM_D = 1, RB_SUM_D2 = 2, RB_F64_COUNT = 8 };
enum { RC_NPAIRS = 0, RC_NBAND = 1, RC_NLIST = 2, RC_I64_COUNT = 8 };
struct ResultBlock {
    double f[RB_F64_COUNT];
    unsigned long long c[RC_I64_COUNT];
};

template <class T> struct SweepArgs {
    const RecT<T>* rec_i;
    const RecT<T>* rec_j;
    const int* cell_start_i;
    const int* cell_start_j;
    const Tile* tiles;
    int* dscal;          // DS_NTILES (read), DS_WORK (atomic tile counter)
    ResultBlock* res;
    int nx, ny, nz;      // device cells along the fast / middle / slow axis = reference dims (3,2,1) in 3-D, (2,1,-) in 2-D
    int lf, sub, self;   // lf = lcell * sub: stencil reach in device cells
    int rec_cap_i, rec_cap_j;   // record capacities: a build whose record count overflowed them is repeated by the host,
                                // and the sweep queued behind it must not touch the truncated arrays
    unsigned sub_magic;  // ceil(2^32 / sub): x / sub == __umulhi(x, sub_magic) for the cell indices in use
    T rc2;
    signed char hw[(2 * LF_MAX + 1) * (2 * LF_MAX + 1)];   // [dslow + lf][dmid + lf]: half-width along the row, -1 = skip
    signed char rdz[(2 * LF_MAX + 1) * (2 * LF_MAX + 1)], rdy[(2 * LF_MAX + 1) * (2 * LF_MAX + 1)];   // stencil row -> (dslow, dmid)
};

template <class T> struct Ctx {   // what a functor sees for the tile in flight
    int ki;          // record index of particle i
    int lane, islot, slice;
    bool active;     // this lane holds a particle that may act as i
    RecT<T> ri;
};

template <class T> __device__ __forceinline__ T block_sum(T v, T* smem /* >= 4 */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    T r = T(0);
    if (threadIdx.x == 0) for (int k = 0; k < SWEEP_THREADS / 32; ++k) r += smem[k];
    return r;  // valid on thread 0
}

// ===================================================================================================
// functor catalogue.  Every functor: Acc (per thread, whole kernel), IAcc (per
// particle i, one tile), init / begin / pair / end (folds the tile into Acc) / finish.  pair() is called by all 32 lanes with a
// `hit` predicate so that warp collectives inside it are legal.
// ===================================================================================================

// f1/f2 test functor (test/modules/Testing.jl:23-26) + pair count + 1-ulp cutoff band count
template <class T> struct FSum {
    T rc2_lo, rc2_hi;   // prevfloat/nextfloat of cutoff^2
    struct Acc { double sd, sd2; unsigned long long n, band; };
    struct IAcc {};
    static constexpr bool NEEDS_BAND = true, EXACT_D2 = true, AUX = false;
    __device__ void init(Acc& a) const { a.sd = 0; a.sd2 = 0; a.n = 0; a.band = 0; }
    __device__ void begin(IAcc&, const Ctx<T>&) const {}
    __device__ void pair(Acc& a, IAcc&, const Ctx<T>&, bool hit, bool ok, const RecT<T>&, int, T, T, T, T d2) const {
        if (hit) { a.sd += (double)xsqrt(d2); a.sd2 += (double)d2; a.n += 1; }
        if (ok && d2 >= rc2_lo && d2 <= rc2_hi) a.band += 1;
    }
    __device__ void end(Acc&, IAcc&, const Ctx<T>&) const {}
    __device__ void finish(Acc& a, ResultBlock* res) const {
        __shared__ double sm[4];
        __shared__ unsigned long long smc[4];
        double sd = block_sum(a.sd, sm), sd2 = block_sum(a.sd2, sm);
        unsigned long long n = block_sum(a.n, smc), band = block_sum(a.band, smc);
        if (threadIdx.x == 0) {
            atomicAdd(&res->f[RB_SUM_D], sd); atomicAdd(&res->f[RB_SUM_D2], sd2);
            atomicAdd(&res->c[RC_NPAIRS], n); atomicAdd(&res->c[RC_NBAND], band);
        }
    }
};

template <class T> __device__ __forceinline__ T fast_rcp(T x);
template <> __device__ __forceinline__ float fast_rcp<float>(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// FP64: hardware seed (MUFU.RCP64H, ~20 bits) + one cubic correction r0 * (1 + e + e^2), e = 1 - x r0: relative error
// ~2^-60, three DFMAs.  The IEEE division 1.0 / x costs two more DFMAs and a slow-path branch for nothing the 1e-10
// tolerance of the force maps could see.
template <> __device__ __forceinline__ double fast_rcp<double>(double x) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}
// 1/sqrt(x) for the tolerance-parity sums: hardware seed + one cubic correction y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2
template <class T> __device__ __forceinline__ T fast_rsqrt(T x);
template <> __device__ __forceinline__ float fast_rsqrt<float>(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <> __device__ __forceinline__ double fast_rsqrt<double>(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}
template <class T> __device__ __forceinline__ T xfma(T a, T b, T c);
template <> __device__ __forceinline__ float xfma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double xfma<double>(double a, double b, double c) { return fma(a, b, c); }

// write per-particle force accumulators: combine the j-slices, rotate back to the input frame
// (pair.x/pair.y are inv_rotation * coordinates, self.jl:171-178), store once per real particle.
template <class T> struct ForceOut {
    T* forces;        // n x dim, AoS
    int dim, accumulate, rotated;
    T inv_rot[9];
    __device__ __forceinline__ void store(const Ctx<T>& c, T fx, T fy, T fz) const {
#pragma unroll
        for (int o = TILE_I; o < 32; o <<= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (!c.active || c.slice != 0) return;
        if (rotated) {
            const T a = inv_rot[0] * fx + inv_rot[1] * fy + inv_rot[2] * fz;
            const T b = inv_rot[3] * fx + inv_rot[4] * fy + inv_rot[5] * fz;
            const T d = inv_rot[6] * fx + inv_rot[7] * fy + inv_rot[8] * fz;
            fx = a; fy = b; fz = d;
        }
        T* f = forces + (size_t)(c.ri.tag & TagT<T>::MASK) * dim;
        if (accumulate) { f[0] += fx; f[1] += fy; if (dim == 3) f[2] += fz; }
        else { f[0] = fx; f[1] = fy; if (dim == 3) f[2] = fz; }
    }
};

// Lennard-Jones c12/d2^6 - c6/d2^3 (test/applications/gromacs/compare_with_gromacs.jl:9-13), optional forces.
// With q = (c12/c6) / d2^3 the pair energy is (c6^2/c12) * (q^2 - q) and the scalar force over d is
// (6 c6^2/c12) * (2 q^2 - q) / d2: the constants leave the inner loop (applied once per tile / per kernel) and
// the pair costs rcp + 10 FP32 instructions.  Pure c6 or pure c12 potentials use the direct form.
template <class T, bool FORCES, bool NORM> struct FLJ {
    T c6, c12;
    T s2, escale, fscale;   // NORM: s2 = cbrt(c12/c6), escale = c6^2/c12, fscale = 6 c6^2/c12; !NORM: the direct form
    ForceOut<T> fo;
    // the energy is summed in T only over ONE tile (a few hundred terms per thread) and folded into a double per thread
    // at the end of the tile: a Float32 accumulator that lives for the whole persistent kernel loses the 1e-5 bar on
    // million-particle systems (thousands of mixed-sign terms per thread)
    struct Acc { double e; };
    struct IAcc { T fx, fy, fz, e; };
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = !FORCES, AUX = false;   // the full-shell force sweep has tolerance parity
    __device__ void init(Acc& a) const { a.e = 0.0; }
    __device__ void begin(IAcc& p, const Ctx<T>&) const { p.fx = p.fy = p.fz = T(0); p.e = T(0); }
    __device__ __forceinline__ void pair(Acc&, IAcc& p, const Ctx<T>&, bool hit, bool, const RecT<T>&, int, T dx, T dy, T dz, T d2) const {
        const T inv = hit ? fast_rcp<T>(d2) : T(0);
        if (NORM) {
            const T w = inv * s2;
            const T q = w * w * w;
            const T u = xfma(q, q, -q);        // q^2 - q
            p.e += u;
            if (FORCES) {
                const T fs = inv * xfma(q, q, u);   // (2 q^2 - q) / d2
                p.fx = xfma(fs, dx, p.fx); p.fy = xfma(fs, dy, p.fy); p.fz = xfma(fs, dz, p.fz);
            }
        } else {
            const T r6 = inv * inv * inv;
            p.e = xfma(r6, xfma(c12, r6, -c6), p.e);
            if (FORCES) {
                const T fs = inv * r6 * xfma(T(12) * c12, r6, T(-6) * c6);
                p.fx = xfma(fs, dx, p.fx); p.fy = xfma(fs, dy, p.fy); p.fz = xfma(fs, dz, p.fz);
            }
        }
    }
    __device__ void end(Acc& a, IAcc& p, const Ctx<T>& c) const {
        a.e += (double)p.e;
        if (FORCES) { const T k = NORM ? fscale : T(1); fo.store(c, k * p.fx, k * p.fy, k * p.fz); }
    }
    __device__ void finish(Acc& a, ResultBlock* res) const {
        __shared__ double sm[4];
        double e = block_sum(a.e, sm);
        if (threadIdx.x == 0) atomicAdd(&res->f[RB_ENERGY], e * (NORM ? (double)escale : 1.0));
    }
#ifndef __CUDACC_RTC__
    static bool can_normalise(T c6_, T c12_) { return c6_ > T(0) && c12_ > T(0); }   // host
    void set(T c6_, T c12_) {   // host
        c6 = c6_; c12 = c12_; s2 = T(0); escale = T(1); fscale = T(1);
        if (NORM) { s2 = std::cbrt(c12 / c6); escale = c6 * c6 / c12; fscale = T(6) * escale; }
    }
#endif
};

// Coulomb-like k*w_i*w_j/d (test/examples/gravitational_potential.jl:30-34, gravitational_force.jl:38-44)
template <class T, bool FORCES> struct FCoul {
    T k;
    const T* w_i;   // weights gathered into record order of set i / set j: one record-sized slot (w, 0, 0, 0) per record
    const T* w_j;
    ForceOut<T> fo;
    struct Acc { double e; };   // per-tile partial in T, folded into a double per tile (see FLJ)
    struct IAcc { T fx, fy, fz, wi, e; };
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = !FORCES, AUX = true;   // side array staged with the records
    __device__ __forceinline__ const RecT<T>* aux_j() const { return reinterpret_cast<const RecT<T>*>(w_j); }
    __device__ void init(Acc& a) const { a.e = 0.0; }
    __device__ void begin(IAcc& p, const Ctx<T>& c) const { p.fx = p.fy = p.fz = T(0); p.e = T(0); p.wi = c.active ? k * w_i[(size_t)c.ki * 4] : T(0); }
    __device__ __forceinline__ void pair(Acc&, IAcc& p, const Ctx<T>&, bool hit, bool, const RecT<T>&, const RecT<T>& aj, T dx, T dy, T dz, T d2) const {
        const T invd = hit ? fast_rsqrt<T>(d2) : T(0);
        // the side-array slot of a padding record is never initialised (it may hold NaN / Inf left in shared memory by an
        // earlier kernel, and NaN * 0 = NaN): a miss contributes an explicit zero
        const T q = hit ? p.wi * aj.x * invd : T(0);     // k w_i w_j / d
        p.e += q;
        if (FORCES) {
            const T g = q * invd * invd;  // k w_i w_j / d^3
            p.fx = xfma(g, dx, p.fx); p.fy = xfma(g, dy, p.fy); p.fz = xfma(g, dz, p.fz);
        }
    }
    __device__ void end(Acc& a, IAcc& p, const Ctx<T>& c) const { a.e += (double)p.e; if (FORCES) fo.store(c, p.fx, p.fy, p.fz); }
    __device__ void finish(Acc& a, ResultBlock* res) const {
        __shared__ double sm[4];
        double e = block_sum(a.e, sm);
        if (threadIdx.x == 0) atomicAdd(&res->f[RB_ENERGY], e);
    }
};

// histogram storage shared by the two histogram functors: per-thread private bins in shared memory
// (bank = thread, conflict free, no atomics) when nbins <= NB_PRIV_MAX, block-shared atomics otherwise
// PRIV: 1 / 0 = storage kind fixed at compile time (one code path per pair body), -1 = chosen at run time (`priv`)
template <class T, bool SUMS, int PRIV = -1> struct HistBins {
    int nbins, priv;
    int off;                        // byte offset of the bins in dynamic shared memory (after the staging buffers; set by the host)
    unsigned long long* g_counts;   // [nbins] global accumulators
    double* g_sums;                 // [nbins]
    __device__ __forceinline__ unsigned int* cnt() const { extern __shared__ __align__(128) unsigned char dsm_raw[]; return reinterpret_cast<unsigned int*>(dsm_raw + off); }
    __device__ __forceinline__ T* sum() const {
        extern __shared__ __align__(128) unsigned char dsm_raw[];
        const size_t nslots = (size_t)nbins * (is_priv() ? SWEEP_THREADS : 1);
        return reinterpret_cast<T*>(dsm_raw + off + ((nslots * 4 + 15) / 16) * 16);
    }
    __device__ __forceinline__ bool is_priv() const { return (PRIV < 0) ? (priv != 0) : (PRIV != 0); }
    __device__ void init() const {
        const int nslots = nbins * (is_priv() ? SWEEP_THREADS : 1);
        for (int k = threadIdx.x; k < nslots; k += SWEEP_THREADS) { cnt()[k] = 0u; if (SUMS) sum()[k] = T(0); }
        __syncthreads();
    }
    __device__ __forceinline__ void add(int b, T v) const {
        if (is_priv()) {
            const int s = b * SWEEP_THREADS + threadIdx.x;
            cnt()[s] += 1u;
            if (SUMS) sum()[s] += v;
        } else {
            atomicAdd(&cnt()[b], 1u);
            if (SUMS) atomicAdd(&sum()[b], v);
        }
    }
    __device__ void flush() const {
        __syncthreads();
        if (is_priv()) {
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            for (int b = w; b < nbins; b += SWEEP_THREADS / 32) {
                unsigned long long c = 0; double s = 0;
                for (int t = lane; t < SWEEP_THREADS; t += 32) { c += cnt()[b * SWEEP_THREADS + t]; if (SUMS) s += (double)sum()[b * SWEEP_THREADS + t]; }
                c = warp_sum(c);
                if (SUMS) s = warp_sum(s);
                if (lane == 0 && c) { atomicAdd(&g_counts[b], c); if (SUMS) atomicAdd(&g_sums[b], s); }
            }
        } else {
            for (int b = threadIdx.x; b < nbins; b += SWEEP_THREADS)
                if (cnt()[b]) { atomicAdd(&g_counts[b], (unsigned long long)cnt()[b]); if (SUMS) atomicAdd(&g_sums[b], (double)sum()[b]); }
        }
    }
};

// distance histogram: counts[floor(d/width)] += 1 (test/examples/distance_histogram.jl:22-26)
// The bin of a pair is floor(sqrt_rn(d2) / width) in T arithmetic -- monotone in d2, so it is decided EXACTLY by a table
// of d2 thresholds computed on the host with the same correctly rounded operations (thr[b] = smallest d2 whose bin is
// >= b).  The device guesses the bin with the hardware rsqrt seed and corrects the guess against the table: no IEEE
// square root or division in the pair body.  thr == nullptr: the direct form.
template <class T> __device__ __forceinline__ T rsqrt_seed(T x);
template <> __device__ __forceinline__ float rsqrt_seed<float>(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <> __device__ __forceinline__ double rsqrt_seed<double>(double x) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
template <class T, int PRIV> struct FHist {
    T width, inv_width;
    const T* thr;     // nbins + 1 ascending thresholds on d2 (device), or nullptr
    HistBins<T, false, PRIV> hb;
    struct Acc {};
    struct IAcc {};
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = true, AUX = false;
    __device__ void init(Acc&) const { hb.init(); }
    __device__ void begin(IAcc&, const Ctx<T>&) const {}
    __device__ __forceinline__ void pair(Acc&, IAcc&, const Ctx<T>&, bool hit, bool, const RecT<T>&, int, T, T, T, T d2) const {
        if (hit) {
            if (thr != nullptr) {
                int b = (int)fmin(d2 * rsqrt_seed<T>(fmax(d2, T(1e-30))) * inv_width, T(hb.nbins));   // guess; d2 = 0 -> bin 0
                b = max(b, 0);
                while (b > 0 && d2 < __ldg(thr + b)) --b;
                while (b < hb.nbins && d2 >= __ldg(thr + b + 1)) ++b;
                if (b < hb.nbins) hb.add(b, T(0));
            } else {
                const T q = floor(xdiv(xsqrt(d2), width));
                if (q >= T(0) && q < T(hb.nbins)) hb.add((int)q, T(0));
            }
        }
    }
    __device__ void end(Acc&, IAcc&, const Ctx<T>&) const {}
    __device__ void finish(Acc&, ResultBlock*) const { hb.flush(); }
};

// halotools-style mean pairwise velocity (test/examples/pairwise_velocities.jl:17-24)
template <class T> struct Vec4T;
template <> struct Vec4T<float> { typedef float4 type; };
template <> struct Vec4T<double> { typedef double4 type; };
constexpr int VEL_EDGES_INLINE = 33;   // bin edges kept in the kernel parameters (constant bank) up to this many
template <class T, int EDGE_MODE, int PRIV> struct FVel {   // EDGE_MODE 1: <= 8 edges inline, 2: <= VEL_EDGES_INLINE inline, 0: edges in global memory
    const T* v_i;     // velocities gathered into record order, 4 components per record, aligned frame
    const T* v_j;
    const T* rbins;   // nbins+1 ascending edges (device)
    // inline_edges != 0 (nbins + 1 <= VEL_EDGES_INLINE): thr2[e] = the smallest d2 whose correctly rounded square root
    // exceeds edge e, computed on the host -- sqrt_rn is monotone, so (rbins[e] < sqrt(d2)) == (d2 >= thr2[e]) EXACTLY
    // and the bin search needs no square root
    T thr2[VEL_EDGES_INLINE];
    int inline_edges;
    HistBins<T, true, PRIV> hb;
    struct Acc {};
    struct IAcc { T vx, vy, vz; };
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = true, AUX = true;   // side array staged with the records
    __device__ __forceinline__ const RecT<T>* aux_j() const { return reinterpret_cast<const RecT<T>*>(v_j); }
    __device__ void init(Acc&) const { hb.init(); }
    __device__ void begin(IAcc& p, const Ctx<T>& c) const {
        p.vx = p.vy = p.vz = T(0);
        if (c.active) { p.vx = v_i[(size_t)c.ki * 4]; p.vy = v_i[(size_t)c.ki * 4 + 1]; p.vz = v_i[(size_t)c.ki * 4 + 2]; }
    }
    __device__ __forceinline__ void pair(Acc&, IAcc& p, const Ctx<T>&, bool hit, bool, const RecT<T>&, const RecT<T>& aj, T dx, T dy, T dz, T d2) const {
        if (hit) {
            int first = 0;   // searchsortedfirst(rbins, r): number of edges < r = sqrt(d2)
            if (EDGE_MODE == 1) {          // straight-line compares against the constant bank (unused slots hold +inf)
#pragma unroll
                for (int e = 0; e < 8; ++e) first += (d2 >= thr2[e]) ? 1 : 0;
            } else if (EDGE_MODE == 2) {
#pragma unroll 1
                for (int e = 0; e <= hb.nbins; ++e) first += (d2 >= thr2[e]) ? 1 : 0;
            } else {
                const T r = xsqrt(d2);
                for (int e = 0; e <= hb.nbins; ++e) first += (__ldg(rbins + e) < r) ? 1 : 0;
            }
            const int b = first - 1;
            if (b >= 0 && b < hb.nbins) {
                const T ux = p.vx - aj.x, uy = p.vy - aj.y, uz = p.vz - aj.z;
                hb.add(b, ((ux * dx + uy * dy) + uz * dz) * fast_rsqrt<T>(d2));   // sums have tolerance parity; counts are exact
            }
        }
    }
    __device__ void end(Acc&, IAcc&, const Ctx<T>&) const {}
    __device__ void finish(Acc&, ResultBlock*) const { hb.flush(); }
};

// minimum distance (test/examples/nearest_neighbor.jl:9-16): smallest d2, ties broken by (i, j)
struct MinPartial { double d2; long long i, j; long long pad; };
template <class T> struct FMin {
    MinPartial* partial;   // [gridDim.x]
    struct Acc { T d2; long long i, j; };
    struct IAcc {};
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = true, AUX = false;
    __device__ void init(Acc& a) const { a.d2 = CUDART_INF_T<T>(); a.i = 0; a.j = 0; }
    __device__ void begin(IAcc&, const Ctx<T>&) const {}
    __device__ static __forceinline__ bool better(T d2, long long i, long long j, T e2, long long ei, long long ej) {
        return (d2 < e2) || (d2 == e2 && (i < ei || (i == ei && j < ej)));
    }
    __device__ __forceinline__ void pair(Acc& a, IAcc&, const Ctx<T>& c, bool hit, bool, const RecT<T>& rj, int, T, T, T, T d2) const {
        if (hit && d2 <= a.d2) {
            const long long i = (long long)(c.ri.tag & TagT<T>::MASK) + 1, j = (long long)(rj.tag & TagT<T>::MASK) + 1;
            if (better(d2, i, j, a.d2, a.i, a.j)) { a.d2 = d2; a.i = i; a.j = j; }
        }
    }
    __device__ void end(Acc&, IAcc&, const Ctx<T>&) const {}
    __device__ void finish(Acc& a, ResultBlock*) const {
        __shared__ MinPartial sm[SWEEP_THREADS / 32];
        T d2 = a.d2; long long i = a.i, j = a.j;
        for (int o = 16; o > 0; o >>= 1) {
            const T e2 = __shfl_xor_sync(0xffffffffu, d2, o);
            const long long ei = __shfl_xor_sync(0xffffffffu, i, o), ej = __shfl_xor_sync(0xffffffffu, j, o);
            if (better(e2, ei, ej, d2, i, j)) { d2 = e2; i = ei; j = ej; }
        }
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) { sm[w].d2 = (double)d2; sm[w].i = i; sm[w].j = j; }
        __syncthreads();
        if (threadIdx.x == 0) {
            MinPartial b = sm[0];
            for (int k = 1; k < SWEEP_THREADS / 32; ++k)
                if ((sm[k].d2 < b.d2) || (sm[k].d2 == b.d2 && (sm[k].i < b.i || (sm[k].i == b.i && sm[k].j < b.j)))) b = sm[k];
            partial[blockIdx.x] = b;
        }
    }
};

// neighbour-list emission: push_pair! (internals/neighborlist.jl:67-76) as a warp-ballot compaction;
// records are Julia's Tuple{Int,Int,T}: {int64 i; int64 j; T d (+pad)} = 24 bytes
// Emission is staged per warp in shared memory: hits are appended to a warp-private buffer, and a full buffer is
// flushed with ONE global atomicAdd (instead of one per warp step on a single counter) and fully coalesced 8-byte
// stores (instead of 24-byte-strided scattered ones).
constexpr int LIST_STAGE_RECORDS = 160;                               // per warp: 160 records x 24 B = 3840 B
constexpr int LIST_STAGE_BYTES = LIST_STAGE_RECORDS * 24;
template <class T> struct FList {
    unsigned long long* out;        // capacity * 3 words
    unsigned long long capacity;
    T rc2_lo, rc2_hi;               // prevfloat / nextfloat of cutoff^2: the at-cutoff band is counted next to the list
    struct Acc { int cnt; unsigned band; };   // warp-uniform: records waiting in the warp's staging buffer, at-cutoff pairs seen
    struct IAcc {};
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = true, AUX = false;
    __device__ __forceinline__ unsigned long long* stage() const {
        extern __shared__ __align__(128) unsigned char dsm_raw[];
        return reinterpret_cast<unsigned long long*>(dsm_raw + StageTotal<T>::value + (threadIdx.x >> 5) * LIST_STAGE_BYTES);
    }
    __device__ void init(Acc& a) const { a.cnt = 0; a.band = 0u; }
    __device__ void begin(IAcc&, const Ctx<T>&) const {}
    __device__ static __forceinline__ unsigned long long dbits(float d) { return (unsigned long long)__float_as_uint(d); }
    __device__ static __forceinline__ unsigned long long dbits(double d) { return (unsigned long long)__double_as_longlong(d); }
    __device__ __forceinline__ void flush(Acc& a, ResultBlock* res) const {
        const int lane = threadIdx.x & 31;
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&res->c[RC_NLIST], (unsigned long long)a.cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned long long* st = stage();
        const unsigned long long room = (base < capacity) ? (capacity - base) : 0ull;
        const int nrec = (int)((unsigned long long)a.cnt < room ? (unsigned long long)a.cnt : room);   // overflow: counted, not written
        unsigned long long* dst = out + base * 3ull;
        for (int w = lane; w < nrec * 3; w += 32) dst[w] = st[w];
        a.cnt = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void pair(Acc& a, IAcc&, const Ctx<T>& c, bool hit, bool ok, const RecT<T>& rj, int, T, T, T, T d2, ResultBlock* res) const {
        // pairs whose d2 is within 1 ulp of cutoff^2 (north_star: "reported separately"; the reference documents that such
        // pairs may fall on either side, docs/src/neighborlists.md:12)
        a.band += __popc(__ballot_sync(0xffffffffu, ok && d2 >= rc2_lo && d2 <= rc2_hi));
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m == 0u) return;
        if (hit) {
            unsigned long long* r = stage() + (a.cnt + __popc(m & ((1u << c.lane) - 1u))) * 3;
            r[0] = (unsigned long long)(c.ri.tag & TagT<T>::MASK) + 1ull;
            r[1] = (unsigned long long)(rj.tag & TagT<T>::MASK) + 1ull;
            r[2] = dbits(xsqrt(d2));
        }
        a.cnt += __popc(m);
        if (a.cnt > LIST_STAGE_RECORDS - 32) flush(a, res);
    }
    __device__ void end(Acc&, IAcc&, const Ctx<T>&) const {}
    __device__ void finish(Acc& a, ResultBlock* res) const {
        if (a.cnt > 0) flush(a, res);
        if ((threadIdx.x & 31) == 0 && a.band) atomicAdd(&res->c[RC_NBAND], (unsigned long long)a.band);
    }
};
template <class F> struct IsList { static constexpr bool value = false; };
template <class T> struct IsList<FList<T>> { static constexpr bool value = true; };

// ===================================================================================================
// the sweep kernel
// ===================================================================================================
template <class T> __device__ __forceinline__ T huge_coord();
template <> __device__ __forceinline__ float huge_coord<float>() { return 1.0e30f; }
template <> __device__ __forceinline__ double huge_coord<double>() { return 1.0e200; }

// ---- per-warp staging of partner records in shared memory (TMA 1-D bulk copies + mbarrier) ---------------------
constexpr int STAGE_PAD = 32;                            // dummy records after the staged ones: the flat loop needs no bounds logic
template <class T, bool AUX = false> struct StageCap { static constexpr int value = StageBytes<T, AUX>::value / (int)sizeof(RecT<T>) - STAGE_PAD; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion counted on `mbar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ RecT<float> ldrec_s(const RecT<float>* p) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    RecT<float> r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.tag = __float_as_uint(v.w);
    return r;
}
__device__ __forceinline__ RecT<double> ldrec_s(const RecT<double>* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = q[0], b = q[1];
    RecT<double> r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.tag = (uint64_t)__double_as_longlong(b.y);
    return r;
}

// min / max over the particles of a tile (every j-slice holds a copy of the same TILE_I particles, so a whole-warp
// reduction gives the same result): one CREDUX per value in FP32 (sm_100a redux.sync.min/max.f32), shuffles in FP64
__device__ __forceinline__ float tile_min(float v) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float tile_max(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ double tile_min(double v) {
#pragma unroll
    for (int o = 1; o < TILE_I; o <<= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double tile_max(double v) {
#pragma unroll
    for (int o = 1; o < TILE_I; o <<= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Row classes of one tile: skipped; STAGED = bulk-copied into the warp's shared-memory buffer, culled and swept by one
// flat loop.  MODE_HALF rows inside the home reference row additionally have a DIRECT part (the reference cells the tile
// itself touches): bulk-copied uncompacted and swept with per-lane record thresholds.  The own row of the full-shell self sweep is staged too: the tile's own records are blanked in the
// staging buffer and the pairs inside the tile are evaluated from registers (shuffles), which needs no self test in the
// flat loop.
enum { ROW_SKIP = 0, ROW_STAGED = 2 };

// resident CTAs per SM the register allocation aims at (measured, C2 LJ forces: F32 8 CTAs = 64 registers 0.473 ms,
// 10 CTAs = 48 registers 0.515 ms, 6 CTAs 0.488 ms; F64 is limited to 6 CTAs by its 8 KB staging buffers: 80 registers,
// no spills 1.03 ms vs 1.11 ms with 64 registers)
// Functors that stage a side array double the staging memory (F32 4 CTAs, F64 3 CTAs per SM): no reason to squeeze their registers.
template <class T, bool AUX> struct SweepMinBlocks { static constexpr int value = (sizeof(T) == 4) ? (AUX ? 4 : 8) : (AUX ? 3 : 6); };
template <class T, int MODE, class F>
__global__ void __launch_bounds__(SWEEP_THREADS, SweepMinBlocks<T, F::AUX>::value)
k_sweep(const __grid_constant__ SweepArgs<T> a, const __grid_constant__ F f) {
    typedef TagT<T> TG;
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ti = TILE_I, nslice = NSLICE;
    const int lf = a.lf, sub = a.sub, hww = 2 * lf + 1;
    const unsigned smagic = a.sub_magic;
    auto div_sub = [&](int v) { return (sub == 1) ? v : (int)__umulhi((unsigned)v, smagic); };   // 2^32 / 1 does not fit the magic
    const int nrows_st = (a.nz == 1) ? hww : hww * hww;
    constexpr int CAP = StageCap<T, F::AUX>::value;
    constexpr int SBYTES = StageBytes<T, F::AUX>::value;
    RecT<T>* const buf = reinterpret_cast<RecT<T>*>(dsm_raw + warp * SBYTES);
    // functors with a per-record side array (weights, velocities) stage it in a second per-warp buffer, slot for slot
    RecT<T>* const abuf = reinterpret_cast<RecT<T>*>(dsm_raw + StageTotal<T, F::AUX>::value + warp * SBYTES);
    const uint32_t abuf_addr = smem_u32(abuf);
    const uint32_t buf_addr = smem_u32(buf);
    const uint32_t mbar = smem_u32(dsm_raw + (SWEEP_THREADS / 32) * SBYTES + warp * 8);
    uint32_t parity = 0;
    if (lane == 0) { mbar_init(mbar, 1); fence_proxy_async(); }
    __syncwarp();
    if (a.dscal[DS_NTOT] > a.rec_cap_i || a.dscal[DS_SET_STRIDE_DEV + DS_NTOT] > a.rec_cap_j) return;   // overflowed build: results are discarded
    typename F::Acc acc;
    f.init(acc);
    const int ntiles = a.dscal[DS_NTILES];
    // the next tile index is fetched one tile ahead: the atomic's round trip overlaps the current tile's work
    int t_next = 0;
    if (lane == 0) t_next = atomicAdd(&a.dscal[DS_WORK], 1);
    for (;;) {
        const int t = __shfl_sync(0xffffffffu, t_next, 0);
        if (t >= ntiles) break;
        if (lane == 0) t_next = atomicAdd(&a.dscal[DS_WORK], 1);
        const Tile tl = a.tiles[t];
        Ctx<T> c;
        c.lane = lane; c.islot = lane & (ti - 1); c.slice = lane >> LOG2_TILE_I;
        const bool valid = c.islot < tl.cnt;
        c.ki = tl.k0 + (valid ? c.islot : 0);
        c.ri = ldrec(a.rec_i + c.ki);
        const bool real_i = (c.ri.tag & TG::GHOST) == 0;
        c.active = valid && ((c.ri.tag & TG::FOREIGN) == 0) && ((MODE == MODE_HALF) ? ((c.ri.tag & TG::HOME) != 0) : real_i);
        const typename TG::type idx_i = c.ri.tag & TG::MASK;
        typename F::IAcc ia;
        f.begin(ia, c);
        // lanes that hold no particle i are parked far away: every distance test fails, no predicate in the loops
        const T xi = c.active ? c.ri.x : huge_coord<T>(), yi = c.ri.y, zi = c.ri.z;
        // bounding box of the tile's particles i (empty box when no lane is active: everything is culled)
        T blo[3], bhi[3];
        {
            const T inf = CUDART_INF_T<T>();
            blo[0] = tile_min(c.active ? c.ri.x : inf); blo[1] = tile_min(c.active ? c.ri.y : inf); blo[2] = tile_min(c.active ? c.ri.z : inf);
            bhi[0] = tile_max(c.active ? c.ri.x : -inf); bhi[1] = tile_max(c.active ? c.ri.y : -inf); bhi[2] = tile_max(c.active ? c.ri.z : -inf);
        }
        const int iy = tl.yz & 0xffff, iz = tl.yz >> 16, tile_row = iz * a.ny + iy;
        const int cxa = tl.cx & 0xffff, cxb = tl.cx >> 16;
        // MODE_HALF: the reference cell of particle i along the row (the other two follow from the tile's row)
        int rfx_i = 0;
        if (MODE == MODE_HALF) {
            const int* cs = a.cell_start_i + (size_t)tile_row * (a.nx + 1);   // per-cell arrays: row pitch nx + 1 (clm_build.cuh)
            int cx = cxa;
            while (cx < cxb && cs[cx + 1] <= c.ki) ++cx;
            rfx_i = div_sub(cx);
        }
        const int ry_i = div_sub(iy), rz_i = div_sub(iz);

        // the pair body shared by both sweeps
        auto pair_body = [&](const RecT<T>& rj, const RecT<T>& aj, const int jc, bool ok) {
            const T dx = xsub(xi, rj.x), dy_ = xsub(yi, rj.y), dz_ = xsub(zi, rj.z);
            T d2;
            if (F::EXACT_D2) d2 = xadd(xadd(xmul(dx, dx), xmul(dy_, dy_)), xmul(dz_, dz_));
            else d2 = xfma(dz_, dz_, xfma(dy_, dy_, dx * dx));
            const bool hit = ok && (d2 <= a.rc2);
            if constexpr (IsList<F>::value) f.pair(acc, ia, c, hit, ok, rj, jc, dx, dy_, dz_, d2, a.res);
            else if constexpr (F::AUX) f.pair(acc, ia, c, hit, ok, rj, aj, dx, dy_, dz_, d2);   // aj: the partner's side-array slot
            else f.pair(acc, ia, c, hit, ok, rj, jc, dx, dy_, dz_, d2);
        };

        for (int rb = 0; rb < nrows_st; rb += 32) {
            // ---- lane r classifies stencil row r and fetches its record range --------------------------------
            const int r = rb + lane;
            int cls = ROW_SKIP, j0 = 0, j1 = 0, rowbase = 0, own = 0;   // own: the tile's own row
            int dj0 = 0, dj1 = 0;   // MODE_HALF, same reference row: the part of the row that needs the per-lane (record-order) rule
            if (r < nrows_st) {
                const int dz = a.rdz[r], dy = a.rdy[r];
                const int z2 = iz + dz, y2 = iy + dy;
                const int w = a.hw[(dz + lf) * hww + dy + lf];
                bool use = (z2 >= 0 && z2 < a.nz && y2 >= 0 && y2 < a.ny && w >= 0);
                int rel = 1;   // partner row's reference cells vs the home reference cell: < 0 behind, 0 same reference row, > 0 forward
                if (MODE == MODE_HALF && use) {
                    const int rz_j = div_sub(z2), ry_j = div_sub(y2);
                    rel = (rz_j != rz_i) ? (rz_j - rz_i) : (ry_j - ry_i);
                    use = rel >= 0;
                }
                if (use) {
                    own = (dy == 0 && dz == 0) ? 1 : 0;
                    rowbase = (z2 * a.ny + y2) * (a.nx + 1);
                    const int xa = max(cxa - w, 0), xb = min(cxb + w, a.nx - 1);
                    j0 = a.cell_start_j[rowbase + xa];
                    j1 = a.cell_start_j[rowbase + xb + 1];
                    if (MODE == MODE_HALF && rel == 0) {
                        // Partners before the reference cell of the tile's first record are never taken; partners after the
                        // reference cell of its LAST record follow the plain forward rule for every lane and go through the
                        // staged path (cull + flat loop) like the rows of later reference rows.  Only the reference cells
                        // the tile itself touches need the per-lane rule.
                        const int* csj = a.cell_start_j + rowbase;
                        const int tsplit = csj[min((div_sub(cxb) + 1) * sub, a.nx)];
                        dj0 = max(j0, csj[div_sub(cxa) * sub]);
                        dj1 = min(j1, tsplit);
                        j0 = max(j0, tsplit);
                    }
                    if (j1 > j0) cls = ROW_STAGED;
                }
            }
            // ---- direct rows -------------------------------------------------------------------------------------
            if (MODE == MODE_HALF) {
                unsigned mdir = __ballot_sync(0xffffffffu, dj1 > dj0);
                if (mdir) {
                    // All direct segments of the tile are bulk-copied in ONE pass (one expect_tx, one wait) when they fit the
                    // staging buffer together -- they are short: the reference cells the tile touches, a few device rows --
                    // and row by row, piecewise, when they do not.  Record order is kept: the rules below need the record index.
                    const int dlen = max(dj1 - dj0, 0);
                    int dincl = dlen;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, dincl, o); if (lane >= o) dincl += v; }
                    const int doff = dincl - dlen, dtotal = __shfl_sync(0xffffffffu, dincl, 31);
                    const bool batched = dtotal <= CAP;
                    if (batched) {
                        if (lane == 0) { fence_proxy_async(); mbar_expect_tx(mbar, (uint32_t)dtotal * (uint32_t)sizeof(RecT<T>) * (F::AUX ? 2u : 1u)); }
                        __syncwarp();
                        if (dlen > 0) {
                            bulk_g2s(buf_addr + (uint32_t)doff * (uint32_t)sizeof(RecT<T>), a.rec_j + dj0, (uint32_t)dlen * (uint32_t)sizeof(RecT<T>), mbar);
                            if constexpr (F::AUX) bulk_g2s(abuf_addr + (uint32_t)doff * (uint32_t)sizeof(RecT<T>), f.aux_j() + dj0, (uint32_t)dlen * (uint32_t)sizeof(RecT<T>), mbar);
                        }
                        mbar_wait(mbar, parity);
                        parity ^= 1u;
                        __syncwarp();
                    }
                    while (mdir) {
                        const int src = __ffs(mdir) - 1;
                        mdir &= mdir - 1;
                        const int bj0 = __shfl_sync(0xffffffffu, dj0, src), bj1 = __shfl_sync(0xffffffffu, dj1, src);
                        const int brow = __shfl_sync(0xffffffffu, rowbase, src), boff = __shfl_sync(0xffffffffu, doff, src);
                        // Partners in a later reference cell (j >= thrA) follow the forward rule; partners in the SAME reference
                        // cell (thrB <= j < thrA) are taken once: real-real pairs from the earlier record, real-image pairs from
                        // the real particle (the distance is symmetric, so this deviation from the reference's slot order changes
                        // nothing -- and it keeps the rule independent of the record order, which differs between the ranks of a
                        // slab-decomposed system)
                        const int* csj = a.cell_start_j + brow;
                        const int thrA = csj[min((rfx_i + 1) * sub, a.nx)], thrB = csj[rfx_i * sub];
                        for (int p0 = bj0; p0 < bj1; p0 += CAP) {
                            const int pn = min(CAP, bj1 - p0);
                            const RecT<T>* base = buf + boff;
                            if (!batched) {
                                base = buf;
                                if (lane == 0) {
                                    fence_proxy_async();
                                    mbar_expect_tx(mbar, (uint32_t)pn * (uint32_t)sizeof(RecT<T>) * (F::AUX ? 2u : 1u));
                                    bulk_g2s(buf_addr, a.rec_j + p0, (uint32_t)pn * (uint32_t)sizeof(RecT<T>), mbar);
                                    if constexpr (F::AUX) bulk_g2s(abuf_addr, f.aux_j() + p0, (uint32_t)pn * (uint32_t)sizeof(RecT<T>), mbar);
                                }
                                mbar_wait(mbar, parity);
                                parity ^= 1u;
                                __syncwarp();
                            }
                            auto body = [&](const RecT<T>* q, const int jc, const bool inb) {
                                const RecT<T> rj = ldrec_s(q);
                                const bool gi = (c.ri.tag & TG::GHOST) != 0, gj = (rj.tag & TG::GHOST) != 0;
                                const bool ok = inb && ((jc >= thrA) ? !(gi && gj) : (jc >= thrB && !gi && (gj || jc > c.ki)));
                                RecT<T> aj = rj;
                                if constexpr (F::AUX) aj = ldrec_s(abuf + (q - buf));
                                pair_body(rj, aj, jc, ok);
                            };
                            const RecT<T>* q = base + c.slice;
                            int jc = p0 + c.slice;
                            const int nfull = pn / nslice;
                            int s_ = 0;
#pragma unroll 1
                            for (; s_ + 4 <= nfull; s_ += 4) {
                                body(q, jc, true); body(q + nslice, jc + nslice, true); body(q + 2 * nslice, jc + 2 * nslice, true); body(q + 3 * nslice, jc + 3 * nslice, true);
                                q += 4 * nslice; jc += 4 * nslice;
                            }
#pragma unroll 1
                            for (; s_ < nfull; ++s_) { body(q, jc, true); q += nslice; jc += nslice; }
                            if (jc - c.slice < p0 + pn) { const bool inb = jc < p0 + pn; body(inb ? q : base, jc, inb); }
                            if (!batched) __syncwarp();
                        }
                    }
                    __syncwarp();
                }
            }
            // ---- staged rows: prefix of the segment lengths, then chunks of at most CAP records --------------------
            const int len = (cls == ROW_STAGED) ? (j1 - j0) : 0;
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int off = incl - len, total = __shfl_sync(0xffffffffu, incl, 31);
            // full-shell self sweep: where the tile's own records sit in the staged sequence (-1: own row not in this batch)
            int own_lo = -1;
            if (MODE == MODE_ALL && a.self) {
                const unsigned mown = __ballot_sync(0xffffffffu, (own & 1) && cls == ROW_STAGED);
                if (mown) own_lo = __shfl_sync(0xffffffffu, off + (tl.k0 - j0), __ffs(mown) - 1);
            }
            for (int c0 = 0; c0 < total; c0 += CAP) {
                const int cn = min(total - c0, CAP);
                const int lo = max(off, c0), hi = min(off + len, c0 + cn);
                if (lane == 0) { fence_proxy_async(); mbar_expect_tx(mbar, (uint32_t)cn * (uint32_t)sizeof(RecT<T>) * (F::AUX ? 2u : 1u)); }
                __syncwarp();
                if (hi > lo) {
                    bulk_g2s(buf_addr + (uint32_t)(lo - c0) * (uint32_t)sizeof(RecT<T>), a.rec_j + (j0 + (lo - off)), (uint32_t)(hi - lo) * (uint32_t)sizeof(RecT<T>), mbar);
                    if constexpr (F::AUX) bulk_g2s(abuf_addr + (uint32_t)(lo - c0) * (uint32_t)sizeof(RecT<T>), f.aux_j() + (j0 + (lo - off)), (uint32_t)(hi - lo) * (uint32_t)sizeof(RecT<T>), mbar);
                }
                mbar_wait(mbar, parity);
                parity ^= 1u;
                __syncwarp();
                // far-away dummy records: (1) round the chunk up to a whole number of warp-wide cull steps (no bounds logic
                // in the cull loop), (2) blank the tile's own records (full-shell self sweep: those pairs come from registers)
                if (cn + lane < ((cn + 31) & ~31)) strec(buf + cn + lane, -huge_coord<T>(), T(0), T(0), (typename TG::type)0);
                if (MODE == MODE_ALL && a.self) {
                    const int pos = own_lo - c0 + lane;
                    if (own_lo >= 0 && lane < tl.cnt && pos >= 0 && pos < cn) strec(buf + pos, -huge_coord<T>(), T(0), T(0), (typename TG::type)0);
                }
                __syncwarp();
                // cull: keep the staged records within the cutoff of the tile's bounding box, compacted in place.
                // The box distance is evaluated with the same operation order as the pair distance; rounding is
                // monotone, so box distance <= pair distance for every particle i of the tile: no pair is lost.
                int ns = 0;
                {
                    const unsigned lt = (1u << lane) - 1u;
                    const RecT<T>* q = buf + lane;
                    for (int k0 = 0; k0 < cn; k0 += 32, q += 32) {
                        const RecT<T> rq = ldrec_s(q);
                        RecT<T> aq = rq;
                        if constexpr (F::AUX) aq = ldrec_s(abuf + (q - buf));
                        const T ex = fmax(fmax(xsub(blo[0], rq.x), xsub(rq.x, bhi[0])), T(0));
                        const T ey = fmax(fmax(xsub(blo[1], rq.y), xsub(rq.y, bhi[1])), T(0));
                        const T ez = fmax(fmax(xsub(blo[2], rq.z), xsub(rq.z, bhi[2])), T(0));
                        T dd;
                        if (F::EXACT_D2) dd = xadd(xadd(xmul(ex, ex), xmul(ey, ey)), xmul(ez, ez));
                        else dd = xfma(ez, ez, xfma(ey, ey, ex * ex));
                        const bool keep = (dd <= a.rc2);
                        const unsigned m = __ballot_sync(0xffffffffu, keep);   // every lane has read its record: in-place writes are safe
                        if (keep) {
                            const int pos = ns + __popc(m & lt);
                            strec(buf + pos, rq.x, rq.y, rq.z, rq.tag);
                            if constexpr (F::AUX) strec(abuf + pos, aq.x, aq.y, aq.z, aq.tag);
                        }
                        ns += __popc(m);
                    }
                }
                // dummy far-away records round the survivors up to a whole warp step (one record per j-slice)
                const int nsteps = (ns + nslice - 1) / nslice;
                if (ns + lane < nsteps * nslice) strec(buf + ns + lane, -huge_coord<T>(), T(0), T(0), (typename TG::type)0);
                __syncwarp();
                const RecT<T>* p = buf + c.slice;
                auto sbody = [&](const RecT<T>* q) {
                    const RecT<T> rj = ldrec_s(q);
                    RecT<T> aj = rj;
                    if constexpr (F::AUX) aj = ldrec_s(abuf + (q - buf));
                    bool ok = true;
                    if (MODE == MODE_HALF) ok = (((c.ri.tag & rj.tag) & TG::GHOST) == 0);
                    else if (MODE == MODE_TRI) ok = (idx_i < (rj.tag & TG::MASK));
                    pair_body(rj, aj, 0, ok);
                };
                int g = 0;
#pragma unroll 1
                for (; g + 8 <= nsteps; g += 8) {
                    sbody(p); sbody(p + nslice); sbody(p + 2 * nslice); sbody(p + 3 * nslice);
                    sbody(p + 4 * nslice); sbody(p + 5 * nslice); sbody(p + 6 * nslice); sbody(p + 7 * nslice);
                    p += 8 * nslice;
                }
#pragma unroll 1
                for (; g < nsteps; ++g) { sbody(p); p += nslice; }
                __syncwarp();
            }
        }
        // full-shell self sweep: the pairs INSIDE the tile, partner records taken from the registers of the lanes that hold them
        if (MODE == MODE_ALL && a.self) {
#pragma unroll
            for (int s_ = 0; s_ < TILE_I / NSLICE; ++s_) {
                const int js = c.slice + NSLICE * s_;
                RecT<T> rj;
                rj.x = __shfl_sync(0xffffffffu, c.ri.x, js); rj.y = __shfl_sync(0xffffffffu, c.ri.y, js); rj.z = __shfl_sync(0xffffffffu, c.ri.z, js);
                rj.tag = __shfl_sync(0xffffffffu, c.ri.tag, js);
                const bool ok = (js != c.islot) && (js < tl.cnt);
                RecT<T> aj = rj;
                if constexpr (F::AUX) aj = ldrec(f.aux_j() + (tl.k0 + (js < tl.cnt ? js : 0)));
                pair_body(rj, aj, tl.k0 + js, ok);
            }
        }
        f.end(acc, ia, c);
    }
    f.finish(acc, a.res);
}

// ---- small epilogue kernels ---------------------------------------------------------------------------
struct MinResult { double d2; long long i, j; long long pad; };
static __global__ void k_min_final(const MinPartial* __restrict__ p, int n, MinResult* __restrict__ out) {
    __shared__ MinPartial sm[256];
    auto better = [](const MinPartial& q, const MinPartial& b) { return (q.d2 < b.d2) || (q.d2 == b.d2 && (q.i < b.i || (q.i == b.i && q.j < b.j))); };
    MinPartial b; b.d2 = CUDART_INF_T<double>(); b.i = 0; b.j = 0; b.pad = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) { const MinPartial q = p[k]; if (q.i != 0 && better(q, b)) b = q; }
    sm[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < blockDim.x; ++k) { const MinPartial q = sm[k]; if (q.i != 0 && (b.i == 0 || better(q, b))) b = q; }
        out->d2 = b.d2; out->i = b.i; out->j = b.j; out->pad = 0;
    }
}
// device-resident (i, j, d) output of the minimum-distance map; reset = false keeps an existing smaller d
template <class T> __global__ void k_min_store(const MinResult* __restrict__ r, long long* i_out, long long* j_out, T* d_out, int accumulate) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const T d = (r->i != 0) ? xsqrt((T)r->d2) : CUDART_INF_T<T>();
    if (accumulate && !(d < *d_out)) return;
    *i_out = r->i; *j_out = r->j; *d_out = d;
}

// out[k] = (accumulate ? out[k] : 0) + scale * src[k]   (device-resident outputs, CLM_OUT_DEVICE)
template <class T> __global__ void k_store_real(T* __restrict__ out, const double* __restrict__ src, int n, double scale, int accumulate) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (T)((accumulate ? (double)out[k] : 0.0) + scale * src[k]);
}
static __global__ void k_store_i64(long long* __restrict__ out, const unsigned long long* __restrict__ src, int n, int accumulate, int shift) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (accumulate ? out[k] : 0ll) + (long long)(src[k] >> shift);
}

}  // namespace clm

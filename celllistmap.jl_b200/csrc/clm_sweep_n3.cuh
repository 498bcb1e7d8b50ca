// Force sweep with Newton's third law: every in-cutoff pair of a self-set system is evaluated ONCE and updates both
// particles (the reference's own scheme: f[i] += df; f[j] -= df, docs/src/ParticleSystem/examples.md:41-47, pair
// enumeration of src/internals/self.jl:143-184 / vicinal_cells.jl:21-65).  Replaces the full-shell k_sweep<MODE_ALL>
// for the self-set force maps of the catalogue (LJ, Coulomb).
//
// Lane layout (the transpose of k_sweep's): the 32 lanes of a warp hold 32 DIFFERENT partners j -- coordinates and the
// partner's force accumulator live in registers -- and the TILE_I particles i of the tile are broadcast from shared
// memory, one per warp step.  f_j needs no cross-lane reduction at all: it is flushed once per (tile, partner) with one
// 128-bit vector reduction (red.global.add.v4.f32, REDG.E.ADD.F32x4 in SASS); f_i is accumulated in TILE_I x 3
// registers per lane and reduce-scattered over the warp once per tile.
//
// The reductions go to a RECORD-ordered accumulator (one 4 x T row per record slot, facc): consecutive lanes hold
// consecutive records of a row, so a warp's flush touches a few 128-byte lines even when the caller's particle
// numbering is random (tools/red_microbench.cu: 3.7e11 coalesced vs 1.9e11 random lane-REDs/s with 1 M slots, 3.0e11
// vs 6.1e10 with 8 M).  For that the build writes a second record array (rec_n3) whose 4th word is the slot of the
// particle's REAL record (an image points at its original) | GHOST | HOME | cell parity: images add straight into their
// original's row and k_force_finish gathers the rows into the caller's particle order through slot_of[] (scale, inverse
// rotation, reset/accumulate).  (Measured alternative: no twin array, the row of a partner gathered from slot_of[index]
// by every lane of every chunk -- the uncoalesced 4-byte gathers cost 0.045 ms of the 0.465 ms sweep; the twin array
// costs 0.006 ms of the build.)
//
// Exactly-once rules, identical pair sets to k_sweep<MODE_HALF / MODE_TRI> (which are bit-exact against the oracle):
//   MODE_HALF  forward reference rows / cells: every record, except image-image pairs; partners in the reference cells
//              the tile itself touches ("direct" part) carry their reference cell index as a key: later cell -> forward
//              rule, same cell -> real i only, image partner or later record slot.
//   MODE_TRI   full stencil, i real, index_i < index_j: the reference's rule, so that of the two symmetric image pairs of
//              a pair the same one is evaluated (their Float32 coordinates round differently).  For triclinic cells the
//              4th word of rec_n3 holds the particle INDEX and the accumulator rows are in particle order.
#pragma once
#include "clm_sweep.cuh"

namespace clm {

static_assert(TILE_I == 8, "the reduce-scatter of the i-side accumulators is written for 8 particles per tile");
// rec_n3 4th word: slot of the particle's REAL record (an image points at its original) | flags.  PAR = parity of the
// record's reference cell along the row (the key of the record-order rules follows from it, see k_sweep_n3)
constexpr uint32_t N3_GHOST = 0x80000000u, N3_HOME = 0x40000000u, N3_SLOT = 0x1fffffffu;
constexpr int N3_PAR_SHIFT = 29;

// per-warp shared memory: [mbarrier | i-side keys | i-side positions | staging buffer (| side-array staging buffer)].
// With 80-96 registers per thread 5-6 CTAs are resident per SM whatever the buffer size up to 8 KB, so the buffer is
// sized for the whole partner sequence of a typical tile in one pass
#ifndef CLM_N3_STAGE_BYTES_F32
#define CLM_N3_STAGE_BYTES_F32 8192
#endif
#ifndef CLM_N3_STAGE_BYTES_F64
#define CLM_N3_STAGE_BYTES_F64 8192
#endif
// (LEAN = a functor without force outputs -- energy only.  Measured on the C2 workload, Float32, tools/tune_n3_energy.sh,
//  profiles/r2_tune_n3_energy.txt: the kernel wants ~116 registers spill-free even without the accumulators; 5 resident CTAs
//  at 96 registers with the 8 KB buffer are fastest -- 0.329 ms against 0.361 ms of k_sweep<MODE_HALF>; 8 CTAs at 64 registers
//  spill and take 0.409 ms.  Float64 stays on k_sweep<MODE_HALF>: 0.66 ms here against 0.574 ms.)
#ifndef CLM_N3E_STAGE_BYTES_F32
#define CLM_N3E_STAGE_BYTES_F32 8192
#endif
#ifndef CLM_N3E_STAGE_BYTES_F64
#define CLM_N3E_STAGE_BYTES_F64 8192
#endif
template <class T, bool AUX, bool LEAN = false> struct N3Cap {
    static constexpr int SB = AUX ? 6144 : (LEAN ? ((sizeof(T) == 4) ? CLM_N3E_STAGE_BYTES_F32 : CLM_N3E_STAGE_BYTES_F64)
                                                 : ((sizeof(T) == 4) ? CLM_N3_STAGE_BYTES_F32 : CLM_N3_STAGE_BYTES_F64));
    static constexpr int REC = (int)sizeof(RecT<T>);
    static constexpr int SLOTS = SB / REC;
    static constexpr int CAPP = SLOTS - 64;      // staged per pass: the carried partial chunk (< 32) and the padding (< 32) share the buffer
    static constexpr int IKEY_OFF = 128, IPOS_OFF = 256, BUF_OFF = 512, ABUF_OFF = BUF_OFF + SB;
    static constexpr int WSTRIDE = BUF_OFF + SB * (AUX ? 2 : 1);
};
template <class T, bool AUX, bool LEAN = false> struct N3Smem { static constexpr int value = (SWEEP_THREADS / 32) * N3Cap<T, AUX, LEAN>::WSTRIDE; };

__device__ __forceinline__ void red_add3(float* p, float x, float y, float z) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}
__device__ __forceinline__ void red_add3(double* p, double x, double y, double z) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(y) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 2), "d"(z) : "memory");
}
// accumulator row of a record slot: facc + slot * 4 elements as ONE 32 x 32 -> 64-bit multiply-add (the compiler's own
// 64-bit address arithmetic for a masked 29-bit slot is seven instructions per flush)
template <class T> __device__ __forceinline__ T* facc_row(T* facc, uint32_t slot) {
    unsigned long long p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(slot), "r"((uint32_t)(4 * sizeof(T))), "l"(facc));
    return reinterpret_cast<T*>(p);
}
// true unless all three are +0 (a lane that saw no pair holds exact +0s; -0 only costs a redundant reduction)
__device__ __forceinline__ bool any_nonzero(float x, float y, float z) { return (__float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z)) != 0u; }
__device__ __forceinline__ bool any_nonzero(double x, double y, double z) { return x != 0.0 || y != 0.0 || z != 0.0; }
// slot word of a staged rec_n3 record
__device__ __forceinline__ uint32_t slotword(const RecT<float>& r) { return r.tag; }
__device__ __forceinline__ uint32_t slotword(const RecT<double>& r) { return (uint32_t)r.tag; }

// ---- pair-force functors: fs(hit, d2, e, wi, wj) returns the scalar s with F_i += s (x_i - x_j), F_j -= s (x_i - x_j);
//      e accumulates the (unscaled) pair energy of the tile ----
template <class T, bool NORM, bool ENERGY, bool FORCES_ = true> struct N3LJ {
    T c6, c12, s2, escale, fscale;      // see FLJ (clm_sweep.cuh): q = (c12/c6)/d2^3 form when both constants are positive
    static constexpr bool AUX = false;
    static constexpr bool FORCES = FORCES_;   // false: energy only (the exactly-once energy map rides the same lean sweep)
    __device__ __forceinline__ const RecT<T>* aux_j() const { return nullptr; }
    __device__ __forceinline__ T wi(int) const { return T(0); }
    __device__ __forceinline__ T fs(bool hit, T d2, T& e, T, T) const {
        const T inv = hit ? fast_rcp<T>(d2) : T(0);
        if (NORM) {
            const T w = inv * s2;
            const T q = w * w * w;
            const T u = xfma(q, q, -q);
            if (ENERGY) e += u;
            return inv * xfma(q, q, u);
        } else {
            const T r6 = inv * inv * inv;
            if (ENERGY) e = xfma(r6, xfma(c12, r6, -c6), e);
            if (!FORCES) return T(0);
            return inv * r6 * xfma(T(12) * c12, r6, T(-6) * c6);
        }
    }
    __device__ __forceinline__ double energy_scale() const { return NORM ? (double)escale : 1.0; }
#ifndef __CUDACC_RTC__
    void set(T c6_, T c12_) {
        c6 = c6_; c12 = c12_; s2 = T(0); escale = T(1); fscale = T(1);
        if (NORM) { s2 = std::cbrt(c12 / c6); escale = c6 * c6 / c12; fscale = T(6) * escale; }
    }
#endif
};
template <class T, bool ENERGY> struct N3Coul {
    T k;
    const T* w_rec;   // weights gathered into record order, one record-sized slot (w, 0, 0, 0) per record (Engine::gather_aux)
    T fscale;         // 1
    static constexpr bool AUX = true;
    static constexpr bool FORCES = true;
    __device__ __forceinline__ const RecT<T>* aux_j() const { return reinterpret_cast<const RecT<T>*>(w_rec); }
    __device__ __forceinline__ T wi(int ki) const { return k * w_rec[(size_t)ki * 4]; }
    __device__ __forceinline__ T fs(bool hit, T d2, T& e, T wi_, T wj) const {
        const T invd = hit ? fast_rsqrt<T>(d2) : T(0);
        const T q = hit ? wi_ * wj * invd : T(0);     // a padding slot's side-array entry is never trusted (NaN * 0)
        if (ENERGY) e += q;
        return q * invd * invd;
    }
    __device__ __forceinline__ double energy_scale() const { return 1.0; }
};

template <class T> struct Vec4S;
template <> struct __align__(16) Vec4S<float> { float x, y, z, w; };
template <> struct __align__(32) Vec4S<double> { double x, y, z, w; };
__device__ __forceinline__ Vec4S<float> ldvec4_s(const Vec4S<float>* p) { const float4 v = *reinterpret_cast<const float4*>(p); Vec4S<float> r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
__device__ __forceinline__ Vec4S<double> ldvec4_s(const Vec4S<double>* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = q[0], b = q[1];
    Vec4S<double> r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y; return r;
}

enum { N3_PLAIN = 0, N3_KEYED = 1, N3_GENERAL = 2, N3_TRI = 3 };

#ifndef CLM_N3_MINB_F32
#define CLM_N3_MINB_F32 5
#endif
#ifndef CLM_N3_HALVES_F64
#define CLM_N3_HALVES_F64 1   // 2: measured slower (see the comment at the i-side accumulators)
#endif
#ifndef CLM_N3_MINB_F64
#define CLM_N3_MINB_F64 ((CLM_N3_HALVES_F64 == 2) ? 5 : 4)
#endif
template <class T> struct N3Halves { static constexpr int value = (sizeof(T) == 8) ? CLM_N3_HALVES_F64 : 1; };
#ifndef CLM_N3E_MINB_F32
#define CLM_N3E_MINB_F32 5
#endif
#ifndef CLM_N3E_MINB_F64
#define CLM_N3E_MINB_F64 3
#endif
template <class T, bool AUX, bool LEAN = false> struct N3MinBlocks {
    static constexpr int value = (LEAN && !AUX) ? ((sizeof(T) == 4) ? CLM_N3E_MINB_F32 : CLM_N3E_MINB_F64)
                                                : ((sizeof(T) == 4) ? (AUX ? 4 : CLM_N3_MINB_F32) : (AUX ? 3 : CLM_N3_MINB_F64));
};

// Per tile (TILE_I consecutive records of one row, particles i):
//   1. lane r classifies stencil row r and fetches its record range; rows of the tile's own REFERENCE row split into
//      the KEYED part D (the reference cells rbase, rbase + 1 the tile's particles live in: the record-order rules
//      apply) and the forward part F (plain forward rule).  The partner sequence of the tile is S = [D rows | F rows].
//   2. S is bulk-copied (TMA 1-D copies, one or two per row) behind the partners that wait at buf[0 .. nl), culled
//      against the bounding box of the tile and compacted in place; nk = number of keyed partners at the front.
//   3. whole chunks of 32 partners are swept (one partner per lane, TILE_I warp steps per chunk); the partial chunk at
//      the end is carried to the next pass; the last pass pads it and sweeps it too.
// The key of a keyed partner (its reference cell along the row) is rbase + (parity bit of its record ^ parity of
// rbase): no key array.  A tile that spans more than two reference cells (sparse rows only) is processed in several
// trips of the rbase loop, each with the particles i of two reference cells.
#ifdef CLM_N3_MAXNREG   // tuning builds: an explicit register cap instead of the occupancy target (tools/tune_n3_cta.sh)
#define CLM_N3_BOUNDS __maxnreg__(CLM_N3_MAXNREG)
#else
#define CLM_N3_BOUNDS __launch_bounds__(SWEEP_THREADS, N3MinBlocks<T, F::AUX, !F::FORCES>::value)
#endif
template <class T, int MODE, class F>
__global__ void CLM_N3_BOUNDS
k_sweep_n3(const __grid_constant__ SweepArgs<T> a, const __grid_constant__ F f, T* __restrict__ facc) {
    typedef TagT<T> TG;
    typedef typename TG::type tag_t;
    typedef N3Cap<T, F::AUX, !F::FORCES> CP;
    constexpr bool FORCES = F::FORCES;
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    const int lane = threadIdx.x & 31;
    constexpr int NH = N3Halves<T>::value, IH = TILE_I / NH;
    // one base address per warp; everything else sits at a compile-time offset from it
    unsigned char* const wb = dsm_raw + (threadIdx.x >> 5) * CP::WSTRIDE;
    // [TILE_I] i-side keys.  MODE_HALF: x = (reference cell - rbase) << 30 | own slot (an image particle: | 0x3fffffff, so that
    // no partner of its own cell passes), y = image flag; MODE_TRI: x = particle index.  A partner's key is built the same way
    // (cell 2 = beyond the keyed part; image partner: slot bits 0x3fffffff) and a pair counts iff key_j > key_i (unsigned).
    int2* const ikey = reinterpret_cast<int2*>(wb + CP::IKEY_OFF);
    Vec4S<T>* const ipos = reinterpret_cast<Vec4S<T>*>(wb + CP::IPOS_OFF);  // [TILE_I]: x, y, z, functor weight
    RecT<T>* const buf = reinterpret_cast<RecT<T>*>(wb + CP::BUF_OFF);
    RecT<T>* const abuf = reinterpret_cast<RecT<T>*>(wb + CP::ABUF_OFF);
    constexpr uint32_t REC = (uint32_t)sizeof(RecT<T>);
    const uint32_t mbar = smem_u32(wb), buf_addr = smem_u32(buf), abuf_addr = smem_u32(abuf);
    uint32_t parity = 0;
    if (lane == 0) { mbar_init(mbar, 1); fence_proxy_async(); }
    __syncwarp();
    if (a.dscal[DS_NTOT] > a.rec_cap_i) return;   // overflowed build: the host repeats build + map
    const int lf = a.lf, sub = a.sub, hww = 2 * lf + 1;
    const unsigned smagic = a.sub_magic;
    auto div_sub = [&](int v) { return (sub == 1) ? v : (int)__umulhi((unsigned)v, smagic); };
    const int nrows_st = (a.nz == 1) ? hww : hww * hww;
    double e_acc = 0.0;
    const int ntiles = a.dscal[DS_NTILES];
    const T rc2 = a.rc2;
    const unsigned lt = (1u << lane) - 1u;
    const RecT<T>* const rec = a.rec_j;           // slot-tagged records (rec_n3)
    // stencil row of this lane (row offset, half width along the row), looked up ONCE when the stencil has at most 32 rows
    // (lcell * sub <= 2): the two dependent indexed constant loads leave the per-tile path
    int rowpack = 0;
    const bool rows_fit = nrows_st <= 32;
    if (rows_fit && lane < nrows_st) {
        const int dz = a.rdz[lane], dy = a.rdy[lane];
        rowpack = (a.hw[(dz + lf) * hww + dy + lf] & 0xff) | ((dy & 0xff) << 8) | ((dz & 0xff) << 16);
    }
    int t_next = 0;
    if (lane == 0) t_next = atomicAdd(&a.dscal[DS_WORK], 1);
    for (;;) {
        const int t = __shfl_sync(0xffffffffu, t_next, 0);
        if (t >= ntiles) break;
        if (lane == 0) t_next = atomicAdd(&a.dscal[DS_WORK], 1);
        const Tile tl = a.tiles[t];
        const int islot = lane & (TILE_I - 1);
        const bool valid = islot < tl.cnt;
        const int ki = tl.k0 + (valid ? islot : 0);
        const RecT<T> ri = ldrec(rec + ki);
        const uint32_t swi = slotword(ri);
        const bool ghost_i = (swi & N3_GHOST) != 0u;
        const bool active0 = valid && ((MODE == MODE_HALF) ? ((swi & N3_HOME) != 0u) : !ghost_i);
        const int iy = tl.yz & 0xffff, iz = tl.yz >> 16;
        const int cxa = tl.cx & 0xffff, cxb = tl.cx >> 16;
        int rfx_i = 0, rfa = 0, rfb = 0;
        if (MODE == MODE_HALF) {
            const int* cs = a.cell_start_i + (size_t)(iz * a.ny + iy) * (a.nx + 1);
            // device cell of record ki: cxa + the number of cells of the tile that end at or before ki (cell ends are monotone).
            // The ends are loaded by the lanes in parallel and compared through shuffles: one memory round trip instead of a
            // chain of dependent loads
            int cx = cxa;
            const int ncs = min(cxb - cxa, 32);
            const int vend = (lane < ncs) ? cs[cxa + 1 + lane] : 0x7fffffff;
            for (int k = 0; k < ncs; ++k) cx += (__shfl_sync(0xffffffffu, vend, k) <= ki) ? 1 : 0;
            if (cxb - cxa > 32) { while (cx >= cxa + 32 && cx < cxb && cs[cx + 1] <= ki) ++cx; }
            rfx_i = div_sub(cx); rfa = div_sub(cxa); rfb = div_sub(cxb);
        }
        const int ry_i = div_sub(iy), rz_i = div_sub(iz);
        // i-side accumulators: NH == 1: all TILE_I particles of the tile, reduced once per tile.  NH == 2 (tuning builds,
        // -DCLM_N3_HALVES_F64=2): the chunks of a pass are swept twice, once per HALF of the tile's particles, with IH x 3
        // accumulators that are zeroed before and reduced after every half-sweep.  Built for Float64, where 8 x 3 doubles live
        // through the pair loop cost 128 registers + spills (16 warps / SM, 5.8e8 warp instructions), and measured SLOWER:
        // 0.984 ms (128 registers, no spills) / 1.047 ms (96 registers, 20 warps) against 0.920 ms of NH == 1 and 0.906 ms of the
        // full shell -- every partner is loaded and its force flushed (three scalar f64 reductions) twice
        // (tools/tune_n3_f64.sh, profiles/r2_tune_n3_f64.txt)
        T fi[FORCES ? IH : 1][3];
#pragma unroll
        for (int i = 0; i < (FORCES ? IH : 1); ++i) fi[i][0] = fi[i][1] = fi[i][2] = T(0);
        T e_tile = T(0);
        // f_i: reduce-scatter of the IH x 3 per-lane partial sums over the warp: after log2(IH) halving steps the 32 / IH lanes
        // of group g hold the partial sums of particle i0 + g, butterfly steps finish them; the accumulators are zeroed
        auto reduce_fi = [&](const int i0) {
            if constexpr (!FORCES) return;
            T v[3 * IH];
#pragma unroll
            for (int i = 0; i < IH; ++i) { const int ia = FORCES ? i : 0; v[3 * i] = fi[ia][0]; v[3 * i + 1] = fi[ia][1]; v[3 * i + 2] = fi[ia][2]; fi[ia][0] = fi[ia][1] = fi[ia][2] = T(0); }
#pragma unroll
            for (int half = 3 * IH / 2, o = 16; half >= 3; half >>= 1, o >>= 1) {
                const bool upper = (lane & o) != 0;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                    const T send = upper ? v[k] : v[k + half];
                    const T keep = upper ? v[k + half] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            constexpr int GL = 32 / IH;       // lanes per particle
#pragma unroll
            for (int o = GL / 2; o >= 1; o >>= 1) {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], o); v[1] += __shfl_xor_sync(0xffffffffu, v[1], o); v[2] += __shfl_xor_sync(0xffffffffu, v[2], o);
            }
            const int g = i0 + lane / GL;
            const int slot_g = (int)(__shfl_sync(0xffffffffu, swi, g) & N3_SLOT);
            const bool act_g = __shfl_sync(0xffffffffu, active0 ? 1 : 0, g) != 0;
            if ((lane & (GL - 1)) == 0 && act_g && (v[0] != T(0) || v[1] != T(0) || v[2] != T(0))) red_add3(facc_row(facc, (uint32_t)slot_g), v[0], v[1], v[2]);
        };

#pragma unroll 1
        for (int rbase = rfa; rbase <= rfb; rbase += 2) {
            const bool active = active0 && (MODE != MODE_HALF || (unsigned)(rfx_i - rbase) < 2u);
            T blo[3], bhi[3];
            {
                const T inf = CUDART_INF_T<T>();
                blo[0] = tile_min(active ? ri.x : inf); blo[1] = tile_min(active ? ri.y : inf); blo[2] = tile_min(active ? ri.z : inf);
                bhi[0] = tile_max(active ? ri.x : -inf); bhi[1] = tile_max(active ? ri.y : -inf); bhi[2] = tile_max(active ? ri.z : -inf);
            }
            const bool any_ghost_i = (MODE == MODE_HALF) && (__ballot_sync(0xffffffffu, active && ghost_i) != 0u);
            __syncwarp();   // every lane is done with the i-side data of the previous trip / tile
            if (lane < TILE_I) {
                Vec4S<T> v;
                v.x = active ? ri.x : huge_coord<T>(); v.y = ri.y; v.z = ri.z; v.w = (F::AUX && active) ? f.wi(ki) : T(0);
                ipos[lane] = v;
                if (MODE == MODE_TRI) ikey[lane] = make_int2((int)(swi & N3_SLOT), 0);
                else ikey[lane] = make_int2((int)(((unsigned)(rfx_i - rbase) << 30) | (ghost_i ? 0x3fffffffu : (unsigned)ki)), ghost_i ? 1 : 0);
            }
            __syncwarp();
            // keyed part along the row (MODE_HALF): device cells [xD0, xD1) = reference cells rbase .. min(rbase + 1, rfb)
            const int xD0 = rbase * sub, xD1 = min((min(rbase + 1, rfb) + 1) * sub, a.nx);
            const uint32_t rpar = (uint32_t)(rbase & 1);
            int nl = 0, nk = 0;   // partners waiting at buf[0 .. nl); the first nk of them are keyed

            // ---- one chunk = 32 partners, one per lane; one warp step per particle i of the tile ------------------
            auto chunk = [&](auto kind_tag, auto half_tag, const int s) {
                constexpr int KIND = decltype(kind_tag)::value;
                constexpr int I0 = decltype(half_tag)::value * IH;
                const RecT<T> rj = ldrec_s(buf + s);
                T wj = T(0);
                if constexpr (F::AUX) wj = ldrec_s(abuf + s).x;
                const uint32_t sw = slotword(rj);
                const bool gj = (sw & N3_GHOST) != 0u;
                const int slot_j = (int)(sw & N3_SLOT);
                // key of the partner: see ikey
                unsigned KJ = 0u;
                if (MODE == MODE_TRI) KJ = (unsigned)slot_j;
                else if (KIND == N3_KEYED || KIND == N3_GENERAL)
                    KJ = ((s < nk) ? ((((sw >> N3_PAR_SHIFT) & 1u) ^ rpar) << 30) : 0x80000000u) | (gj ? 0x3fffffffu : (unsigned)slot_j);
                T fjx = T(0), fjy = T(0), fjz = T(0);
#pragma unroll
                for (int ii = 0; ii < IH; ++ii) {
                    const int i = I0 + ii;
                    const Vec4S<T> pi = ldvec4_s(ipos + i);
                    const T dx = pi.x - rj.x, dy = pi.y - rj.y, dz = pi.z - rj.z;
                    const T d2 = xfma(dz, dz, xfma(dy, dy, dx * dx));
                    bool ok = true;
                    if (KIND == N3_KEYED || KIND == N3_TRI) ok = KJ > (unsigned)ikey[i].x;
                    else if (KIND == N3_GENERAL) { const int2 ik = ikey[i]; ok = (KJ > (unsigned)ik.x) && !(ik.y && gj); }
                    const bool hit = ok && (d2 <= rc2);
                    const T sc = f.fs(hit, d2, e_tile, pi.w, wj);
                    if constexpr (FORCES) {
                        fi[ii][0] = xfma(sc, dx, fi[ii][0]); fi[ii][1] = xfma(sc, dy, fi[ii][1]); fi[ii][2] = xfma(sc, dz, fi[ii][2]);
                        fjx = xfma(-sc, dx, fjx); fjy = xfma(-sc, dy, fjy); fjz = xfma(-sc, dz, fjz);
                    }
                }
                if constexpr (FORCES) { if (any_nonzero(fjx, fjy, fjz)) red_add3(facc_row(facc, (uint32_t)slot_j), fjx, fjy, fjz); }
            };
            // one cull step: 32 staged records, survivors compacted in place behind buf[ns)
            auto cull_step = [&](const RecT<T>* q, int& ns) -> unsigned {
                const RecT<T> rq = ldrec_s(q);
                RecT<T> aq = rq;
                if constexpr (F::AUX) aq = ldrec_s(abuf + (q - buf));
                const T ex = fmax(fmax(blo[0] - rq.x, rq.x - bhi[0]), T(0));
                const T ey = fmax(fmax(blo[1] - rq.y, rq.y - bhi[1]), T(0));
                const T ez = fmax(fmax(blo[2] - rq.z, rq.z - bhi[2]), T(0));
                const T dd = xfma(ez, ez, xfma(ey, ey, ex * ex));
                const bool keep = (dd <= rc2);
                const unsigned m = __ballot_sync(0xffffffffu, keep);   // every lane has read its slot: in-place writes are safe
                if (keep) {
                    const int pos = ns + __popc(m & lt);
                    strec(buf + pos, rq.x, rq.y, rq.z, rq.tag);
                    if constexpr (F::AUX) strec(abuf + pos, aq.x, aq.y, aq.z, aq.tag);
                }
                return m;
            };

#pragma unroll 1
            for (int rb = 0; rb < nrows_st; rb += 32) {
                // ---- lane r classifies stencil row r (same rules as k_sweep) ---------------------------------------
                const int r = rb + lane;
                int j0 = 0, j1 = 0, d0 = 0, d1 = 0;
                if (r < nrows_st) {
                    int dz, dy, w;
                    if (rows_fit) { w = (int)(signed char)(rowpack & 0xff); dy = (int)(signed char)((rowpack >> 8) & 0xff); dz = (int)(signed char)((rowpack >> 16) & 0xff); }
                    else { dz = a.rdz[r]; dy = a.rdy[r]; w = a.hw[(dz + lf) * hww + dy + lf]; }
                    const int z2 = iz + dz, y2 = iy + dy;
                    bool use = ((unsigned)z2 < (unsigned)a.nz) && ((unsigned)y2 < (unsigned)a.ny) && (w >= 0);
                    int rel = 1;
                    if (MODE == MODE_HALF && use) {
                        const int rz_j = div_sub(z2), ry_j = div_sub(y2);
                        rel = (rz_j != rz_i) ? (rz_j - rz_i) : (ry_j - ry_i);
                        use = rel >= 0;
                    }
                    if (use) {
                        const int* csj = a.cell_start_j + (size_t)(z2 * a.ny + y2) * (a.nx + 1);
                        j0 = csj[max(cxa - w, 0)];
                        j1 = csj[min(cxb + w, a.nx - 1) + 1];
                        if (MODE == MODE_HALF && rel == 0) {
                            const int sD1 = csj[xD1];
                            d0 = max(j0, csj[xD0]);
                            d1 = min(j1, sD1);
                            j0 = max(j0, sD1);
                        }
                    }
                }
                const int dlen = (MODE == MODE_HALF) ? max(d1 - d0, 0) : 0, flen = max(j1 - j0, 0);
                int dincl = dlen, fincl = flen;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, fincl, o);
                    if (lane >= o) fincl += v;
                    if (MODE == MODE_HALF) { const int u = __shfl_up_sync(0xffffffffu, dincl, o); if (lane >= o) dincl += u; }
                }
                const int tD = (MODE == MODE_HALF) ? __shfl_sync(0xffffffffu, dincl, 31) : 0;
                const int total = tD + __shfl_sync(0xffffffffu, fincl, 31);
                // where this lane's segments sit in S, and what they read: (global record index - offset in S)
                const int offD = dincl - dlen, offF = tD + fincl - flen;
                const int srcD = d0 - offD, srcF = j0 - offF;

                // ---- passes over S --------------------------------------------------------------------------------
#pragma unroll 1
                for (int c0 = 0; c0 < total;) {
                    const int cn = min(total - c0, CP::CAPP - nl);
                    {
                        if (lane == 0) { fence_proxy_async(); mbar_expect_tx(mbar, (uint32_t)cn * REC * (F::AUX ? 2u : 1u)); }
                        __syncwarp();
                        const int wdst = nl - c0;   // S offset -> buffer slot
#pragma unroll
                        for (int part = (MODE == MODE_HALF ? 0 : 1); part < 2; ++part) {
                            const int off = part ? offF : offD, len = part ? flen : dlen, src = part ? srcF : srcD;
                            const int lo = max(off, c0), hi = min(off + len, c0 + cn);
                            if (hi > lo) {
                                bulk_g2s(buf_addr + (uint32_t)(wdst + lo) * REC, rec + (src + lo), (uint32_t)(hi - lo) * REC, mbar);
                                if constexpr (F::AUX) bulk_g2s(abuf_addr + (uint32_t)(wdst + lo) * REC, f.aux_j() + (src + lo), (uint32_t)(hi - lo) * REC, mbar);
                            }
                        }
                        mbar_wait(mbar, parity);
                        parity ^= 1u;
                        __syncwarp();
                    }
                    // far-away dummies round the staged records up to whole cull steps
                    if (cn + lane < ((cn + 31) & ~31)) strec(buf + nl + cn + lane, -huge_coord<T>(), T(0), T(0), (tag_t)0);
                    __syncwarp();
                    int ns = nl;
                    {
                        const RecT<T>* q = buf + nl + lane;
                        int k0 = 0;
                        if (MODE == MODE_HALF) {
                            const int kend = min(tD - c0, cn);   // staged records [0, kend) of this pass are keyed
#pragma unroll 1
                            for (; k0 < kend; k0 += 32, q += 32) {
                                const unsigned m = cull_step(q, ns);
                                const int left = kend - k0;
                                nk = ns + __popc(m & ((left >= 32) ? 0xffffffffu : ((1u << left) - 1u)));
                                ns += __popc(m);
                            }
                        }
#pragma unroll 1
                        for (; k0 < cn; k0 += 32, q += 32) ns += __popc(cull_step(q, ns));
                    }
                    c0 += cn;
                    const bool last = (c0 >= total);
                    int nfull = ns >> 5;
                    if (last && (ns & 31)) {   // pad the partial chunk with far-away dummies and sweep it too
                        if ((ns & 31) + lane < 32) strec(buf + ns + lane, -huge_coord<T>(), T(0), T(0), (tag_t)0);
                        nfull += 1;
                        ns = nfull * 32;
                    }
                    __syncwarp();
                    // one loop per chunk kind (each kind is inlined exactly once): no per-chunk dispatch
                    auto sweep_chunks = [&](auto half_tag) {
                        if (MODE == MODE_TRI) {
#pragma unroll 1
                            for (int ch = 0; ch < nfull; ++ch) chunk(std::integral_constant<int, N3_TRI>{}, half_tag, ch * 32 + lane);
                        } else if (any_ghost_i) {
#pragma unroll 1
                            for (int ch = 0; ch < nfull; ++ch) chunk(std::integral_constant<int, N3_GENERAL>{}, half_tag, ch * 32 + lane);
                        } else {
                            const int nkc = min(nfull, (nk + 31) >> 5);     // chunks that hold keyed partners come first
#pragma unroll 1
                            for (int ch = 0; ch < nkc; ++ch) chunk(std::integral_constant<int, N3_KEYED>{}, half_tag, ch * 32 + lane);
#pragma unroll 1
                            for (int ch = nkc; ch < nfull; ++ch) chunk(std::integral_constant<int, N3_PLAIN>{}, half_tag, ch * 32 + lane);
                        }
                    };
                    sweep_chunks(std::integral_constant<int, 0>{});
                    if constexpr (NH == 2) {
                        if (nfull > 0) reduce_fi(0);
                        sweep_chunks(std::integral_constant<int, 1>{});
                        if (nfull > 0) reduce_fi(IH);
                    }
                    const int rem = ns & 31;
                    if (nfull > 0 && rem > 0) {
                        // carry the partial chunk to the front
                        const RecT<T> rq = ldrec_s(buf + nfull * 32 + lane);
                        RecT<T> aq = rq;
                        if constexpr (F::AUX) aq = ldrec_s(abuf + nfull * 32 + lane);
                        __syncwarp();
                        if (lane < rem) {
                            strec(buf + lane, rq.x, rq.y, rq.z, rq.tag);
                            if constexpr (F::AUX) strec(abuf + lane, aq.x, aq.y, aq.z, aq.tag);
                        }
                    }
                    nk = max(nk - nfull * 32, 0);
                    nl = rem;
                    __syncwarp();
                }
            }
        }

        if constexpr (NH == 1) reduce_fi(0);   // Float32: once per tile
        e_acc += (double)e_tile;
    }
    {
        __shared__ double sm[4];
        const double e = block_sum(e_acc, sm);
        if (threadIdx.x == 0 && e != 0.0) atomicAdd(&a.res->f[RB_ENERGY], e * f.energy_scale());
    }
}

// accumulator rows -> the caller's per-particle force array:
// out[idx] = (accumulate ? out[idx] : 0) + scale * R^-1 f(row of idx).  One thread per PARTICLE: the output is written in
// particle order (coalesced) and the particle's row is gathered through slot_of[idx] (k_place) -- rows are in record
// order, image rows are never written (images add into their original's row).  slot_of == nullptr: rows in particle order
// (triclinic cells, whose exactly-once rule compares particle indices).
template <class T> struct Row4;
template <> struct Row4<float> { typedef float4 type; };
template <> struct Row4<double> { typedef double4 type; };
// the accumulator rows back to zero for the next map: one coalesced fill (scattered 16-byte zero stores from the gather
// above would be partial-sector writes)
template <class T>
__global__ void __launch_bounds__(256) k_zero_rows(T* __restrict__ facc, const int* __restrict__ dscal, int rec_cap) {
    typedef typename Row4<T>::type row_t;
    const int ntot = min(dscal[DS_NTOT], rec_cap);
    row_t z; z.x = T(0); z.y = T(0); z.z = T(0); z.w = T(0);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ntot; k += gridDim.x * blockDim.x) reinterpret_cast<row_t*>(facc)[k] = z;
}
template <class T>
__global__ void __launch_bounds__(256)
k_force_finish(const int* __restrict__ slot_of, T* __restrict__ facc, const int* __restrict__ dscal, int rec_cap, int n, T* __restrict__ out, int dim,
               T scale, int accumulate, int rotated, const __grid_constant__ GeomT<T> g) {
    typedef typename Row4<T>::type row_t;
    if (dscal[DS_NTOT] > rec_cap) return;   // overflowed build
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int k = slot_of ? slot_of[idx] : idx;
    if (k >= rec_cap) return;
    const row_t v = *(reinterpret_cast<const row_t*>(facc) + k);   // the rows are zeroed by the next build (k_gather) or by k_zero_rows
    T fx = v.x * scale, fy = v.y * scale, fz = v.z * scale;
    if (rotated) {
        const T p = g.inv_rot[0] * fx + g.inv_rot[1] * fy + g.inv_rot[2] * fz;
        const T q = g.inv_rot[3] * fx + g.inv_rot[4] * fy + g.inv_rot[5] * fz;
        const T s = g.inv_rot[6] * fx + g.inv_rot[7] * fy + g.inv_rot[8] * fz;
        fx = p; fy = q; fz = s;
    }
    T* o = out + (size_t)idx * dim;
    if (accumulate) { o[0] += fx; o[1] += fy; if (dim == 3) o[2] += fz; }
    else { o[0] = fx; o[1] = fy; if (dim == 3) o[2] = fz; }
}

}  // namespace clm

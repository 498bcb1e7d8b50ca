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
// particle's REAL record (an image points at its original) | GHOST: images add straight into their original's row and
// k_force_finish gathers the real rows into the caller's particle order (scale, inverse rotation, reset/accumulate).
//
// Exactly-once rules, identical pair sets to k_sweep<MODE_HALF / MODE_TRI> (which are bit-exact against the oracle):
//   MODE_HALF  forward reference rows / cells: every record, except image-image pairs; partners in the reference cells
//              the tile itself touches ("direct" part) carry their reference cell index as a key: later cell -> forward
//              rule, same cell -> real i only, image partner or later record slot.
//   MODE_TRI   full stencil, i real, slot of i's real record < slot of j's real record (any strict total order on the
//              particles selects one of the two symmetric evaluations; the reference uses the particle index).
#pragma once
#include "clm_sweep.cuh"

namespace clm {

static_assert(TILE_I == 8, "the reduce-scatter of the i-side accumulators is written for 8 particles per tile");
constexpr uint32_t N3_GHOST = 0x80000000u, N3_SLOT = 0x7fffffffu;   // rec_n3 4th word: slot of the real record | GHOST
constexpr int N3_KEYED_CAP = 128;                   // records of the keyed ("direct") part staged per pass
constexpr int N3_KBUF = N3_KEYED_CAP + 32;          // keys: one per keyed slot + the carried partial chunk
constexpr int N3_IBUF_BYTES = 8 * 4 * 8 + 8 * 16 + N3_KBUF * 4;   // per warp: TILE_I x (x, y, z, w) of T (<= double), TILE_I x int4, keys

// per-warp staging buffer: with ~80 registers per thread 6 CTAs are resident per SM whatever the buffer size up to
// 8 KB, so the buffer is sized for the whole forward part of a typical tile in one pass
#ifndef CLM_N3_STAGE_BYTES_F32
#define CLM_N3_STAGE_BYTES_F32 8192
#endif
#ifndef CLM_N3_STAGE_BYTES_F64
#define CLM_N3_STAGE_BYTES_F64 8192
#endif
template <class T, bool AUX> struct N3Cap {
    static constexpr int SB = AUX ? 6144 : ((sizeof(T) == 4) ? CLM_N3_STAGE_BYTES_F32 : CLM_N3_STAGE_BYTES_F64);
    static constexpr int REC = (int)sizeof(RecT<T>);
    static constexpr int SLOTS = SB / REC;
    static constexpr int FWD = SLOTS - 32 - STAGE_PAD;   // staged per pass: the carried partial chunk (< 32) and the padding share the buffer
    static constexpr int MBAR_OFF = (SWEEP_THREADS / 32) * SB;
    static constexpr int ABUF_OFF = MBAR_OFF + 64;
    static constexpr int IBUF_OFF = ABUF_OFF + (AUX ? (SWEEP_THREADS / 32) * SB : 0);
};
template <class T, bool AUX> struct N3Smem { static constexpr int value = N3Cap<T, AUX>::IBUF_OFF + (SWEEP_THREADS / 32) * N3_IBUF_BYTES; };

__device__ __forceinline__ void red_add3(float* p, float x, float y, float z) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}
__device__ __forceinline__ void red_add3(double* p, double x, double y, double z) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(y) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 2), "d"(z) : "memory");
}
// slot word of a staged rec_n3 record
__device__ __forceinline__ uint32_t slotword(const RecT<float>& r) { return r.tag; }
__device__ __forceinline__ uint32_t slotword(const RecT<double>& r) { return (uint32_t)r.tag; }

// ---- pair-force functors: fs(hit, d2, e, wi, wj) returns the scalar s with F_i += s (x_i - x_j), F_j -= s (x_i - x_j);
//      e accumulates the (unscaled) pair energy of the tile ----
template <class T, bool NORM, bool ENERGY> struct N3LJ {
    T c6, c12, s2, escale, fscale;      // see FLJ (clm_sweep.cuh): q = (c12/c6)/d2^3 form when both constants are positive
    static constexpr bool AUX = false;
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
#ifndef CLM_N3_MINB_F64
#define CLM_N3_MINB_F64 4
#endif
template <class T, bool AUX> struct N3MinBlocks { static constexpr int value = (sizeof(T) == 4) ? (AUX ? 4 : CLM_N3_MINB_F32) : (AUX ? 3 : CLM_N3_MINB_F64); };

template <class T, int MODE, class F>
__global__ void __launch_bounds__(SWEEP_THREADS, N3MinBlocks<T, F::AUX>::value)
k_sweep_n3(const __grid_constant__ SweepArgs<T> a, const __grid_constant__ F f, const RecT<T>* __restrict__ rec_tag, T* __restrict__ facc) {
    typedef TagT<T> TG;
    typedef typename TG::type tag_t;
    typedef N3Cap<T, F::AUX> CP;
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lf = a.lf, sub = a.sub, hww = 2 * lf + 1;
    const unsigned smagic = a.sub_magic;
    auto div_sub = [&](int v) { return (sub == 1) ? v : (int)__umulhi((unsigned)v, smagic); };
    const int nrows_st = (a.nz == 1) ? hww : hww * hww;
    constexpr uint32_t REC = (uint32_t)sizeof(RecT<T>);
    RecT<T>* const buf = reinterpret_cast<RecT<T>*>(dsm_raw + warp * CP::SB);
    RecT<T>* const abuf = reinterpret_cast<RecT<T>*>(dsm_raw + CP::ABUF_OFF + warp * CP::SB);
    unsigned char* const ibraw = dsm_raw + CP::IBUF_OFF + warp * N3_IBUF_BYTES;
    Vec4S<T>* const ipos = reinterpret_cast<Vec4S<T>*>(ibraw);             // [TILE_I]: x, y, z, functor weight
    int4* const ikey = reinterpret_cast<int4*>(ibraw + TILE_I * 4 * 8);     // [TILE_I]: reference cell x, own slot, image flag, slot of the real record
    int* const kbuf = reinterpret_cast<int*>(ibraw + TILE_I * 4 * 8 + TILE_I * 16);   // [N3_KBUF]: reference cell x of the keyed partners
    const uint32_t abuf_addr = smem_u32(abuf), buf_addr = smem_u32(buf);
    const uint32_t mbar = smem_u32(dsm_raw + CP::MBAR_OFF + warp * 8);
    uint32_t parity = 0;
    if (lane == 0) { mbar_init(mbar, 1); fence_proxy_async(); }
    __syncwarp();
    if (a.dscal[DS_NTOT] > a.rec_cap_i) return;   // overflowed build: the host repeats build + map
    double e_acc = 0.0;
    const int ntiles = a.dscal[DS_NTILES];
    const T rc2 = a.rc2;
    const unsigned lt = (1u << lane) - 1u;
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
        const RecT<T> ri = ldrec(rec_tag + ki);                      // coordinates + tag (HOME / GHOST / FOREIGN)
        const bool ghost_i = (ri.tag & TG::GHOST) != 0;
        const bool active = valid && ((ri.tag & TG::FOREIGN) == 0) && ((MODE == MODE_HALF) ? ((ri.tag & TG::HOME) != 0) : !ghost_i);
        T blo[3], bhi[3];
        {
            const T inf = CUDART_INF_T<T>();
            blo[0] = tile_min(active ? ri.x : inf); blo[1] = tile_min(active ? ri.y : inf); blo[2] = tile_min(active ? ri.z : inf);
            bhi[0] = tile_max(active ? ri.x : -inf); bhi[1] = tile_max(active ? ri.y : -inf); bhi[2] = tile_max(active ? ri.z : -inf);
        }
        const int iy = tl.yz & 0xffff, iz = tl.yz >> 16, tile_row = iz * a.ny + iy;
        const int cxa = tl.cx & 0xffff, cxb = tl.cx >> 16;
        int rfx_i = 0;
        if (MODE == MODE_HALF) {
            const int* cs = a.cell_start_i + (size_t)tile_row * (a.nx + 1);
            int cx = cxa;
            while (cx < cxb && cs[cx + 1] <= ki) ++cx;
            rfx_i = div_sub(cx);
        }
        const int ry_i = div_sub(iy), rz_i = div_sub(iz);
        const bool any_ghost_i = __ballot_sync(0xffffffffu, active && ghost_i) != 0u;
        const int slot_i = (int)(slotword(ldrec(a.rec_j + ki)) & N3_SLOT);   // slot of i's real record (own slot for a real particle)
        __syncwarp();   // every lane is done with the previous tile's i-side data
        if (lane < TILE_I) {
            Vec4S<T> v;
            v.x = active ? ri.x : huge_coord<T>(); v.y = ri.y; v.z = ri.z; v.w = (F::AUX && active) ? f.wi(ki) : T(0);
            ipos[lane] = v;
            ikey[lane] = make_int4(rfx_i, ki, ghost_i ? 1 : 0, slot_i);
        }
        __syncwarp();
        T fi[TILE_I][3];
#pragma unroll
        for (int i = 0; i < TILE_I; ++i) fi[i][0] = fi[i][1] = fi[i][2] = T(0);
        T e_tile = T(0);
        // keyed ("direct") part along the row (MODE_HALF): the reference cells rfa .. rfb the tile touches
        const int rfa = div_sub(cxa), rfb = div_sub(cxb);
        const int xlo_dir = rfa * sub, xhi_dir = min((rfb + 1) * sub, a.nx);
        // the partner list of the tile: survivors of the cull wait at buf[0 .. nl); the first nkey of them carry keys
        int nl = 0, nkey = 0;

        // ---- one chunk = 32 partners, one per lane; one warp step per particle i of the tile ---------------------
        // PLAIN: forward partners of a tile without image particles i (every pair counts); KEYED: chunks that hold keyed
        // partners; GENERAL: tiles with image particles i; TRI: triclinic rule
        auto chunk = [&](auto kind_tag, const int s) {
            constexpr int KIND = decltype(kind_tag)::value;
            const RecT<T> rj = ldrec_s(buf + s);
            T wj = T(0);
            if constexpr (F::AUX) wj = ldrec_s(abuf + s).x;
            const uint32_t sw = slotword(rj);
            const bool gj = (sw & N3_GHOST) != 0u;
            const int slot_j = (int)(sw & N3_SLOT);
            int K1 = 0x7fffffff;
            const int K2 = (MODE == MODE_TRI) ? slot_j : (gj ? 0x7fffffff : slot_j);
            if (KIND == N3_KEYED || KIND == N3_GENERAL) { if (s < nkey) K1 = kbuf[s]; }
            T fjx = T(0), fjy = T(0), fjz = T(0);
#pragma unroll
            for (int i = 0; i < TILE_I; ++i) {
                const Vec4S<T> pi = ldvec4_s(ipos + i);
                const T dx = pi.x - rj.x, dy = pi.y - rj.y, dz = pi.z - rj.z;
                const T d2 = xfma(dz, dz, xfma(dy, dy, dx * dx));
                bool ok = true;
                if (KIND == N3_KEYED) { const int4 ik = ikey[i]; ok = (K1 > ik.x) || (K1 == ik.x && K2 > ik.y); }
                else if (KIND == N3_GENERAL) { const int4 ik = ikey[i]; ok = ik.z ? (K1 > ik.x && !gj) : ((K1 > ik.x) || (K1 == ik.x && K2 > ik.y)); }
                else if (KIND == N3_TRI) { const int4 ik = ikey[i]; ok = K2 > ik.w; }
                const bool hit = ok && (d2 <= rc2);
                const T sc = f.fs(hit, d2, e_tile, pi.w, wj);
                fi[i][0] = xfma(sc, dx, fi[i][0]); fi[i][1] = xfma(sc, dy, fi[i][1]); fi[i][2] = xfma(sc, dz, fi[i][2]);
                fjx = xfma(-sc, dx, fjx); fjy = xfma(-sc, dy, fjy); fjz = xfma(-sc, dz, fjz);
            }
            if (fjx != T(0) || fjy != T(0) || fjz != T(0)) red_add3(facc + (size_t)slot_j * 4, fjx, fjy, fjz);
        };

        for (int rb = 0; rb < nrows_st; rb += 32) {
            // ---- lane r classifies stencil row r (same rules as k_sweep) -------------------------------------------
            const int r = rb + lane;
            int j0 = 0, j1 = 0, rowbase = 0, dj0 = 0, dj1 = 0, bnd1 = 0x7fffffff;
            if (r < nrows_st) {
                const int dz = a.rdz[r], dy = a.rdy[r];
                const int z2 = iz + dz, y2 = iy + dy;
                const int w = a.hw[(dz + lf) * hww + dy + lf];
                bool use = (z2 >= 0 && z2 < a.nz && y2 >= 0 && y2 < a.ny && w >= 0);
                int rel = 1;
                if (MODE == MODE_HALF && use) {
                    const int rz_j = div_sub(z2), ry_j = div_sub(y2);
                    rel = (rz_j != rz_i) ? (rz_j - rz_i) : (ry_j - ry_i);
                    use = rel >= 0;
                }
                if (use) {
                    rowbase = (z2 * a.ny + y2) * (a.nx + 1);
                    const int xa = max(cxa - w, 0), xb = min(cxb + w, a.nx - 1);
                    j0 = a.cell_start_j[rowbase + xa];
                    j1 = a.cell_start_j[rowbase + xb + 1];
                    if (MODE == MODE_HALF && rel == 0) {
                        const int* csj = a.cell_start_j + rowbase;
                        const int tsplit = csj[xhi_dir];
                        dj0 = max(j0, csj[xlo_dir]);
                        dj1 = min(j1, tsplit);
                        j0 = max(j0, tsplit);
                        if (rfb > rfa) bnd1 = csj[(rfa + 1) * sub];   // first record of the tile's second reference cell in this row
                    }
                }
            }
            const int dlen = (MODE == MODE_HALF) ? max(dj1 - dj0, 0) : 0, flen = max(j1 - j0, 0);
            // prefixes of the segment lengths: where each row's records go in the staged sequence of its part
            int dincl = dlen, fincl = flen;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, fincl, o);
                if (lane >= o) fincl += v;
                if (MODE == MODE_HALF) { const int u = __shfl_up_sync(0xffffffffu, dincl, o); if (lane >= o) dincl += u; }
            }
            const int dtotal = (MODE == MODE_HALF) ? __shfl_sync(0xffffffffu, dincl, 31) : 0, ftotal = __shfl_sync(0xffffffffu, fincl, 31);
            // ---- passes: the keyed ("direct") part first, then the forward part; each pass stages at most CAPP records
            //      behind the partners that wait at buf[0 .. nl), culls them against the tile's bounding box, compacts in
            //      place and sweeps the whole chunks of the list; the partial chunk at its end is carried to the next pass.
            //      The last forward pass of the batch sweeps everything. ----
            bool keyed = dtotal > 0;
            int c0 = 0;
            for (;;) {
                const int total = keyed ? dtotal : ftotal, capp = keyed ? N3_KEYED_CAP : CP::FWD;
                if (keyed && c0 >= total) { keyed = false; c0 = 0; continue; }
                const int len = keyed ? dlen : flen, off = keyed ? (dincl - dlen) : (fincl - flen), seg0 = keyed ? dj0 : j0;
                const int cn = min(total - c0, capp);
                const int lo = max(off, c0), hi = min(off + len, c0 + cn);
                if (cn > 0) {
                    if (lane == 0) { fence_proxy_async(); mbar_expect_tx(mbar, (uint32_t)cn * REC * (F::AUX ? 2u : 1u)); }
                    __syncwarp();
                    if (hi > lo) {
                        bulk_g2s(buf_addr + (uint32_t)(nl + lo - c0) * REC, a.rec_j + (seg0 + (lo - off)), (uint32_t)(hi - lo) * REC, mbar);
                        if constexpr (F::AUX) bulk_g2s(abuf_addr + (uint32_t)(nl + lo - c0) * REC, f.aux_j() + (seg0 + (lo - off)), (uint32_t)(hi - lo) * REC, mbar);
                    }
                    if (MODE == MODE_HALF && keyed) {
                        // keys of the staged records, row by row: reference cell (along the row) = rfa + the number of reference-cell
                        // starts of that row at or before the record's slot in the global array
                        unsigned md = __ballot_sync(0xffffffffu, hi > lo);
                        while (md) {
                            const int src = __ffs(md) - 1;
                            md &= md - 1;
                            const int slo = __shfl_sync(0xffffffffu, lo, src), shi = __shfl_sync(0xffffffffu, hi, src);
                            const int sbase = __shfl_sync(0xffffffffu, seg0 - off, src), sb1 = __shfl_sync(0xffffffffu, bnd1, src);
                            const int srb = __shfl_sync(0xffffffffu, rowbase, src);
#pragma unroll 1
                            for (int p = slo + lane; p < shi; p += 32) {
                                const int jc = sbase + p;
                                int key = rfa + ((jc >= sb1) ? 1 : 0);
                                if (rfb > rfa + 1) {   // a sparse row: the tile spans more than two reference cells
#pragma unroll 1
                                    for (int rc = rfa + 2; rc <= rfb; ++rc) key += (jc >= a.cell_start_j[srb + rc * sub]) ? 1 : 0;
                                }
                                kbuf[nl + p - c0] = key;
                            }
                        }
                    }
                    mbar_wait(mbar, parity);
                    parity ^= 1u;
                    __syncwarp();
                }
                if (cn + lane < ((cn + 31) & ~31)) strec(buf + nl + cn + lane, -huge_coord<T>(), T(0), T(0), (tag_t)0);
                __syncwarp();
                int ns = nl;
                {
                    const RecT<T>* q = buf + nl + lane;
#pragma unroll 1
                    for (int k0 = 0; k0 < cn; k0 += 32, q += 32) {
                        const RecT<T> rq = ldrec_s(q);
                        RecT<T> aq = rq;
                        if constexpr (F::AUX) aq = ldrec_s(abuf + (q - buf));
                        int key = 0;
                        if (MODE == MODE_HALF && keyed) key = kbuf[nl + k0 + lane];
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
                            if (MODE == MODE_HALF && keyed) kbuf[pos] = key;
                        }
                        ns += __popc(m);
                    }
                }
                if (keyed) nkey = ns;      // the keyed part comes first: every waiting partner is keyed
                const bool last = !keyed && (c0 + cn >= total);
                int nfull = ns >> 5;
                if (last && (ns & 31)) {   // pad the partial chunk with far-away dummies and sweep it too
                    if ((ns & 31) + lane < 32) strec(buf + ns + lane, -huge_coord<T>(), T(0), T(0), (tag_t)0);
                    nfull += 1;
                    ns = nfull * 32;
                }
                __syncwarp();
#pragma unroll 1
                for (int ch = 0; ch < nfull; ++ch) {
                    const int s = ch * 32 + lane;
                    if (MODE == MODE_TRI) chunk(std::integral_constant<int, N3_TRI>{}, s);
                    else if (any_ghost_i) chunk(std::integral_constant<int, N3_GENERAL>{}, s);
                    else if (ch * 32 < nkey) chunk(std::integral_constant<int, N3_KEYED>{}, s);
                    else chunk(std::integral_constant<int, N3_PLAIN>{}, s);
                }
                const int rem = ns & 31;
                if (nfull > 0 && rem > 0) {
                    // carry the partial chunk to the front
                    const RecT<T> rq = ldrec_s(buf + nfull * 32 + lane);
                    RecT<T> aq = rq;
                    if constexpr (F::AUX) aq = ldrec_s(abuf + nfull * 32 + lane);
                    int key = 0;
                    if (MODE == MODE_HALF && keyed) key = kbuf[nfull * 32 + lane];
                    __syncwarp();
                    if (lane < rem) {
                        strec(buf + lane, rq.x, rq.y, rq.z, rq.tag);
                        if constexpr (F::AUX) strec(abuf + lane, aq.x, aq.y, aq.z, aq.tag);
                        if (MODE == MODE_HALF && keyed) kbuf[lane] = key;
                    }
                }
                if (nfull > 0) nkey = keyed ? rem : 0;
                nl = rem;
                __syncwarp();
                c0 += cn;
                if (last) break;
            }
        }

        // ---- f_i: reduce-scatter of the TILE_I x 3 per-lane partial sums over the warp: after three halving steps the
        //      four lanes 4g .. 4g+3 hold the partial sums of particle g, two butterfly steps finish them ----
        {
            T v[24];
#pragma unroll
            for (int i = 0; i < TILE_I; ++i) { v[3 * i] = fi[i][0]; v[3 * i + 1] = fi[i][1]; v[3 * i + 2] = fi[i][2]; }
#pragma unroll
            for (int half = 12, o = 16; half >= 3; half >>= 1, o >>= 1) {
                const bool upper = (lane & o) != 0;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                    const T send = upper ? v[k] : v[k + half];
                    const T keep = upper ? v[k + half] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
#pragma unroll
            for (int o = 2; o >= 1; o >>= 1) {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], o); v[1] += __shfl_xor_sync(0xffffffffu, v[1], o); v[2] += __shfl_xor_sync(0xffffffffu, v[2], o);
            }
            const int g = lane >> 2;
            const int slot_g = __shfl_sync(0xffffffffu, slot_i, g);
            const bool act_g = __shfl_sync(0xffffffffu, active ? 1 : 0, g) != 0;
            if ((lane & 3) == 0 && act_g && (v[0] != T(0) || v[1] != T(0) || v[2] != T(0))) red_add3(facc + (size_t)slot_g * 4, v[0], v[1], v[2]);
        }
        e_acc += (double)e_tile;
    }
    {
        __shared__ double sm[4];
        const double e = block_sum(e_acc, sm);
        if (threadIdx.x == 0 && e != 0.0) atomicAdd(&a.res->f[RB_ENERGY], e * f.energy_scale());
    }
}

// record-ordered accumulator rows -> the caller's per-particle force array (particle order), and back to zero:
// out[idx] = (accumulate ? out[idx] : 0) + scale * R^-1 f.  Every particle has exactly one real record.
template <class T>
__global__ void __launch_bounds__(256)
k_force_finish(const RecT<T>* __restrict__ rec_tag, T* __restrict__ facc, const int* __restrict__ dscal, int rec_cap, T* __restrict__ out, int dim,
               T scale, int accumulate, int rotated, const __grid_constant__ GeomT<T> g) {
    typedef TagT<T> TG;
    const int ntot = dscal[DS_NTOT];
    if (ntot > rec_cap) return;   // overflowed build
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ntot; k += gridDim.x * blockDim.x) {
        const typename TG::type tag = rec_tag[k].tag;
        if (tag & (TG::GHOST | TG::FOREIGN)) continue;   // image rows are never written: images add into their original's row
        T* row = facc + (size_t)k * 4;
        T fx = row[0] * scale, fy = row[1] * scale, fz = row[2] * scale;
        row[0] = T(0); row[1] = T(0); row[2] = T(0);
        if (rotated) {
            const T p = g.inv_rot[0] * fx + g.inv_rot[1] * fy + g.inv_rot[2] * fz;
            const T q = g.inv_rot[3] * fx + g.inv_rot[4] * fy + g.inv_rot[5] * fz;
            const T s = g.inv_rot[6] * fx + g.inv_rot[7] * fy + g.inv_rot[8] * fz;
            fx = p; fy = q; fz = s;
        }
        T* o = out + (size_t)(tag & TG::MASK) * dim;
        if (accumulate) { o[0] += fx; o[1] += fy; if (dim == 3) o[2] += fz; }
        else { o[0] = fx; o[1] = fy; if (dim == 3) o[2] = fz; }
    }
}

}  // namespace clm

// Shared device/host definitions of the B200 cutoff-pair engine.
//
// HBM layout (one per particle set):
//   rec[n_total]      cell-sorted packed particle records (x, y, z, tag): one 128-bit (F32) or two
//                     128-bit (F64) loads per particle; image ("ghost") particles are materialised
//                     exactly as the reference does (internals/Box.jl:556-566), with the same tag.
//   cell_start[]      first record of every cell of the DEVICE grid, row pitch nx + 1 (entry nx = end of the row).
//                     The device grid is the reference grid (Box.jl:209-220) with every reference cell split into
//                     `sub` sub-cells per dimension (sub = 1: identical grids); the LAST reference dimension runs
//                     fastest, so one row of cells along it is ONE contiguous range of rec[] (rows themselves are
//                     placed in arbitrary order: their bases come from an atomic counter, clm_build.cuh).
//   tiles[n_tiles]    work items: TI consecutive records of one row + the range of cells they span.
#pragma once
#ifdef __CUDACC_RTC__   // run-time compilation of user pair functions (clm_rtc.cu): no host headers
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace clm {

// ---- rounding-exact arithmetic (no FMA contraction): bit parity with the CPU oracle -------------
__device__ __forceinline__ float  xadd(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  xsub(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float  xmul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  xdiv(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float  xsqrt(float a)  { return __fsqrt_rn(a); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

// ---- packed particle records ----------------------------------------------------------------------
// tag: bit (W-1) = image particle (real == false), bit (W-2) = lives in a REFERENCE cell that contains at
// least one real particle ("home" cell: the cells the reference sweeps, self.jl:56-57), bit (W-3) = foreign:
// a particle (or an image of one) owned by another rank of a slab-decomposed system -- it is a partner j
// like any other record but never acts as particle i; low bits = 0-based index (owned particles first,
// then the foreign ones).
template <class T> struct RecT;
template <> struct __align__(16) RecT<float> {
    float x, y, z;
    uint32_t tag;
};
template <> struct __align__(32) RecT<double> {
    double x, y, z;
    uint64_t tag;
};
template <class T> struct TagT;
template <> struct TagT<float> {
    typedef uint32_t type;
    static constexpr uint32_t GHOST = 0x80000000u, HOME = 0x40000000u, FOREIGN = 0x20000000u, MASK = 0x1fffffffu;
};
template <> struct TagT<double> {
    typedef uint64_t type;
    static constexpr uint64_t GHOST = 0x8000000000000000ull, HOME = 0x4000000000000000ull, FOREIGN = 0x2000000000000000ull, MASK = 0x1fffffffffffffffull;
};

__device__ __forceinline__ RecT<float> ldrec(const RecT<float>* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    RecT<float> r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.tag = __float_as_uint(v.w);
    return r;
}
__device__ __forceinline__ RecT<double> ldrec(const RecT<double>* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    RecT<double> r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.tag = (uint64_t)__double_as_longlong(b.y);
    return r;
}
__device__ __forceinline__ void strec(RecT<float>* p, float x, float y, float z, uint32_t tag) {
    *reinterpret_cast<float4*>(p) = make_float4(x, y, z, __uint_as_float(tag));
}
__device__ __forceinline__ void strec(RecT<double>* p, double x, double y, double z, uint64_t tag) {
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(x, y);
    q[1] = make_double2(z, __longlong_as_double((long long)tag));
}

// ---- geometry handed to the kernels (all values already rounded to T on the host) ------------------
template <class T> struct GeomT {
    // wrap: p = M \ x through the closed-form cofactor solve, frac, M * p, then rotation * p
    T cof[9];       // row-major: frac_k = ((cof[3k]*x0 + cof[3k+1]*x1) + cof[3k+2]*x2) / det   (3-D)
    T det;
    T m[9];         // input unit cell, row-major rows: out_k = (m[3k]*p0 + m[3k+1]*p1) + m[3k+2]*p2
    T rot[9];       // rotation, row-major
    T inv_rot[9];   // inverse rotation, row-major
    T shift[27][3]; // aligned_unit_cell * idx for idx in {-1,0,1}^N, first index fastest
    T cb_min[3], cb_max[3], cs[3];
    T cutoff, cutoff_sqr;
    int nc[3];      // reference grid (Box.nc); 2-D: nc[2] = 1
    int sub;        // sub-cells per reference cell and dimension of the device grid
    int lcell;
    int dim;
    int cell_type;  // clm_cell_type
    int rotated;    // 1 iff rotation != identity (triclinic)
    // non-periodic systems that REUSE the box of the previous build (_limits_fit_in_box, src/internals/ParticleSystem.jl:165-174):
    // the coordinate limits that box was made from; a particle outside them flags DS_NOFIT and the build is redone with new limits
    int np_check;
    T np_lo[3], np_hi[3];
};

struct Tile {       // 16 bytes
    int k0;         // first record of the tile
    int cnt;        // number of records (<= TI)
    int yz;         // row coordinates: y | (z << 16)   (row = y + ny * z)
    int cx;         // cxa | (cxb << 16): x-range of the cells the records live in
};

constexpr int WARP = 32;
constexpr int CLM_ORTHO_CT = 0, CLM_TRICLINIC_CT = 1, CLM_NONPERIODIC_CT = 2;  // == enum clm_cell_type

template <class T> __device__ __forceinline__ T CUDART_INF_T();
template <> __device__ __forceinline__ float CUDART_INF_T<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double CUDART_INF_T<double>() { return __longlong_as_double(0x7ff0000000000000ll); }

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace clm

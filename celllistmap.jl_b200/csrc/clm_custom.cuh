// User-defined pair functions, compiled at run time with NVRTC (clm_rtc.cu) into the SAME sweep kernel as the
// compiled-in catalogue: the B200 replacement of the Julia closure `f(pair, output)` of pairwise!(f, sys)
// (src/API/pairwise.jl:48-63) together with the default output protocol copy_output / reset_output! / reducer! = `+`
// for numbers, static vectors and arrays thereof (src/API/parallel_custom.jl:53-54, :116-123, :213).
//
// A user functor is a stateless struct in namespace clm_user:
//
//     struct MyPair {
//         static constexpr int NSCALAR = 1;   // scalar outputs, summed over all pairs            (0..8)
//         static constexpr int NPART   = 3;   // per-particle output components (forces-like)     (0..4)
//         static constexpr int NAUX    = 1;   // per-particle input components (masses, charges)  (0..4)
//         static constexpr int HIST    = 0;   // 1: a histogram output (bin count chosen at call time)
//         // optional: bit k set = scalar output k (k < 4) is reduced with min / max instead of + (the custom reducer of
//         // src/API/parallel_custom.jl:196-214; written with out.min_scalar(k, v) / out.max_scalar(k, v), starts at +-Inf)
//         static constexpr unsigned SCALAR_MIN = 0, SCALAR_MAX = 0;
//         template <class T, class Out>
//         __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
//             const T w = par[0] * p.ai[0] * p.aj[0] / p.d();
//             out.add_scalar(0, w);                                   // output += w
//             for (int k = 0; k < 3; ++k) out.add_i(k, w * (p.y[k] - p.x[k]) / p.d2);   // output[i] += ...
//         }
//     };
//
// NeighborPair mirrors src/API/NeighborPair.jl:19-33 (i, j 1-based and unordered, x, y with y - x the minimum-image
// vector, d2, lazy d).  Exactly-once semantics are the reference's: functors without per-particle outputs run in the
// reference's own mode (forward stencil / index rule / cross), each pair once.  Functors WITH per-particle outputs run
// the full-shell sweep: operator() is called once per ORDERED pair and add_i() adds to particle p.i only (no atomics);
// scalar and histogram outputs of self-set systems are halved on output.  The functor must therefore be symmetric under
// the exchange of the two particles -- which the reference requires too, because the orientation of (i, j) is unspecified.
#pragma once
#include "clm_sweep.cuh"

namespace clm {

constexpr int CUSTOM_MAX_PAR = 16, CUSTOM_MAX_SCALAR = 8, CUSTOM_MAX_PART = 4, CUSTOM_MAX_AUX = 4;
constexpr int RC_MINMAX0 = 4;   // ResultBlock.c[4 .. 7]: min / max scalar outputs 0 .. 3, order-preserving encoding (below)

// optional members of a user functor: which scalar outputs are min / max reductions
template <class U, class = void> struct ScalarMinMask { static constexpr unsigned value = 0u; };
template <class U> struct ScalarMinMask<U, decltype((void)U::SCALAR_MIN)> { static constexpr unsigned value = U::SCALAR_MIN; };
template <class U, class = void> struct ScalarMaxMask { static constexpr unsigned value = 0u; };
template <class U> struct ScalarMaxMask<U, decltype((void)U::SCALAR_MAX)> { static constexpr unsigned value = U::SCALAR_MAX; };
// doubles as unsigned integers of the same order; 0 is below every number, so that a zeroed slot means "no value":
// max outputs are kept as atomicMax(enc(v)), min outputs as atomicMax(~enc(v))
__device__ __forceinline__ unsigned long long order_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <class T> struct NeighborPair {
    long long i, j;
    T x[3], y[3];
    T d2;
    const T* ai;   // side-array row of particle i (NAUX components), nullptr when NAUX == 0
    const T* aj;
    __device__ __forceinline__ T d() const { return xsqrt(d2); }
};

// device-resident min / max scalar outputs: decode the ordered-integer slots (accumulate: the value found in out joins in)
template <class T> __global__ void k_store_minmax(T* __restrict__ out, const unsigned long long* __restrict__ slots, unsigned min_mask, unsigned max_mask, int accumulate) {
    const int k = threadIdx.x;
    if (k >= 4 || !(((min_mask | max_mask) >> k) & 1u)) return;
    const bool is_min = ((min_mask >> k) & 1u) != 0u;
    unsigned long long c = slots[k];
    double v = is_min ? CUDART_INF_T<double>() : -CUDART_INF_T<double>();
    if (c != 0ull) {
        if (is_min) c = ~c;
        v = __longlong_as_double((long long)((c >> 63) ? (c & 0x7fffffffffffffffull) : ~c));
    }
    if (accumulate) v = is_min ? fmin(v, (double)out[k]) : fmax(v, (double)out[k]);
    out[k] = (T)v;
}

// what a user functor writes into
template <class T, int NS, int NP, bool HIST> struct PairOutput {
    T* s;
    T* pi;
    const HistBins<T, true>* hb;
    __device__ __forceinline__ void add_scalar(int k, T v) { s[k] += v; }
    __device__ __forceinline__ void min_scalar(int k, T v) { s[k] = fmin(s[k], v); }   // scalar k must be declared in SCALAR_MIN
    __device__ __forceinline__ void max_scalar(int k, T v) { s[k] = fmax(s[k], v); }   // ... in SCALAR_MAX
    __device__ __forceinline__ void add_i(int c, T v) { pi[c] += v; }
    // histogram output: counts[bin] += 1, sums[bin] += v; bins outside [0, nbins) are ignored
    __device__ __forceinline__ void add_hist(int bin, T v) { if (HIST && bin >= 0 && bin < hb->nbins) hb->add(bin, v); }
    __device__ __forceinline__ int nbins() const { return HIST ? hb->nbins : 0; }
};

// the data members do NOT depend on U: the host fills a FCustom<T, HostStub> and passes its bytes to the kernel
template <class T, class U> struct FCustom {
    T par[CUSTOM_MAX_PAR];
    const T* ax_i;     // side arrays gathered into record order, one record-sized slot (4 x T) per record
    const T* ax_j;
    T* part_out;       // n x NPART per-particle output
    int part_accumulate, rotated, dim, pad_;
    T inv_rot[9];
    HistBins<T, true> hb;

    static constexpr int NS = U::NSCALAR, NP = U::NPART, NA = U::NAUX;
    static constexpr bool HAS_HIST = (U::HIST != 0);
    static constexpr unsigned MINM = ScalarMinMask<U>::value, MAXM = ScalarMaxMask<U>::value;
    static_assert((MINM & MAXM) == 0u && (MINM | MAXM) < 16u && ((MINM | MAXM) >> (NS > 0 ? NS : 0)) == 0u,
                  "SCALAR_MIN / SCALAR_MAX: disjoint bit masks over the scalar outputs 0..3");
    // identity of the reduction of scalar k, in T / double
    __device__ static __forceinline__ T ident(int k) { return ((MINM >> k) & 1u) ? CUDART_INF_T<T>() : (((MAXM >> k) & 1u) ? -CUDART_INF_T<T>() : T(0)); }
    __device__ static __forceinline__ double fold(int k, double a, double b) { return ((MINM >> k) & 1u) ? fmin(a, b) : (((MAXM >> k) & 1u) ? fmax(a, b) : a + b); }
    static_assert(NS >= 0 && NS <= CUSTOM_MAX_SCALAR, "NSCALAR must be in 0..8");
    static_assert(NP >= 0 && NP <= CUSTOM_MAX_PART, "NPART must be in 0..4");
    static_assert(NA >= 0 && NA <= CUSTOM_MAX_AUX, "NAUX must be in 0..4");
    // scalar outputs: summed in T over one tile only, folded into a double per thread at the end of the tile (a Float32
    // accumulator that lives for the whole persistent kernel loses the 1e-5 bar on million-particle systems)
    struct Acc { double s[NS > 0 ? NS : 1]; };
    struct IAcc { T v[NP > 0 ? NP : 1]; T a[4]; T s[NS > 0 ? NS : 1]; };
    static constexpr bool NEEDS_BAND = false, EXACT_D2 = true, AUX = (NA > 0);

    __device__ __forceinline__ const RecT<T>* aux_j() const { return reinterpret_cast<const RecT<T>*>(ax_j); }
    __device__ void init(Acc& a) const {
#pragma unroll
        for (int k = 0; k < (NS > 0 ? NS : 1); ++k) a.s[k] = (double)ident(k);
        if (HAS_HIST) hb.init();
    }
    __device__ void begin(IAcc& p, const Ctx<T>& c) const {
#pragma unroll
        for (int k = 0; k < (NP > 0 ? NP : 1); ++k) p.v[k] = T(0);
#pragma unroll
        for (int k = 0; k < (NS > 0 ? NS : 1); ++k) p.s[k] = ident(k);
#pragma unroll
        for (int k = 0; k < 4; ++k) p.a[k] = (AUX && c.active && k < NA) ? ax_i[(size_t)c.ki * 4 + k] : T(0);
    }
    __device__ __forceinline__ void call(Acc&, IAcc& p, const Ctx<T>& c, const RecT<T>& rj, const T* aj, T d2) const {
        NeighborPair<T> np;
        np.i = (long long)(c.ri.tag & TagT<T>::MASK) + 1;
        np.j = (long long)(rj.tag & TagT<T>::MASK) + 1;
        if (rotated) {   // pair.x / pair.y are inv_rotation * coordinates (self.jl:171-178, cross.jl:117-124)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                np.x[k] = inv_rot[3 * k] * c.ri.x + inv_rot[3 * k + 1] * c.ri.y + inv_rot[3 * k + 2] * c.ri.z;
                np.y[k] = inv_rot[3 * k] * rj.x + inv_rot[3 * k + 1] * rj.y + inv_rot[3 * k + 2] * rj.z;
            }
        } else {
            np.x[0] = c.ri.x; np.x[1] = c.ri.y; np.x[2] = c.ri.z;
            np.y[0] = rj.x; np.y[1] = rj.y; np.y[2] = rj.z;
        }
        np.d2 = d2;
        np.ai = p.a; np.aj = aj;
        PairOutput<T, NS, NP, HAS_HIST> out;
        out.s = p.s; out.pi = p.v; out.hb = &hb;
        U()(np, par, out);
    }
    // functors without side arrays
    __device__ __forceinline__ void pair(Acc& a, IAcc& p, const Ctx<T>& c, bool hit, bool, const RecT<T>& rj, int, T, T, T, T d2) const {
        if (hit) call(a, p, c, rj, nullptr, d2);
    }
    // functors with side arrays: aj is the partner's staged side-array slot
    __device__ __forceinline__ void pair(Acc& a, IAcc& p, const Ctx<T>& c, bool hit, bool, const RecT<T>& rj, const RecT<T>& aj, T, T, T, T d2) const {
        if (hit) {
            const T av[4] = {aj.x, aj.y, aj.z, T(0)};
            call(a, p, c, rj, av, d2);
        }
    }
    __device__ void end(Acc& a, IAcc& p, const Ctx<T>& c) const {
#pragma unroll
        for (int k = 0; k < NS; ++k) a.s[k] = fold(k, a.s[k], (double)p.s[k]);
        if (NP == 0) return;
        T v[NP > 0 ? NP : 1];
#pragma unroll
        for (int k = 0; k < (NP > 0 ? NP : 1); ++k) {
            v[k] = p.v[k];
#pragma unroll
            for (int o = TILE_I; o < 32; o <<= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);   // combine the j-slices
        }
        if (!c.active || c.slice != 0) return;
        T* f = part_out + (size_t)(c.ri.tag & TagT<T>::MASK) * NP;
#pragma unroll
        for (int k = 0; k < (NP > 0 ? NP : 1); ++k) { if (part_accumulate) f[k] += v[k]; else f[k] = v[k]; }
    }
    __device__ void finish(Acc& a, ResultBlock* res) const {
        __shared__ double sm[4];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if (((MINM | MAXM) >> k) & 1u) {
                // min / max: warp butterfly, then one ordered-integer atomicMax per warp
                double v = a.s[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fold(k, v, __shfl_xor_sync(0xffffffffu, v, o));
                const bool is_min = ((MINM >> k) & 1u) != 0u;
                if ((threadIdx.x & 31) == 0 && v != (double)ident(k)) atomicMax(&res->c[RC_MINMAX0 + k], is_min ? ~order_bits(v) : order_bits(v));
            } else {
                const double s = block_sum(a.s[k], sm);
                if (threadIdx.x == 0) atomicAdd(&res->f[k], s);
            }
        }
        if (HAS_HIST) hb.flush();
    }
};

}  // namespace clm

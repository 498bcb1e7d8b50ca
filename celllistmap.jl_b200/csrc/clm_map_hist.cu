// clm_map_dist_hist / clm_map_pairvel: histogram maps of the catalogue.
#include <cmath>
#include <limits>
#include "clm_engine.cuh"

namespace clm {

static inline size_t hist_smem(int nbins, bool priv, size_t sum_bytes) {
    const size_t slots = (size_t)nbins * (priv ? SWEEP_THREADS : 1);
    return ((slots * 4 + 15) / 16) * 16 + slots * sum_bytes;
}

template <class T> int Engine<T>::map_dist_hist(const void* width, int nbins, int flags, int64_t* counts) {
    if (!width || !counts) return fail(CLM_ERR_ARGUMENT, "width / counts pointer is NULL");
    if (nbins < 1 || nbins > 2048) return fail(CLM_ERR_ARGUMENT, "nbins must be in 1..2048");
    for (;;) {
    if (int rc = prepare_map(flags)) return rc;
    CLM_CK(d_hcount.ensure((size_t)nbins));
    CLM_CK(cudaMemsetAsync(d_hcount.p, 0, (size_t)nbins * sizeof(unsigned long long), stream));
    auto run = [&](auto fn) -> int {
        fn.width = *(const T*)width;
        fn.hb.nbins = nbins; fn.hb.priv = (nbins <= NB_PRIV_MAX) ? 1 : 0; fn.hb.off = StageTotal<T, false>::value; fn.hb.g_counts = d_hcount.p; fn.hb.g_sums = nullptr;
        return launch_reduce(fn, hist_smem(nbins, fn.hb.priv != 0, 0));
    };
    if (int rc = (nbins <= NB_PRIV_MAX) ? run(FHist<T, 1>()) : run(FHist<T, 0>())) return rc;
    const int v = build_validate();
    if (v == CLM_RETRY_INTERNAL) continue;
    if (v) return v;
    break;
    }
    std::vector<unsigned long long> hc;
    if (!(flags & CLM_OUT_DEVICE)) {
        hc.resize((size_t)nbins);
        CLM_CK(cudaMemcpyAsync(hc.data(), d_hcount.p, hc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
    }
    if (int rc = store_i64(counts, d_hcount.p, hc.data(), nbins, flags)) return rc;
    return finish_map(flags);
}

// smallest x >= 0 with sqrt_rn(x) > edge (std::sqrt is correctly rounded, like the device's __fsqrt_rn / __dsqrt_rn)
template <class T> static T sqrt_threshold(T edge) {
    if (!(edge >= T(0))) return (edge != edge) ? std::numeric_limits<T>::infinity() : T(0);   // NaN edge: never exceeded; negative: always
    if (std::isinf(edge)) return std::numeric_limits<T>::infinity();
    const T inf = std::numeric_limits<T>::infinity();
    T x = edge * edge;
    if (std::isinf(x)) x = std::numeric_limits<T>::max();
    while (x > T(0) && std::sqrt(x) > edge) x = std::nextafter(x, T(-1));       // walk down to a value that does not exceed
    while (x < inf && !(std::sqrt(x) > edge)) x = std::nextafter(x, inf);       // then up to the first one that does
    return x;
}

template <class T> int Engine<T>::map_pairvel(const void* vx, const void* vy, const void* rbins, int nbins, int flags, int64_t* counts, void* sums) {
    if (!vx || !rbins || !counts || !sums) return fail(CLM_ERR_ARGUMENT, "velocity / rbins / output pointer is NULL");
    if (two_sets && !vy) return fail(CLM_ERR_ARGUMENT, "velocities of the second set are required for a two-set system");
    if (nbins < 1 || nbins > 1024) return fail(CLM_ERR_ARGUMENT, "nbins must be in 1..1024");
    if (int rc = build()) return rc;   // per-record side arrays are sized by the record count: validated build first
    if (int rc = prepare_map(flags)) return rc;
    const bool dev = (flags & CLM_OUT_DEVICE) != 0;
    if (int rc = gather_aux(0, (const T*)vx, dim, geom.rotated != 0, dev)) return rc;
    if (two_sets) { if (int rc = gather_aux(1, (const T*)vy, dim, geom.rotated != 0, dev)) return rc; }
    CLM_CK(d_hcount.ensure((size_t)nbins));
    CLM_CK(d_hsum.ensure((size_t)nbins));
    CLM_CK(d_rbins.ensure((size_t)nbins + 1));
    CLM_CK(cudaMemsetAsync(d_hcount.p, 0, (size_t)nbins * sizeof(unsigned long long), stream));
    CLM_CK(cudaMemsetAsync(d_hsum.p, 0, (size_t)nbins * sizeof(double), stream));
    CLM_CK(cudaMemcpyAsync(d_rbins.p, rbins, ((size_t)nbins + 1) * sizeof(T), cudaMemcpyHostToDevice, stream));
    const int aux_bytes = (SWEEP_THREADS / 32) * StageBytes<T, true>::value;   // side-array staging buffers precede the bins
    auto run = [&](auto fn) -> int {
        fn.v_i = sets[0].aux.p; fn.v_j = sets[two_sets ? 1 : 0].aux.p; fn.rbins = d_rbins.p;
        fn.inline_edges = (nbins + 1 <= VEL_EDGES_INLINE) ? 1 : 0;
        for (int e = 0; e < VEL_EDGES_INLINE; ++e) fn.thr2[e] = (fn.inline_edges && e <= nbins) ? sqrt_threshold(((const T*)rbins)[e]) : std::numeric_limits<T>::infinity();
        fn.hb.nbins = nbins; fn.hb.priv = (nbins <= NB_PRIV_MAX) ? 1 : 0; fn.hb.off = StageTotal<T, true>::value + aux_bytes; fn.hb.g_counts = d_hcount.p; fn.hb.g_sums = d_hsum.p;
        return launch_reduce(fn, (size_t)aux_bytes + hist_smem(nbins, fn.hb.priv != 0, sizeof(T)));
    };
    int lrc;
    if (nbins + 1 <= 8) lrc = run(FVel<T, 1, 1>());                       // <= 7 bins: private bins (NB_PRIV_MAX = 16)
    else if (nbins <= NB_PRIV_MAX) lrc = run(FVel<T, 2, 1>());
    else if (nbins + 1 <= VEL_EDGES_INLINE) lrc = run(FVel<T, 2, 0>());
    else lrc = run(FVel<T, 0, 0>());
    if (lrc) return lrc;
    std::vector<unsigned long long> hc;
    std::vector<double> hs;
    if (!dev) {
        hc.resize((size_t)nbins); hs.resize((size_t)nbins);
        CLM_CK(cudaMemcpyAsync(hc.data(), d_hcount.p, hc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaMemcpyAsync(hs.data(), d_hsum.p, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
    }
    if (int rc = store_i64(counts, d_hcount.p, hc.data(), nbins, flags)) return rc;
    if (int rc = store_real(sums, d_hsum.p, hs.data(), nbins, 1.0, flags)) return rc;
    return finish_map(flags);
}

template int Engine<float>::map_dist_hist(const void*, int, int, int64_t*);
template int Engine<double>::map_dist_hist(const void*, int, int, int64_t*);
template int Engine<float>::map_pairvel(const void*, const void*, const void*, int, int, int64_t*, void*);
template int Engine<double>::map_pairvel(const void*, const void*, const void*, int, int, int64_t*, void*);

}  // namespace clm

// clm_map_dist_hist / clm_map_pairvel: histogram maps of the catalogue.
#include <cmath>
#include <limits>
#include "clm_engine.cuh"

namespace clm {

static inline size_t hist_smem(int nbins, bool priv, size_t sum_bytes) {
    const size_t slots = (size_t)nbins * (priv ? SWEEP_THREADS : 1);
    return ((slots * 4 + 15) / 16) * 16 + slots * sum_bytes;
}

// thresholds of the distance histogram: thr[b] = smallest d2 >= 0 with floor(sqrt_rn(d2) / width) >= b, b = 0 .. nbins,
// found by walking neighbouring floating-point numbers around the real-number guess (std::sqrt and the division are
// correctly rounded, like the device's __fsqrt_rn / __fdiv_rn)
template <class T> static bool hist_thresholds(T width, int nbins, std::vector<T>& thr) {
    if (!(width > T(0)) || std::isinf(width)) return false;
    const T inf = std::numeric_limits<T>::infinity();
    auto bin_of_r = [&](T r) { return std::floor(r / width); };
    thr.assign((size_t)nbins + 1, T(0));
    for (int b = 1; b <= nbins; ++b) {
        // smallest r with floor(r / width) >= b: walk from the real-number guess b * width (a few ulps at most)
        T r = (T)b * width;
        if (std::isinf(r)) { thr[b] = inf; continue; }
        int steps = 0;
        while (r > T(0) && bin_of_r(std::nextafter(r, T(-1))) >= (T)b) { r = std::nextafter(r, T(-1)); if (++steps > 64) return false; }
        while (bin_of_r(r) < (T)b) { r = std::nextafter(r, inf); if (++steps > 128) return false; }
        // smallest d2 with sqrt_rn(d2) >= r
        T x = r * r;
        if (std::isinf(x)) { thr[b] = inf; continue; }
        steps = 0;
        while (x > T(0) && std::sqrt(std::nextafter(x, T(-1))) >= r) { x = std::nextafter(x, T(-1)); if (++steps > 64) return false; }
        while (std::sqrt(x) < r) { x = std::nextafter(x, inf); if (++steps > 128) return false; }
        thr[b] = x;   // a walk that does not converge (pathological width: denormal, huge) falls back to the direct form
    }
    return true;
}

template <class T> int Engine<T>::map_dist_hist(const void* width, int nbins, int flags, int64_t* counts) {
    if (!width || !counts) return fail(CLM_ERR_ARGUMENT, "width / counts pointer is NULL");
    if (nbins < 1 || nbins > 2048) return fail(CLM_ERR_ARGUMENT, "nbins must be in 1..2048");
    std::vector<T> thr;
    const bool use_thr = hist_thresholds(*(const T*)width, nbins, thr);
    for (;;) {
    if (int rc = prepare_map(flags)) return rc;
    CLM_CK(d_hcount.ensure((size_t)nbins));
    CLM_CK(cudaMemsetAsync(d_hcount.p, 0, (size_t)nbins * sizeof(unsigned long long), stream));
    if (use_thr) {
        CLM_CK(d_rbins.ensure((size_t)nbins + 1));
        CLM_CK(cudaMemcpyAsync(d_rbins.p, thr.data(), ((size_t)nbins + 1) * sizeof(T), cudaMemcpyHostToDevice, stream));
    }
    auto run = [&](auto fn) -> int {
        fn.width = *(const T*)width; fn.inv_width = T(1) / fn.width; fn.thr = use_thr ? d_rbins.p : nullptr;
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
template <class T> static T sqrt_threshold(T edge, bool& converged) {
    if (!(edge >= T(0))) return (edge != edge) ? std::numeric_limits<T>::infinity() : T(0);   // NaN edge: never exceeded; negative: always
    if (std::isinf(edge)) return std::numeric_limits<T>::infinity();
    const T inf = std::numeric_limits<T>::infinity();
    T x = edge * edge;
    if (std::isinf(x)) x = std::numeric_limits<T>::max();
    // edge * edge is within a few ulps of the threshold: walk down to a value that does not exceed, then up to the first
    // one that does (denormal d2 ranges, where the walk would be long, cannot occur between distinct particles)
    int k = 0;
    for (; k < 4096 && x > T(0) && std::sqrt(x) > edge; ++k) x = std::nextafter(x, T(-1));
    for (; k < 4096 && x < inf && !(std::sqrt(x) > edge); ++k) x = std::nextafter(x, inf);
    if (k >= 4096) converged = false;   // the caller falls back to the direct form (sqrt on the device)
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
    bool thr_ok = true;
    T thr2[VEL_EDGES_INLINE];
    for (int e = 0; e < VEL_EDGES_INLINE; ++e) thr2[e] = (e <= nbins && nbins + 1 <= VEL_EDGES_INLINE) ? sqrt_threshold(((const T*)rbins)[e], thr_ok) : std::numeric_limits<T>::infinity();
    auto run = [&](auto fn) -> int {
        fn.v_i = sets[0].aux.p; fn.v_j = sets[two_sets ? 1 : 0].aux.p; fn.rbins = d_rbins.p;
        fn.inline_edges = (nbins + 1 <= VEL_EDGES_INLINE) ? 1 : 0;
        for (int e = 0; e < VEL_EDGES_INLINE; ++e) fn.thr2[e] = (fn.inline_edges && e <= nbins) ? thr2[e] : std::numeric_limits<T>::infinity();
        fn.hb.nbins = nbins; fn.hb.priv = (nbins <= NB_PRIV_MAX) ? 1 : 0; fn.hb.off = StageTotal<T, true>::value + aux_bytes; fn.hb.g_counts = d_hcount.p; fn.hb.g_sums = d_hsum.p;
        return launch_reduce(fn, (size_t)aux_bytes + hist_smem(nbins, fn.hb.priv != 0, sizeof(T)));
    };
    int lrc;
    if (!thr_ok) lrc = (nbins <= NB_PRIV_MAX) ? run(FVel<T, 0, 1>()) : run(FVel<T, 0, 0>());
    else if (nbins + 1 <= 8) lrc = run(FVel<T, 1, 1>());                       // <= 7 bins: private bins (NB_PRIV_MAX = 16)
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

// clm_map_sum_d_d2 (test functor), clm_map_mindist, clm_neighborlist(+copy).
#include <cmath>
#include <limits>
#include "clm_engine.cuh"

namespace clm {

template <class T> int Engine<T>::map_sum(int flags, void* sd, void* sd2, int64_t* np) {
    for (;;) {
        if (int rc = prepare_map(flags)) return rc;
        FSum<T> fn;
        fn.rc2_lo = std::nextafter(geom.cutoff_sqr, T(0));
        fn.rc2_hi = std::nextafter(geom.cutoff_sqr, std::numeric_limits<T>::infinity());
        if (int rc = launch_reduce(fn, 0)) return rc;
        const int v = build_validate();
        if (v == CLM_RETRY_INTERNAL) continue;
        if (v) return v;
        break;
    }
    if (!(flags & CLM_OUT_DEVICE)) {
        if (int rc = fetch_results()) return rc;
        stats.n_pairs = (int64_t)h_res->c[RC_NPAIRS];
        stats.n_cutoff_band = (int64_t)h_res->c[RC_NBAND];
    }
    if (int rc = store_real(sd, &d_res.p->f[RB_SUM_D], &h_res->f[RB_SUM_D], 1, 1.0, flags)) return rc;
    if (int rc = store_real(sd2, &d_res.p->f[RB_SUM_D2], &h_res->f[RB_SUM_D2], 1, 1.0, flags)) return rc;
    if (int rc = store_i64(np, &d_res.p->c[RC_NPAIRS], &h_res->c[RC_NPAIRS], 1, flags)) return rc;
    return finish_map(flags);
}

template <class T> int Engine<T>::map_mindist(int flags, int64_t* i, int64_t* j, void* d) {
    if (!i || !j || !d) return fail(CLM_ERR_ARGUMENT, "output pointer is NULL");
    for (;;) {
        if (int rc = prepare_map(flags)) return rc;
        CLM_CK(d_minpart.ensure((size_t)n_sm * 16));
        FMin<T> fn;
        fn.partial = d_minpart.p;
        if (int rc = launch_reduce(fn, 0)) return rc;
        const int v = build_validate();
        if (v == CLM_RETRY_INTERNAL) continue;
        if (v) return v;
        break;
    }
    k_min_final<<<1, 256, 0, stream>>>(d_minpart.p, last_grid, d_minres.p);
    CLM_CK(cudaGetLastError());
    stats.launches += 1;
    const int accumulate = (flags & CLM_RESET) ? 0 : 1;
    if (flags & CLM_OUT_DEVICE) {
        k_min_store<T><<<1, 32, 0, stream>>>(d_minres.p, (long long*)i, (long long*)j, (T*)d, accumulate);
        CLM_CK(cudaGetLastError());
        stats.launches += 1;
    } else {
        MinResult* hr = reinterpret_cast<MinResult*>(reinterpret_cast<unsigned char*>(h_res) + sizeof(ResultBlock));
        CLM_CK(cudaMemcpyAsync(hr, d_minres.p, sizeof(MinResult), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
        const T dd = (hr->i != 0) ? std::sqrt((T)hr->d2) : std::numeric_limits<T>::infinity();
        if (!accumulate || dd < *(T*)d) { *i = hr->i; *j = hr->j; *(T*)d = dd; }
    }
    return finish_map(flags);
}

// neighborlist! (API/neighborlist.jl:217-231): size hint from the uniform-density estimate
// (_estimated_n_pairs, internals/neighborlist.jl:43-56), emission, retry with the exact size on overflow
template <class T> int Engine<T>::neighborlist(int flags, int64_t* n_out) {
    if (!n_out) return fail(CLM_ERR_ARGUMENT, "output pointer is NULL");
    if (int rc = prepare_map(flags)) return rc;
    double vol = 1.0;
    {
        const T(*m)[3] = box.in;
        vol = (dim == 3) ? std::fabs((double)(m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0])))
                         : std::fabs((double)(m[0][0] * m[1][1] - m[0][1] * m[1][0]));
    }
    const double sphere = ((dim == 2) ? M_PI : 4.0 * M_PI / 3.0) * std::pow((double)box.cutoff, dim);
    const double nx = (double)sets[0].n, ny = (double)sets[1].n;
    const double est = (two_sets ? nx * ny : nx * (nx - 1) / 2) * sphere / std::max(vol, 1e-300);
    size_t capacity = (size_t)std::min(est * 1.15 + 4096.0, 4.0e10);
    capacity = std::max(capacity, nl.cap / 3);
    for (int attempt = 0; attempt < 3; ++attempt) {
        CLM_CK(nl.ensure(capacity * 3));
        FList<T> fn;
        fn.out = nl.p; fn.capacity = capacity;
        fn.rc2_lo = std::nextafter(geom.cutoff_sqr, T(0));
        fn.rc2_hi = std::nextafter(geom.cutoff_sqr, std::numeric_limits<T>::infinity());
        if (int rc = launch_reduce(fn, (size_t)(SWEEP_THREADS / 32) * LIST_STAGE_BYTES)) return rc;
        {
            const int v = build_validate();
            if (v == CLM_RETRY_INTERNAL) { if (int rc = prepare_map(flags & ~CLM_PROFILE)) return rc; --attempt; continue; }
            if (v) return v;
        }
        if (int rc = fetch_results()) return rc;
        nl_count = (int64_t)h_res->c[RC_NLIST];
        if ((size_t)nl_count <= capacity) break;
        capacity = (size_t)nl_count + 1024;          // overflow: the count is exact, rerun with room for it
        if (int rc = prepare_map(flags & ~CLM_PROFILE)) return rc;
        if (attempt == 2) return fail(CLM_ERR_CAPACITY, "neighbour list did not fit after resizing");
    }
    stats.n_pairs = nl_count;
    stats.n_cutoff_band = (int64_t)h_res->c[RC_NBAND];
    *n_out = nl_count;
    return finish_map(flags);
}

template <class T> int Engine<T>::neighborlist_copy(void* rec, int64_t cap, int on_device) {
    if (cap < nl_count) return fail(CLM_ERR_CAPACITY, "records buffer too small: need " + std::to_string(nl_count));
    if (nl_count == 0) return CLM_OK;
    if (!rec) return fail(CLM_ERR_ARGUMENT, "records pointer is NULL");
    CLM_CK(cudaMemcpyAsync(rec, nl.p, (size_t)nl_count * 24, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    CLM_CK(cudaStreamSynchronize(stream));
    return CLM_OK;
}

template int Engine<float>::map_sum(int, void*, void*, int64_t*);
template int Engine<double>::map_sum(int, void*, void*, int64_t*);
template int Engine<float>::map_mindist(int, int64_t*, int64_t*, void*);
template int Engine<double>::map_mindist(int, int64_t*, int64_t*, void*);
template int Engine<float>::neighborlist(int, int64_t*);
template int Engine<double>::neighborlist(int, int64_t*);
template int Engine<float>::neighborlist_copy(void*, int64_t, int);
template int Engine<double>::neighborlist_copy(void*, int64_t, int);

}  // namespace clm

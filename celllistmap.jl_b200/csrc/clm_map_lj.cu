// clm_map_lj / clm_map_coulomb: energy (+ forces) maps of the catalogue.
#include <cmath>
#include "clm_engine.cuh"

namespace clm {

template <class T> int Engine<T>::map_lj(const void* p, int flags, void* e, void* f) {
    if (!p) return fail(CLM_ERR_ARGUMENT, "LJ parameter pointer is NULL");
    const T* c = (const T*)p;
    double scale = 1.0;
    const bool norm = FLJ<T, true, true>::can_normalise(c[0], c[1]);
    const bool n3 = f && n3_usable();
    const bool n3e = !f && n3_scalar_usable();    // energy only: the same lean sweep without force accumulators
    const bool async = (flags & CLM_ASYNC) != 0;
    if (async) { if (int rc = async_begin(flags)) return rc; }
    for (;;) {   // repeated only when the build's record-capacity estimate was too small (build_validate)
    if (n3 || n3e) n3_request();
    if (int rc = prepare_map(flags)) return rc;
    if (n3) {
        // self-set forces: every pair once, both particles updated (clm_sweep_n3.cuh)
        ForceOut<T> fo;
        if (int rc = forces_begin(f, flags, fo)) return rc;
        const int mode = sweep_mode();
        int rc;
        // energy_out == NULL: the pair loop drops the energy arithmetic (one FADD of ~23 instructions per pair step)
        auto run = [&](auto fn) {
            fn.set(c[0], c[1]);
            const T fsc = norm ? fn.fscale : T(1);
            return (mode == MODE_TRI) ? launch_n3<MODE_TRI>(fn, fo.forces, fo.accumulate, fsc) : launch_n3<MODE_HALF>(fn, fo.forces, fo.accumulate, fsc);
        };
        if (norm) rc = e ? run(N3LJ<T, true, true>()) : run(N3LJ<T, true, false>());
        else rc = e ? run(N3LJ<T, false, true>()) : run(N3LJ<T, false, false>());
        if (rc) return rc;
        scale = 1.0;
    } else if (f) {
        // full-shell sweep: every ordered pair (i real, j any image) adds to f_i only -> one plain store per
        // particle, no atomics, no per-batch force copies; the energy is visited twice in self-set systems
        if (norm) {
            FLJ<T, true, true> fn;
            fn.set(c[0], c[1]);
            if (int rc = forces_begin(f, flags, fn.fo)) return rc;
            if (int rc = launch<MODE_ALL>(fn, 0)) return rc;
        } else {
            FLJ<T, true, false> fn;
            fn.set(c[0], c[1]);
            if (int rc = forces_begin(f, flags, fn.fo)) return rc;
            if (int rc = launch<MODE_ALL>(fn, 0)) return rc;
        }
        scale = two_sets ? 1.0 : 0.5;
    } else if (n3e) {
        // energy only, self-set: the reference's exactly-once pair set with the direct form r6 (c12 r6 - c6), on the lean
        // partner-per-lane sweep (clm_sweep_n3.cuh) -- no force accumulators, so more warps are resident
        N3LJ<T, false, true, false> fn;
        fn.set(c[0], c[1]);
        const int rc = (sweep_mode() == MODE_TRI) ? launch_n3_scalar<MODE_TRI>(fn) : launch_n3_scalar<MODE_HALF>(fn);
        if (rc) return rc;
    } else {
        // energy only: the reference's exactly-once sweep with the direct form (same pair set and arithmetic as the oracle)
        FLJ<T, false, false> fn;
        fn.set(c[0], c[1]);
        std::memset(&fn.fo, 0, sizeof(fn.fo));
        if (int rc = launch_reduce(fn, 0)) return rc;
    }
    if (async) break;   // validated when the next frame is enqueued (or by clm_synchronize)
    const int v = build_validate();
    if (v == CLM_RETRY_INTERNAL) continue;
    if (v) return v;
    break;
    }
    if (async) { if (flags & CLM_PROFILE) { if (int rc = finish_map(flags)) return rc; } return async_end(e, f, f ? (size_t)sets[0].n * dim : 0, scale); }
    if (!(flags & CLM_OUT_DEVICE)) { if (int rc = fetch_results()) return rc; }
    if (int rc = store_real(e, &d_res.p->f[RB_ENERGY], &h_res->f[RB_ENERGY], 1, scale, flags)) return rc;
    if (f) { if (int rc = forces_end(f, flags)) return rc; }
    return finish_map(flags);
}

template <class T> int Engine<T>::map_coulomb(const void* wx, const void* wy, const void* k, int flags, void* e, void* f) {
    if (!wx || !k) return fail(CLM_ERR_ARGUMENT, "weights / k pointer is NULL");
    if (two_sets && !wy) return fail(CLM_ERR_ARGUMENT, "weights of the second set are required for a two-set system");
    const bool n3 = f && n3_usable();
    if (n3) n3_request();
    if (int rc = build()) return rc;   // per-record side arrays are sized by the record count: validated build first
    if (int rc = prepare_map(flags)) return rc;
    const bool dev = (flags & CLM_OUT_DEVICE) != 0;
    if (int rc = gather_aux(0, (const T*)wx, 1, false, dev)) return rc;
    if (two_sets) { if (int rc = gather_aux(1, (const T*)wy, 1, false, dev)) return rc; }
    const T* wi = sets[0].aux.p;
    const T* wj = sets[two_sets ? 1 : 0].aux.p;
    double scale = 1.0;
    if (n3) {
        // self-set forces: every pair once, both particles updated (clm_sweep_n3.cuh)
        ForceOut<T> fo;
        if (int rc = forces_begin(f, flags, fo)) return rc;
        N3Coul<T, true> fn;
        fn.k = *(const T*)k; fn.w_rec = wi; fn.fscale = T(1);
        const int rc = (sweep_mode() == MODE_TRI) ? launch_n3<MODE_TRI>(fn, fo.forces, fo.accumulate, T(1)) : launch_n3<MODE_HALF>(fn, fo.forces, fo.accumulate, T(1));
        if (rc) return rc;
    } else if (f) {
        FCoul<T, true> fn;
        fn.k = *(const T*)k; fn.w_i = wi; fn.w_j = wj;
        if (int rc = forces_begin(f, flags, fn.fo)) return rc;
        if (int rc = launch<MODE_ALL>(fn, (size_t)(SWEEP_THREADS / 32) * StageBytes<T, true>::value)) return rc;
        scale = two_sets ? 1.0 : 0.5;
    } else {
        FCoul<T, false> fn;
        fn.k = *(const T*)k; fn.w_i = wi; fn.w_j = wj;
        std::memset(&fn.fo, 0, sizeof(fn.fo));
        if (int rc = launch_reduce(fn, (size_t)(SWEEP_THREADS / 32) * StageBytes<T, true>::value)) return rc;
    }
    if (!dev) { if (int rc = fetch_results()) return rc; }
    if (int rc = store_real(e, &d_res.p->f[RB_ENERGY], &h_res->f[RB_ENERGY], 1, scale, flags)) return rc;
    if (f) { if (int rc = forces_end(f, flags)) return rc; }
    return finish_map(flags);
}

template int Engine<float>::map_lj(const void*, int, void*, void*);
template int Engine<double>::map_lj(const void*, int, void*, void*);
template int Engine<float>::map_coulomb(const void*, const void*, const void*, int, void*, void*);
template int Engine<double>::map_coulomb(const void*, const void*, const void*, int, void*, void*);

}  // namespace clm

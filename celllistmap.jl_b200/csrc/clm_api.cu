// C ABI of libclm_b200.so (include/clm_b200.h) and the non-map members of Engine<T>:
// box construction, position upload, cell-list build (UpdateCellList! equivalent), result staging.
#include <cstring>
#include <cmath>
#include <limits>
#include <new>
#include "clm_engine.cuh"

namespace clm {

static thread_local std::string g_create_error;

template <class T> int Engine<T>::init(int dim_, int device_) {
    dim = dim_; device = device_; dtype = (sizeof(T) == 4) ? CLM_F32 : CLM_F64;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(CLM_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(CLM_ERR_ARGUMENT, "device ordinal out of range");
    CLM_CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CLM_CK(cudaGetDeviceProperties(&prop, device));
    n_sm = prop.multiProcessorCount;
    CLM_CK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    CLM_CK(cudaEventCreate(&ev0));
    CLM_CK(cudaEventCreate(&ev1));
    CLM_CK(cudaEventCreate(&ev2));
    CLM_CK(cudaEventCreate(&ev3));
    CLM_CK(cudaStreamCreateWithFlags(&pub_stream, cudaStreamNonBlocking));
    CLM_CK(cudaEventCreateWithFlags(&ev_built, cudaEventDisableTiming));
    CLM_CK(cudaEventCreate(&ev_b0));
    CLM_CK(cudaEventCreate(&ev_b1));
    CLM_CK(dscal.ensure(DS_COUNT));
    CLM_CK(d_res.ensure(1));
    CLM_CK(d_minres.ensure(2));
    CLM_CK(cudaHostAlloc((void**)&h_dscal, DS_COUNT * sizeof(int), cudaHostAllocMapped));
    CLM_CK(cudaHostGetDevicePointer((void**)&h_dscal_dev, h_dscal, 0));
    CLM_CK(cudaHostAlloc((void**)&h_ints, READ_INTS_MAX * sizeof(int), cudaHostAllocMapped));
    CLM_CK(cudaHostGetDevicePointer((void**)&h_ints_dev, h_ints, 0));
    CLM_CK(cudaEventCreateWithFlags(&ev_ints, cudaEventDisableTiming));
    CLM_CK(cudaMallocHost((void**)&h_res, sizeof(ResultBlock) + 2 * sizeof(MinResult)));
    std::memset(&stats, 0, sizeof(stats));
    stats.n_sm = n_sm;
    return CLM_OK;
}

template <class T> Engine<T>::~Engine() {
    comm_destroy();
    cudaSetDevice(device);
    if (own_stream) cudaStreamSynchronize(own_stream);
    if (copy_in) { cudaStreamSynchronize(copy_in); cudaStreamDestroy(copy_in); }
    if (copy_out) { cudaStreamSynchronize(copy_out); cudaStreamDestroy(copy_out); }
    for (cudaEvent_t e : {ev_h2d, ev_posfree, ev_done, ev_out[0], ev_out[1]}) if (e) cudaEventDestroy(e);
    d_forces_alt.release(); d_eout.release();
    for (auto& s : sets) { s.pos.release(); s.pos_alt.release(); s.fpos.release(); s.rec.release(); s.rec_n3.release(); s.slot_of.release(); s.fmask.release(); s.cell_start.release(); s.counters.release(); s.aux.release(); }
    dscal.release(); tiles.release(); d_res.release();
    d_hcount.release(); nl.release(); d_hsum.release(); d_rbins.release(); d_forces.release(); d_facc.release(); d_minmax.release(); d_minpart.release(); d_minres.release();
    custom_store_free(custom_store);
    if (h_dscal) cudaFreeHost(h_dscal);
    if (h_ints) cudaFreeHost(h_ints);
    if (ev_ints) cudaEventDestroy(ev_ints);
    if (h_res) cudaFreeHost(h_res);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev2) cudaEventDestroy(ev2);
    if (ev3) cudaEventDestroy(ev3);
    if (ev_built) cudaEventDestroy(ev_built);
    if (ev_b0) cudaEventDestroy(ev_b0);
    if (ev_b1) cudaEventDestroy(ev_b1);
    if (pub_stream) { cudaStreamSynchronize(pub_stream); cudaStreamDestroy(pub_stream); }
    if (own_stream) cudaStreamDestroy(own_stream);
}

// Box(...) dispatch: sides -> orthorhombic matrix (Box.jl:330-335), matrix (Box.jl:191-203), Limits deferred (Box.jl:374-377)
template <class T> int Engine<T>::set_box(int cell_type, const void* uc, int is_matrix, const void* cutoff, int lcell) {
    if (!cutoff) return fail(CLM_ERR_ARGUMENT, "cutoff pointer is NULL");
    if (lcell < 1) return fail(CLM_ERR_ARGUMENT, "lcell must be greater or equal to 1");
    const T rc = *(const T*)cutoff;
    dirty = true;
    if (cell_type == CLM_NONPERIODIC) {
        nonperiodic = true; np_cutoff = rc; np_lcell = lcell; box_set = false; np_have_limits = false;
        return CLM_OK;
    }
    if (cell_type != CLM_ORTHORHOMBIC && cell_type != CLM_TRICLINIC) return fail(CLM_ERR_ARGUMENT, "unknown cell type");
    if (!uc) return fail(CLM_ERR_ARGUMENT, "unit cell pointer is NULL");
    nonperiodic = false;
    T cell[3][3];
    geo::eye(cell);
    const T* c = (const T*)uc;
    if (is_matrix) { for (int col = 0; col < dim; ++col) for (int r = 0; r < dim; ++r) cell[r][col] = c[r + dim * col]; }
    else { for (int k = 0; k < dim; ++k) cell[k][k] = c[k]; }
    if (dim == 2) { cell[2][2] = T(1); }
    const T origin[3] = {T(0), T(0), T(0)};
    box_set = false;
    int rcode = make_box(box, dim, cell_type, cell, rc, lcell, origin, err);
    if (rcode != CLM_OK) return rcode;
    box_set = true;
    return CLM_OK;
}

template <class T> int Engine<T>::get_box(clm_box_info* o) {
    if (!o) return fail(CLM_ERR_ARGUMENT, "output pointer is NULL");
    if (!box_set) return fail(CLM_ERR_STATE, "box not set (non-periodic boxes exist after clm_build)");
    std::memset(o, 0, sizeof(*o));
    o->dim = dim; o->dtype = dtype; o->cell_type = box.cell_type; o->lcell = box.lcell;
    o->cutoff = box.cutoff; o->cutoff_sqr = box.cutoff_sqr;
    for (int k = 0; k < 3; ++k) o->nc[k] = (k < dim) ? box.nc[k] : 1;
    for (int col = 0; col < dim; ++col)
        for (int r = 0; r < dim; ++r) {
            o->input_unit_cell[r + dim * col] = box.in[r][col]; o->aligned_unit_cell[r + dim * col] = box.al[r][col];
            o->rotation[r + dim * col] = box.rot[r][col]; o->inv_rotation[r + dim * col] = box.irot[r][col];
        }
    for (int k = 0; k < dim; ++k) { o->computing_box_min[k] = box.cb_min[k]; o->computing_box_max[k] = box.cb_max[k]; o->cell_size[k] = box.cs[k]; o->origin[k] = box.origin[k]; }
    return CLM_OK;
}

template <class T> int Engine<T>::set_positions(int set, const void* xyz, int64_t n, int on_device) {
    if (set != 0 && set != 1) return fail(CLM_ERR_ARGUMENT, "set must be 0 (x) or 1 (y)");
    if (n < 0) return fail(CLM_ERR_ARGUMENT, "negative particle count");
    if (n > 0 && !xyz) return fail(CLM_ERR_ARGUMENT, "positions pointer is NULL");
    if (n > (int64_t)(TagT<float>::MASK)) return fail(CLM_ERR_UNSUPPORTED, "more than 2^29 particles in one set");
    CLM_CK(cudaSetDevice(device));
    DevSet<T>& s = sets[set];
    CLM_CK(s.pos.ensure((size_t)n * dim));
    if (n) CLM_CK(cudaMemcpyAsync(s.pos.p, xyz, (size_t)n * dim * sizeof(T), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    s.n = n;
    if (set == 1) two_sets = (n > 0) || (xyz != nullptr);   // (NULL, 0) removes the second set; an empty y set gives no pairs
    dirty = true;
    return CLM_OK;
}

// pipelined frames: the coordinates of the NEXT frame are copied from PINNED host memory on the copy-in stream into the
// buffer the frame in flight is not using; the next cell-list build waits for the copy, nothing else does
template <class T> int Engine<T>::set_positions_async(int set, const void* xyz, int64_t n) {
    if (set != 0 && set != 1) return fail(CLM_ERR_ARGUMENT, "set must be 0 (x) or 1 (y)");
    if (n <= 0 || !xyz) return fail(CLM_ERR_ARGUMENT, "clm_set_positions_async needs a non-empty pinned host array");
    if (n > (int64_t)(TagT<float>::MASK)) return fail(CLM_ERR_UNSUPPORTED, "more than 2^29 particles in one set");
    CLM_CK(cudaSetDevice(device));
    if (int rc = pipeline_init()) return rc;
    DevSet<T>& s = sets[set];
    std::swap(s.pos, s.pos_alt);
    CLM_CK(s.pos.ensure((size_t)n * dim));
    // the build that read this buffer two frames ago has finished once the latest enqueued build has (same stream)
    CLM_CK(cudaStreamWaitEvent(copy_in, ev_posfree, 0));
    if (!(dbg & 4)) CLM_CK(cudaMemcpyAsync(s.pos.p, xyz, (size_t)n * dim * sizeof(T), cudaMemcpyHostToDevice, copy_in));
    CLM_CK(cudaEventRecord(ev_h2d, copy_in));
    pending_h2d = true;
    s.n = n;
    if (set == 1) two_sets = true;
    dirty = true;
    return CLM_OK;
}

// foreign particles (slab decomposition): real particles owned by other ranks that lie within the stencil reach of
// this rank's slab.  They are binned (with their periodic images) after the owned ones and never act as particle i.
template <class T> int Engine<T>::set_foreign(int set, const void* xyz, int64_t n, int on_device) {
    if (set != 0 && set != 1) return fail(CLM_ERR_ARGUMENT, "set must be 0 (x) or 1 (y)");
    if (n < 0 || (n > 0 && !xyz)) return fail(CLM_ERR_ARGUMENT, "bad foreign particle array");
    CLM_CK(cudaSetDevice(device));
    DevSet<T>& s = sets[set];
    CLM_CK(s.fpos.ensure((size_t)std::max<int64_t>(n, 1) * dim));
    if (n) CLM_CK(cudaMemcpyAsync(s.fpos.p, xyz, (size_t)n * dim * sizeof(T), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    s.n_foreign = n;
    dirty = true;
    return CLM_OK;
}

// rows of the set's own coordinate array that belong to other ranks (see include/clm_b200.h)
template <class T> int Engine<T>::set_foreign_mask(int set, const uint8_t* mask, int64_t n, int on_device) {
    if (set != 0 && set != 1) return fail(CLM_ERR_ARGUMENT, "set must be 0 (x) or 1 (y)");
    if (n < 0 || (n > 0 && !mask)) return fail(CLM_ERR_ARGUMENT, "bad foreign mask");
    CLM_CK(cudaSetDevice(device));
    DevSet<T>& s = sets[set];
    if (n) {
        CLM_CK(s.fmask.ensure((size_t)n));
        CLM_CK(cudaMemcpyAsync(s.fmask.p, mask, (size_t)n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
        if (!on_device) CLM_CK(cudaStreamSynchronize(stream));   // the caller's array may be pageable and short-lived
    }
    s.n_mask = n;
    dirty = true;
    return CLM_OK;
}

// a few device ints back to the host behind everything enqueued on the handle's stream: a one-warp kernel writes them
// into mapped pinned memory and the host waits for an event.  A cudaMemcpy of the same 16 bytes queues on the device->host
// DMA engine, i.e. behind the force copy-out of the previous pipelined frame (96 MB at 8 M particles per rank: the halo
// exchange's count read-back then took 2.8 ms instead of microseconds, bench_multi.py timeline).
template <class T> int Engine<T>::read_ints(const int32_t* dev, int32_t n, int32_t* host_out) {
    if (n < 0 || n > READ_INTS_MAX || (n > 0 && (!dev || !host_out))) return fail(CLM_ERR_ARGUMENT, "clm_read_ints: 0..64 ints, non-NULL pointers");
    if (n == 0) return CLM_OK;
    CLM_CK(cudaSetDevice(device));
    k_publish_ints<<<1, 64, 0, stream>>>((const int*)dev, n, h_ints_dev);
    CLM_CK(cudaGetLastError());
    CLM_CK(cudaEventRecord(ev_ints, stream));
    CLM_CK(cudaEventSynchronize(ev_ints));
    for (int k = 0; k < n; ++k) host_out[k] = h_ints[k];
    stats.launches += 1;
    return CLM_OK;
}

// reference-cell index along one dimension for arbitrary coordinates, with the arithmetic of the build
template <class T> int Engine<T>::cell_coords(const void* xyz, int64_t n, int on_device, int axis, int32_t* out) {
    if (!box_set) return fail(CLM_ERR_STATE, "clm_set_box must be called first (non-periodic boxes exist after clm_build)");
    if (axis < 0 || axis >= dim) return fail(CLM_ERR_ARGUMENT, "axis out of range");
    if (n <= 0) return CLM_OK;
    if (!xyz || !out) return fail(CLM_ERR_ARGUMENT, "NULL pointer");
    CLM_CK(cudaSetDevice(device));
    GeomT<T> g;
    fill_geom(box, g);
    const T* dx = (const T*)xyz;
    int* dout = (int*)out;
    DBuf<T> tx;
    DBuf<int> to;
    if (!on_device) {
        CLM_CK(tx.ensure((size_t)n * dim));
        CLM_CK(to.ensure((size_t)n));
        CLM_CK(cudaMemcpyAsync(tx.p, xyz, (size_t)n * dim * sizeof(T), cudaMemcpyHostToDevice, stream));
        dx = tx.p; dout = to.p;
    }
    const int nb = (int)((n + 255) / 256);
    if (dim == 3) k_cell_coord<T, 3><<<nb, 256, 0, stream>>>(g, dx, (int)n, axis, dout);
    else k_cell_coord<T, 2><<<nb, 256, 0, stream>>>(g, dx, (int)n, axis, dout);
    CLM_CK(cudaGetLastError());
    stats.launches += 1;
    if (!on_device) {
        CLM_CK(cudaMemcpyAsync(out, to.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
        tx.release(); to.release();
    }
    return CLM_OK;
}

// one-pass face selection for the halo exchange (all pointers on the device, enqueue only)
template <class T> int Engine<T>::select_layers(const void* xyz, int64_t n, int axis, const int32_t* ranges, int merge, void* out_a, void* out_b,
                                                int64_t capacity, int32_t* counts, int32_t* idx_a, int32_t* idx_b) {
    if (!box_set) return fail(CLM_ERR_STATE, "clm_set_box must be called first");
    if (axis < 0 || axis >= dim || !ranges || !counts || (n > 0 && (!xyz || !out_a || !out_b))) return fail(CLM_ERR_ARGUMENT, "bad argument");
    if (n <= 0) return CLM_OK;
    CLM_CK(cudaSetDevice(device));
    GeomT<T> g;
    fill_geom(box, g);
    const int nb = (int)((n + 255) / 256);
    const int4 r = make_int4(ranges[0], ranges[1], ranges[2], ranges[3]);
    const int cap = (int)std::min<int64_t>(capacity, 0x7fffffff);
    if (dim == 3) k_select_layers<T, 3><<<nb, 256, 0, stream>>>(g, (const T*)xyz, (int)n, axis, r, merge, (T*)out_a, (T*)out_b, cap, (int*)counts, (int*)idx_a, (int*)idx_b);
    else k_select_layers<T, 2><<<nb, 256, 0, stream>>>(g, (const T*)xyz, (int)n, axis, r, merge, (T*)out_a, (T*)out_b, cap, (int*)counts, (int*)idx_a, (int*)idx_b);
    CLM_CK(cudaGetLastError());
    stats.launches += 1;
    return CLM_OK;
}


// UpdateCellList! (CellLists.jl:727-927; non-periodic NonPeriodicCells.jl:93-230)
// clm_build: enqueue + validate, repeated when the record-capacity estimate was too small
template <class T> int Engine<T>::build() {
    for (;;) {
        if (int rc = build_enqueue()) return rc;
        const int v = build_validate();
        if (v == CLM_RETRY_INTERNAL) continue;
        return v;
    }
}

// everything the build puts on the stream, ending with the D2H copy of the device scalars (validation flags, record
// counts) and an event; nothing waits on the host
template <class T> int Engine<T>::build_enqueue() {
    if (!dirty) return CLM_OK;
    CLM_CK(cudaSetDevice(device));
    const int nsets = two_sets ? 2 : 1;
    for (int s = 0; s < nsets; ++s)
        if (sets[s].n + sets[s].n_foreign > 0x7fffffffLL / 28 || sets[s].n + sets[s].n_foreign > (int64_t)TagT<float>::MASK) return fail(CLM_ERR_UNSUPPORTED, "too many particles for 32-bit record indices");
    if (pending_h2d) { if (!(dbg & 1)) CLM_CK(cudaStreamWaitEvent(stream, ev_h2d, 0)); pending_h2d = false; }
    CLM_CK(cudaEventRecord(ev_b0, stream));
    if (published) CLM_CK(cudaStreamWaitEvent(stream, ev_built, 0));   // the side-stream publish of the previous build has read the scalars (long ago)
    // device scalars (a kernel, not a copy: see k_dscal_init).  Only the limits pass of a non-periodic system needs them
    // before the counters are zeroed; otherwise the zero fill of set 0 initialises them in passing (one launch less)
    const bool np_reuse = nonperiodic && box_set && np_have_limits && !np_force_limits;
    if (nonperiodic && !np_reuse) {
        k_dscal_init<<<1, 32, 0, stream>>>(dscal.p);
        CLM_CK(cudaGetLastError());
        // Box(limits(x[,y]), cutoff): sides = extent + 2.1*cutoff, origin = minimum coordinates
        // (CellOperations.jl:290-324, Box.jl:38, :374-377); limits by a device reduction, ONE host round trip for both
        // sets and the NaN flags
        T lo[3], hi[3];
        for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<T>::max(); hi[k] = std::numeric_limits<T>::lowest(); }
        int nblk[2] = {0, 0};
        for (int s = 0; s < nsets; ++s) nblk[s] = sets[s].n > 0 ? (int)std::min<int64_t>(1024, (sets[s].n + 255) / 256) : 0;
        CLM_CK(d_minmax.ensure((size_t)std::max(nblk[0] + nblk[1], 1) * 6));
        h_stage.resize((size_t)std::max(nblk[0] + nblk[1], 1) * 6 * sizeof(T));
        for (int s = 0; s < nsets; ++s) {
            if (!nblk[s]) continue;
            T* out = d_minmax.p + (size_t)(s ? nblk[0] : 0) * 6;
            if (dim == 3) k_minmax<T, 3><<<nblk[s], 256, 0, stream>>>(sets[s].pos.p, (int)sets[s].n, out, dscal.p + s * DS_SET_STRIDE);
            else k_minmax<T, 2><<<nblk[s], 256, 0, stream>>>(sets[s].pos.p, (int)sets[s].n, out, dscal.p + s * DS_SET_STRIDE);
            CLM_CK(cudaGetLastError());
            stats.launches += 1;
        }
        if (nblk[0] + nblk[1]) CLM_CK(cudaMemcpyAsync(h_stage.data(), d_minmax.p, (size_t)(nblk[0] + nblk[1]) * 6 * sizeof(T), cudaMemcpyDeviceToHost, stream));
        k_dscal_publish<<<1, 32, 0, stream>>>(dscal.p, h_dscal_dev);
        CLM_CK(cudaGetLastError());
        CLM_CK(cudaStreamSynchronize(stream));
        // a NaN coordinate poisons the limits: report it as the reference does (validation runs first there)
        for (int s = 0; s < nsets; ++s)
            if (h_dscal[s * DS_SET_STRIDE + DS_NAN] != IDX_NONE)
                return fail(CLM_ERR_INVALID_COORDINATES, "Invalid coordinates found (NaN) for particle of index " + std::to_string(h_dscal[s * DS_SET_STRIDE + DS_NAN] + 1) + (s ? " of the second set" : ""));
        const T* h = reinterpret_cast<const T*>(h_stage.data());
        for (int s = 0; s < nsets; ++s) {
            T slo[3] = {T(0), T(0), T(0)}, shi[3] = {T(0), T(0), T(0)};   // empty set: zero limits (_minmax, CellOperations.jl:263-265)
            if (nblk[s]) {
                for (int k = 0; k < 3; ++k) { slo[k] = std::numeric_limits<T>::max(); shi[k] = std::numeric_limits<T>::lowest(); }
                const T* hs = h + (size_t)(s ? nblk[0] : 0) * 6;
                for (int b = 0; b < nblk[s]; ++b)
                    for (int k = 0; k < dim; ++k) { slo[k] = std::min(slo[k], hs[(size_t)b * 6 + k]); shi[k] = std::max(shi[k], hs[(size_t)b * 6 + 3 + k]); }
            }
            for (int k = 0; k < dim; ++k) { lo[k] = std::min(lo[k], slo[k]); hi[k] = std::max(hi[k], shi[k]); }
        }
        T cell[3][3], origin[3] = {T(0), T(0), T(0)};
        geo::eye(cell);
        const T pad = T(210) * np_cutoff / T(100);
        for (int k = 0; k < dim; ++k) { cell[k][k] = (hi[k] - lo[k]) + pad; origin[k] = lo[k]; }
        box_set = false;
        int rcode = make_box(box, dim, CLM_NONPERIODIC, cell, np_cutoff, np_lcell, origin, err);
        if (rcode != CLM_OK) return rcode;
        box_set = true;
        for (int k = 0; k < 3; ++k) { np_lo[k] = (k < dim) ? lo[k] : T(0); np_hi[k] = (k < dim) ? hi[k] : T(0); }
        np_have_limits = true;
        np_force_limits = false;
        // (the limits pass used the flags: the zero fill of set 0 below resets them)
    }
    if (!box_set) return fail(CLM_ERR_STATE, "clm_set_box must be called before clm_build");
    fill_geom(box, geom);
    if (np_reuse) {   // the box of the previous build is kept while every particle stays inside the limits it was made from
        geom.np_check = 1;
        for (int k = 0; k < 3; ++k) { geom.np_lo[k] = np_lo[k]; geom.np_hi[k] = np_hi[k]; }
    }
    double nref_total = 1;
    for (int k = 0; k < dim; ++k) nref_total *= (double)box.nc[k];
    if (nref_total > 2.0e9) return fail(CLM_ERR_UNSUPPORTED, "more than 2e9 computing cells: increase the cutoff or use lcell = 1");
    if (box.lcell > LF_MAX) return fail(CLM_ERR_UNSUPPORTED, "lcell > 15 is not supported by the device stencil table");
    nref = (int64_t)nref_total;
    // device grid: split every reference cell into sub^N sub-cells, aiming at ~4 particles per device cell
    {
        double inner = 1;
        for (int k = 0; k < dim; ++k) inner *= (double)std::max<int64_t>(1, box.nc[k] - 2 * box.lcell - 1);
        const double per_cell = (double)std::max(sets[0].n, two_sets ? sets[1].n : (int64_t)0) / inner;
        int sub = opt_sub > 0 ? opt_sub : (int)std::floor(std::pow(std::max(per_cell / 4.0, 1.0), 1.0 / dim) + 0.35);
        sub = std::max(1, std::min(sub, SUB_MAX / box.lcell));
        while (sub > 1) {
            double nd = nref_total * std::pow((double)sub, dim);
            bool ok = nd <= 4.0e8;
            for (int k = 0; k < dim; ++k) ok = ok && (box.nc[k] * sub <= 32767);
            if (ok) break;
            --sub;
        }
        for (int k = 0; k < dim; ++k) if (box.nc[k] * sub > 32767) return fail(CLM_ERR_UNSUPPORTED, "more than 32767 cells along one dimension");
        geom.sub = sub;
    }
    const int sub = geom.sub, lf = box.lcell * sub;
    // device row = cells along the LAST reference dimension (see cell_of)
    nfast = (int)box.nc[dim - 1] * sub; nmid = (int)((dim == 3) ? box.nc[1] : box.nc[0]) * sub; nslow = (int)((dim == 3) ? box.nc[0] * sub : 1);
    ncells = (int64_t)nfast * nmid * nslow;
    nrows = ncells / nfast;
    const int64_t ncp = nrows * (nfast + 1);   // per-cell arrays have a row pitch of nfast + 1 (clm_build.cuh)
    if (ncp + 2 > 0x7fffffff) return fail(CLM_ERR_UNSUPPORTED, "device cell grid too large");
    // stencil rows: a row offset (dslow, dmid) is at least d_perp away; partners can only sit within
    // sqrt(cutoff^2 - d_perp^2) along the row (1e-4 relative slack covers coordinate rounding at cell borders)
    {
        const double csf = (double)box.cs[dim - 1] / sub, csm = (double)((dim == 3) ? box.cs[1] : box.cs[0]) / sub, css = (double)box.cs[0] / sub;
        const double rc2 = (double)box.cutoff * (double)box.cutoff * (1.0 + 2.0e-4);
        std::memset(row_hw, -1, sizeof(row_hw));
        for (int ds = -lf; ds <= lf; ++ds)
            for (int dm = -lf; dm <= lf; ++dm) {
                if (dim == 2 && ds != 0) continue;
                const double gm = std::max(std::abs(dm) - 1, 0) * csm, gs = (dim == 3) ? std::max(std::abs(ds) - 1, 0) * css : 0.0;
                const double rest = rc2 - gm * gm - gs * gs;
                if (rest <= 0) continue;
                const int w = std::min(lf, (int)std::ceil(std::sqrt(rest) / csf));
                row_hw[(ds + lf) * (2 * lf + 1) + dm + lf] = (signed char)w;
            }
    }
    // tile size: particles per warp tile (the remaining lanes split the partners into j-slices)
    tile_i = TILE_I;
    // record capacity without a host round trip: real + image particles ~ n * volume(computing box) / volume(cell)
    // (uniform density) with 25 % slack; a denser boundary layer is detected at the end-of-build sync and the build
    // is repeated once with the exact size
    double img_factor = 1.0;
    if (box.cell_type != CLM_NONPERIODIC) {
        double vbox = 1.0, vcell;
        for (int k = 0; k < dim; ++k) vbox *= (double)box.cb_max[k] - (double)box.cb_min[k];
        const T(*m)[3] = box.al;
        vcell = (dim == 3) ? std::fabs((double)m[0][0] * ((double)m[1][1] * m[2][2] - (double)m[1][2] * m[2][1]) - (double)m[0][1] * ((double)m[1][0] * m[2][2] - (double)m[1][2] * m[2][0]) +
                                       (double)m[0][2] * ((double)m[1][0] * m[2][1] - (double)m[1][1] * m[2][0]))
                           : std::fabs((double)m[0][0] * m[1][1] - (double)m[0][1] * m[1][0]);
        img_factor = std::min(std::max(vbox / std::max(vcell, 1e-300), 1.0), (dim == 3) ? 8.0 : 4.0);
    }
    {
        // tile array: sized from the record capacity of set 0 (its capacity is settled first)
        {
            DevSet<T>& S0 = sets[0];
            const size_t want0 = std::max<size_t>((size_t)((double)(S0.n + S0.n_foreign) * img_factor * 1.25) + 4096, (size_t)std::max<int64_t>(S0.n_tot, 1));
            CLM_CK(S0.rec.ensure(want0));
            tiles_upper = (int64_t)(S0.rec.cap / tile_i) + nrows + 1;
            CLM_CK(tiles.ensure((size_t)tiles_upper));
        }
        for (int s = 0; s < nsets; ++s) {
            DevSet<T>& S = sets[s];
            const size_t want = std::max<size_t>((size_t)((double)(S.n + S.n_foreign) * img_factor * 1.25) + 4096, (size_t)std::max<int64_t>(S.n_tot, 1));
            CLM_CK(S.rec.ensure(want));
            const bool n3 = want_n3 && s == 0;
            if (n3) {
                CLM_CK(S.rec_n3.ensure(S.rec.cap));
                const size_t old_cap = d_facc.cap;
                CLM_CK(d_facc.ensure(S.rec.cap * 4));
                if (d_facc.cap != old_cap) CLM_CK(cudaMemsetAsync(d_facc.p, 0, d_facc.cap * sizeof(T), stream));
            }
            CLM_CK(S.cell_start.ensure((size_t)ncp + 2));
            CLM_CK(S.counters.ensure((size_t)(2 * ncp + nref)));
            S.cell_count = S.counters.p; S.cell_nact = S.counters.p + ncp; S.ref_real = S.counters.p + 2 * ncp;
            {   // one zero fill of [cell_count | cell_nact | ref_real] (a kernel: see k_dscal_init)
                const long long nz = 2 * ncp + nref;
                k_zero_ints<<<(int)std::min<long long>((nz / 4 + 255) / 256 + 1, (long long)n_sm * 8), 256, 0, stream>>>(S.counters.p, nz, s == 0 ? dscal.p : nullptr);
                CLM_CK(cudaGetLastError());
            }
            int* ds = dscal.p + s * DS_SET_STRIDE;
            const int64_t nall = S.n + S.n_foreign;
            const int nb = (int)std::min<int64_t>((nall + 255) / 256, bin_blocks_per_sm > 0 ? (int64_t)n_sm * bin_blocks_per_sm : (int64_t)0x7fffffff);   // k_bin strides over the rest
            const int rec_cap = (int)std::min<size_t>(S.rec.cap, 0x7fffffff);
            if (S.n_mask != 0 && S.n_mask != S.n) return fail(CLM_ERR_ARGUMENT, "clm_set_foreign_mask: the mask must have one byte per row of clm_set_positions");
            const uint8_t* fm = S.n_mask ? S.fmask.p : nullptr;
            if (nall > 0) {
                CLM_CK(S.slot_of.ensure((size_t)nall));
                if (dim == 3) k_bin<T, 3, false><<<nb, 256, 0, stream>>>(geom, S.pos.p, S.fpos.p, (int)nall, (int)S.n, S.cell_count, S.cell_nact, S.ref_real, nullptr, nullptr, 0, ds, fm);
                else k_bin<T, 2, false><<<nb, 256, 0, stream>>>(geom, S.pos.p, S.fpos.p, (int)nall, (int)S.n, S.cell_count, S.cell_nact, S.ref_real, nullptr, nullptr, 0, ds, fm);
                CLM_CK(cudaGetLastError());
                stats.launches += 1;
            }
            // per-row starts (written one slot up: the scatter pass uses cell_start[c + 1] as the cursor of cell c, which leaves
            // the row's exclusive starts behind once every record is placed) + the tiles of set 0 + the real-cell count
            {
                const int make_tiles = (s == 0) ? 1 : 0;
                const int nbr = (int)std::max<int64_t>(((int64_t)nrows * 32 + 255) / 256, 1);
                const int tcap = (int)std::min<int64_t>(tiles_upper, 0x7fffffff);
                if (dim == 3) k_rows<3><<<nbr, 256, 0, stream>>>(S.cell_count, S.cell_nact, S.ref_real, S.cell_start.p, nfast, nmid, (int)nrows, sub, (int)box.nc[1], (int)box.nc[2], (int)nref, make_tiles, tile_i, tiles.p, tcap, ds);
                else k_rows<2><<<nbr, 256, 0, stream>>>(S.cell_count, S.cell_nact, S.ref_real, S.cell_start.p, nfast, nmid, (int)nrows, sub, 1, (int)box.nc[1], (int)nref, make_tiles, tile_i, tiles.p, tcap, ds);
                CLM_CK(cudaGetLastError());
                stats.launches += 1;
            }
            if (nall > 0) {
                if (dim == 3) k_bin<T, 3, true><<<nb, 256, 0, stream>>>(geom, S.pos.p, S.fpos.p, (int)nall, (int)S.n, S.cell_start.p + 1, S.cell_nact, S.ref_real, S.rec.p, S.slot_of.p, rec_cap, ds, fm);
                else k_bin<T, 2, true><<<nb, 256, 0, stream>>>(geom, S.pos.p, S.fpos.p, (int)nall, (int)S.n, S.cell_start.p + 1, S.cell_nact, S.ref_real, S.rec.p, S.slot_of.p, rec_cap, ds, fm);
                CLM_CK(cudaGetLastError());
                stats.launches += 1;
                if (n3) {
                    // slot-tagged twin of the records + zeroed accumulator rows: one coalesced pass in record order
                    const int by_index = box.cell_type == CLM_TRICLINIC ? 1 : 0;
                    const int nbt = (int)(((int64_t)rec_cap + 255) / 256);
                    if (dim == 3) k_twin<T, 3><<<nbt, 256, 0, stream>>>(geom, S.rec.p, S.slot_of.p, ds + DS_NTOT, rec_cap, S.rec_n3.p, d_facc.p, by_index);
                    else k_twin<T, 2><<<nbt, 256, 0, stream>>>(geom, S.rec.p, S.slot_of.p, ds + DS_NTOT, rec_cap, S.rec_n3.p, d_facc.p, by_index);
                    CLM_CK(cudaGetLastError());
                    stats.launches += 1;
                }
            }
        }
        CLM_CK(cudaEventRecord(ev_b1, stream));
        if (ev_posfree) CLM_CK(cudaEventRecord(ev_posfree, stream));   // pipelined frames: the coordinate buffer has been read
        // the one host round trip of the build (validation flags, record counts, tile count) is only ENQUEUED here; the
        // caller queues its map kernels behind it and then waits for this event, so the GPU never idles on the host.  The
        // publish kernel runs on a SIDE stream: its write to mapped host memory travels the same PCIe direction as the
        // force copy-out of the previous pipelined frame and completed only behind it, which held up the sweep queued
        // behind the publish on the compute stream by ~0.09 ms per frame (tools/diag_e2e.py)
        CLM_CK(cudaStreamWaitEvent(pub_stream, ev_b1, 0));
        k_dscal_publish<<<1, 32, 0, pub_stream>>>(dscal.p, h_dscal_dev);
        CLM_CK(cudaGetLastError());
        CLM_CK(cudaEventRecord(ev_built, pub_stream));
        published = true;
    }
    validate_pending = true;
    dirty = false;
    have_n3 = want_n3;
    facc_clean = want_n3;   // k_gather zeroed the accumulator rows of the new list
    return CLM_OK;
}

// waits for the build's device scalars and checks them.  CLM_RETRY_INTERNAL: the record capacity was too small (the
// scatter pass dropped records); capacity is grown, the build is marked dirty and the caller must redo build + map.
template <class T> int Engine<T>::build_validate() {
    if (!validate_pending) return CLM_OK;
    validate_pending = false;
    CLM_CK(cudaEventSynchronize(ev_built));
    const int nsets = two_sets ? 2 : 1;
    bool overflow = false;
    bool nofit = nonperiodic && h_dscal[DS_NOFIT] != 0;
    for (int s = 0; s < nsets; ++s) {
        const int* hs = h_dscal + s * DS_SET_STRIDE;
        if (hs[DS_NAN] != IDX_NONE) {
            dirty = true;
            return fail(CLM_ERR_INVALID_COORDINATES, "Invalid coordinates found (NaN) for particle of index " + std::to_string(hs[DS_NAN] + 1) + (s ? " of the second set" : ""));
        }
        if (nofit) continue;   // the reused non-periodic box does not hold the new coordinates: rebuilt below
        if (hs[DS_OOB] != IDX_NONE) {
            dirty = true;
            return fail(CLM_ERR_INVALID_COORDINATES, "Invalid coordinates found: particle of index " + std::to_string(hs[DS_OOB] + 1) + " falls outside the computing grid (non-finite coordinate?)");
        }
        sets[s].n_tot = hs[DS_NTOT];
        if ((size_t)sets[s].n_tot > sets[s].rec.cap) overflow = true;
    }
    if (nofit) {   // _limits_fit_in_box failed: new limits, new box, build again
        dirty = true;
        np_force_limits = true;
        return CLM_RETRY_INTERNAL;
    }
    if (overflow) {
        dirty = true;
        if (++build_retries > 3) return fail(CLM_ERR_CUDA, "cell-list build did not converge on a record capacity");
        return CLM_RETRY_INTERNAL;
    }
    build_retries = 0;
    float ms = 0;
    CLM_CK(cudaEventElapsedTime(&ms, ev_b0, ev_b1));
    stats.build_ms = ms;
    stats.n_cells = ncells;
    stats.n_tiles = h_dscal[DS_NTILES];
    for (int s = 0; s < 2; ++s) {
        stats.n_real[s] = (s < nsets) ? sets[s].n : 0;
        stats.n_total[s] = (s < nsets) ? sets[s].n_tot : 0;
        stats.n_cells_real[s] = (s < nsets) ? h_dscal[s * DS_SET_STRIDE + DS_NCELLS_REAL] : 0;
    }
    return CLM_OK;
}

template <class T> int Engine<T>::prepare_map(int flags) {
    CLM_CK(cudaSetDevice(device));
    // accumulating into caller-owned DEVICE buffers cannot be redone: validate the build before anything is added
    if ((flags & CLM_OUT_DEVICE) && !(flags & CLM_RESET)) { if (int rc = build()) return rc; }
    else if (int rc = build_enqueue()) return rc;
    if (int rc = flush_pending_out(ev_b1)) return rc;   // outputs of the previous pipelined frame: copied out next to this frame's sweep
    profile_sweep = (flags & CLM_PROFILE) != 0;
    if (flags & CLM_PROFILE) CLM_CK(cudaEventRecord(ev0, stream));
    static_assert(sizeof(ResultBlock) % 8 == 0 && sizeof(ResultBlock) / 8 <= 32, "k_map_begin zeroes the result block with one warp");
    k_map_begin<<<1, 32, 0, stream>>>(reinterpret_cast<unsigned long long*>(d_res.p), (int)(sizeof(ResultBlock) / 8), dscal.p + DS_WORK);
    CLM_CK(cudaGetLastError());
    return CLM_OK;
}
template <class T> int Engine<T>::finish_map(int flags) {
    if (flags & CLM_PROFILE) {
        CLM_CK(cudaEventRecord(ev1, stream));
        if (flags & CLM_ASYNC) { profile_pending = true; return CLM_OK; }   // pipelined frame: the times are read by clm_get_stats
        return profile_collect();
    }
    return CLM_OK;
}
template <class T> int Engine<T>::profile_collect() {
    profile_pending = false;
    CLM_CK(cudaEventSynchronize(ev1));
    float ms = 0;
    CLM_CK(cudaEventElapsedTime(&ms, ev0, ev1));
    stats.map_ms = ms;
    CLM_CK(cudaEventElapsedTime(&ms, ev2, ev3));
    stats.sweep_ms = ms;
    return CLM_OK;
}
template <class T> int Engine<T>::fetch_results() {
    CLM_CK(cudaMemcpyAsync(h_res, d_res.p, sizeof(ResultBlock), cudaMemcpyDeviceToHost, stream));
    CLM_CK(cudaStreamSynchronize(stream));
    return CLM_OK;
}

// per-particle auxiliary input (weights / velocities) -> record order of `set`
template <class T> int Engine<T>::gather_aux(int set, const T* aux, int ncomp, bool rotate, bool on_device) {
    DevSet<T>& S = sets[set];
    const size_t out_n = (size_t)std::max<int64_t>(S.n_tot, 1) * 4;   // one record-sized slot (4 x T) per record
    // owned particles first, then the foreign (halo) particles of a slab-decomposed system: one row each
    const size_t in_n = (size_t)(S.n + S.n_foreign) * ncomp;
    CLM_CK(S.aux.ensure(out_n + in_n + 4));
    T* staged = S.aux.p + out_n;
    if (S.n + S.n_foreign == 0) return CLM_OK;
    CLM_CK(cudaMemcpyAsync(staged, aux, in_n * sizeof(T), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    const int nb = (int)((S.n_tot + 255) / 256);
    if (nb) k_gather_aux<T><<<nb, 256, 0, stream>>>(S.rec.p, (int)S.n_tot, staged, ncomp, geom, rotate ? 1 : 0, S.aux.p);
    CLM_CK(cudaGetLastError());
    stats.launches += 1;
    return CLM_OK;
}

// scalar / small-vector real outputs: out = (reset ? 0 : out) + scale * accumulator
template <class T> int Engine<T>::store_real(void* out, const double* dev_src, const double* host_src, int n, double scale, int flags) {
    if (!out) return CLM_OK;
    if (flags & CLM_OUT_DEVICE) {
        k_store_real<T><<<(n + 127) / 128, 128, 0, stream>>>((T*)out, dev_src, n, scale, (flags & CLM_RESET) ? 0 : 1);
        CLM_CK(cudaGetLastError());
        stats.launches += 1;
    } else {
        T* o = (T*)out;
        for (int k = 0; k < n; ++k) o[k] = (T)(((flags & CLM_RESET) ? 0.0 : (double)o[k]) + scale * host_src[k]);
    }
    return CLM_OK;
}
template <class T> int Engine<T>::store_i64(int64_t* out, const unsigned long long* dev_src, const unsigned long long* host_src, int n, int flags, int shift) {
    if (!out) return CLM_OK;
    if (flags & CLM_OUT_DEVICE) {
        k_store_i64<<<(n + 127) / 128, 128, 0, stream>>>((long long*)out, dev_src, n, (flags & CLM_RESET) ? 0 : 1, shift);
        CLM_CK(cudaGetLastError());
        stats.launches += 1;
    } else {
        for (int k = 0; k < n; ++k) out[k] = ((flags & CLM_RESET) ? 0 : out[k]) + (int64_t)(host_src[k] >> shift);
    }
    return CLM_OK;
}

// forces: the sweep stores each real particle's force exactly once.  Device outputs are written (or
// accumulated) in place; host outputs go through a device staging buffer and are added on the host.
template <class T> int Engine<T>::forces_begin(void* forces_out, int flags, ForceOut<T>& fo) {
    fo.dim = dim; fo.rotated = geom.rotated;
    for (int k = 0; k < 9; ++k) fo.inv_rot[k] = geom.inv_rot[k];
    if (flags & CLM_OUT_DEVICE) { fo.forces = (T*)forces_out; fo.accumulate = (flags & CLM_RESET) ? 0 : 1; return CLM_OK; }
    DBuf<T>& stage = ((flags & CLM_ASYNC) && (frame & 1)) ? d_forces_alt : d_forces;   // pipelined frames: staging by frame parity
    CLM_CK(stage.ensure((size_t)std::max<int64_t>(sets[0].n, 1) * dim));
    fo.forces = stage.p; fo.accumulate = 0;
    return CLM_OK;
}
template <class T> int Engine<T>::forces_end(void* forces_out, int flags) { return part_end(forces_out, flags, dim); }
template <class T> int Engine<T>::part_end(void* forces_out, int flags, int ncomp) {
    if (flags & CLM_OUT_DEVICE) return CLM_OK;
    const size_t cnt = (size_t)sets[0].n * ncomp;
    if (cnt == 0) return CLM_OK;
    if (flags & CLM_RESET) {
        CLM_CK(cudaMemcpyAsync(forces_out, d_forces.p, cnt * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
    } else {
        h_stage.resize(cnt * sizeof(T));
        CLM_CK(cudaMemcpyAsync(h_stage.data(), d_forces.p, cnt * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
        T* o = (T*)forces_out;
        const T* a = (const T*)h_stage.data();
        for (size_t k = 0; k < cnt; ++k) o[k] += a[k];
    }
    return CLM_OK;
}

template <class T> int Engine<T>::get_stats(clm_stats* out) {
    if (!out) return fail(CLM_ERR_ARGUMENT, "output pointer is NULL");
    { const int v = build_validate(); if (v != CLM_OK && v != CLM_RETRY_INTERNAL) return v; }
    if (profile_pending) { if (int rc = profile_collect()) return rc; }
    *out = stats;
    return CLM_OK;
}
template <class T> int Engine<T>::set_option(const char* name, int64_t v) {
    if (!name) return fail(CLM_ERR_ARGUMENT, "option name is NULL");
    const std::string s(name);
    if (s == "sub") { if (v < 0 || v > SUB_MAX) return fail(CLM_ERR_ARGUMENT, "sub must be in 0..7"); opt_sub = (int)v; dirty = true; return CLM_OK; }
    if (s == "dbg") { dbg = (int)v; return CLM_OK; }
    if (s == "bin_blocks_per_sm") { if (v < 0 || v > 64) return fail(CLM_ERR_ARGUMENT, "bin_blocks_per_sm must be in 0..64"); bin_blocks_per_sm = (int)v; return CLM_OK; }
    if (s == "n3") { opt_n3 = (v < 0) ? -1 : (v ? 1 : 0); return CLM_OK; }
    if (s == "blocks_per_sm") { if (v < -8) return fail(CLM_ERR_ARGUMENT, "blocks_per_sm must be >= -8"); opt_bps = (int)v; return CLM_OK; }   // > 0: cap of resident sweep CTAs per SM; < 0: that many below the occupancy maximum; 0: maximum
    return fail(CLM_ERR_ARGUMENT, "unknown option " + s);
}

template struct Engine<float>;
template struct Engine<double>;

}  // namespace clm

// ---- register-resident FMA microbenchmark: the measured FP32 / FP64 SIMT roofline denominator -------
// (MEASURED_PEAKS.json has no FP-pipe figure; bench.py measures it with this kernel)
namespace clm {
template <class T> __global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = T(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = v[k] * a + b;   // contracted to one FMA each: 8 independent chains
    }
    T s = T(0);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == T(-1)) out[0] = s;   // never true: keeps the chains alive
}
template <class T> static int fma_peak(int device, double* tflops) {
    if (cudaSetDevice(device) != cudaSuccess) return CLM_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CLM_ERR_CUDA;
    T* d = nullptr;
    if (cudaMalloc((void**)&d, 64) != cudaSuccess) return CLM_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, iters = (sizeof(T) == 4) ? 16384 : 2048;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_fma_peak<T><<<blocks, 256>>>(d, iters, T(0.999), T(0.001));
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return CLM_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * (double)iters * blocks * 256.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return CLM_OK;
}
}  // namespace clm

// ===================================================================================================
using clm::EngineBase;
struct clm_handle { EngineBase* e; };

extern "C" {
int clm_version(void) { return 100; }
int clm_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) { clm::g_create_error = "clm_host_register: NULL pointer or empty range"; return CLM_ERR_ARGUMENT; }
    const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { (void)cudaGetLastError(); clm::g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return CLM_ERR_CUDA; }
    return CLM_OK;
}
int clm_host_unregister(void* ptr) {
    if (!ptr) { clm::g_create_error = "clm_host_unregister: NULL pointer"; return CLM_ERR_ARGUMENT; }
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); clm::g_create_error = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); return CLM_ERR_CUDA; }
    return CLM_OK;
}
int clm_measure_fma_peak(int device, int dtype, double* tflops) {
    if (!tflops) return CLM_ERR_ARGUMENT;
    return dtype == CLM_F32 ? clm::fma_peak<float>(device, tflops) : clm::fma_peak<double>(device, tflops);
}
const char* clm_last_error(clm_handle* h) { return h ? h->e->err.c_str() : clm::g_create_error.c_str(); }

int clm_create(clm_handle** out, int dim, int dtype, int device, int ngpus) {
    if (!out) { clm::g_create_error = "handle output pointer is NULL"; return CLM_ERR_ARGUMENT; }
    *out = nullptr;
    if (dim != 2 && dim != 3) { clm::g_create_error = "Dimension must be 2 or 3."; return CLM_ERR_DIMENSION; }
    if (dtype != CLM_F32 && dtype != CLM_F64) { clm::g_create_error = "dtype must be CLM_F32 or CLM_F64"; return CLM_ERR_ARGUMENT; }
    if (ngpus != 1) { clm::g_create_error = "one handle drives one GPU; multi-GPU slabs use one handle per rank (clm_comm_init + clm_slab_update)"; return CLM_ERR_UNSUPPORTED; }
    EngineBase* e = nullptr;
    int rc;
    if (dtype == CLM_F32) { auto* p = new (std::nothrow) clm::Engine<float>(); e = p; rc = p ? p->init(dim, device) : CLM_ERR_CUDA; }
    else { auto* p = new (std::nothrow) clm::Engine<double>(); e = p; rc = p ? p->init(dim, device) : CLM_ERR_CUDA; }
    if (rc != CLM_OK) { clm::g_create_error = e ? e->err : "out of memory"; delete e; return rc; }
    *out = new clm_handle{e};
    return CLM_OK;
}
int clm_destroy(clm_handle* h) { if (h) { delete h->e; delete h; } return CLM_OK; }
#define H_OR_FAIL if (!h) return CLM_ERR_ARGUMENT
int clm_set_stream(clm_handle* h, void* s) { H_OR_FAIL; return h->e->set_stream(s); }
int clm_synchronize(clm_handle* h) { H_OR_FAIL; return h->e->synchronize(); }
int clm_set_box(clm_handle* h, int ct, const void* uc, int is_matrix, const void* cutoff, int lcell) { H_OR_FAIL; return h->e->set_box(ct, uc, is_matrix, cutoff, lcell); }
int clm_get_box(clm_handle* h, clm_box_info* o) { H_OR_FAIL; return h->e->get_box(o); }
int clm_set_positions(clm_handle* h, int set, const void* xyz, int64_t n, int on_device) { H_OR_FAIL; return h->e->set_positions(set, xyz, n, on_device); }
int clm_set_positions_async(clm_handle* h, int set, const void* xyz, int64_t n) { H_OR_FAIL; return h->e->set_positions_async(set, xyz, n); }
int clm_build(clm_handle* h) { H_OR_FAIL; return h->e->build(); }
int clm_set_foreign(clm_handle* h, int set, const void* xyz, int64_t n, int on_device) { H_OR_FAIL; return h->e->set_foreign(set, xyz, n, on_device); }
int clm_set_foreign_mask(clm_handle* h, int set, const uint8_t* mask, int64_t n, int on_device) { H_OR_FAIL; return h->e->set_foreign_mask(set, mask, n, on_device); }
int clm_read_ints(clm_handle* h, const int32_t* dev, int32_t n, int32_t* host_out) { H_OR_FAIL; return h->e->read_ints(dev, n, host_out); }
int clm_cell_coords(clm_handle* h, const void* xyz, int64_t n, int on_device, int axis, int32_t* out) { H_OR_FAIL; return h->e->cell_coords(xyz, n, on_device, axis, out); }
int clm_select_layers(clm_handle* h, const void* xyz, int64_t n, int axis, const int32_t* ranges, int merge, void* out_a, void* out_b, int64_t capacity, int32_t* counts, int32_t* idx_a, int32_t* idx_b) { H_OR_FAIL; return h->e->select_layers(xyz, n, axis, ranges, merge, out_a, out_b, capacity, counts, idx_a, idx_b); }
int clm_map_lj(clm_handle* h, const void* p, int flags, void* e, void* f) { H_OR_FAIL; return h->e->map_lj(p, flags, e, f); }
int clm_map_coulomb(clm_handle* h, const void* wx, const void* wy, const void* k, int flags, void* e, void* f) { H_OR_FAIL; return h->e->map_coulomb(wx, wy, k, flags, e, f); }
int clm_map_dist_hist(clm_handle* h, const void* width, int nbins, int flags, int64_t* counts) { H_OR_FAIL; return h->e->map_dist_hist(width, nbins, flags, counts); }
int clm_map_pairvel(clm_handle* h, const void* vx, const void* vy, const void* rbins, int nbins, int flags, int64_t* counts, void* sums) { H_OR_FAIL; return h->e->map_pairvel(vx, vy, rbins, nbins, flags, counts, sums); }
int clm_map_mindist(clm_handle* h, int flags, int64_t* i, int64_t* j, void* d) { H_OR_FAIL; return h->e->map_mindist(flags, i, j, d); }
int clm_map_sum_d_d2(clm_handle* h, int flags, void* sd, void* sd2, int64_t* np) { H_OR_FAIL; return h->e->map_sum(flags, sd, sd2, np); }
int clm_neighborlist(clm_handle* h, int flags, int64_t* n) { H_OR_FAIL; return h->e->neighborlist(flags, n); }
int clm_neighborlist_copy(clm_handle* h, void* rec, int64_t cap, int on_device) { H_OR_FAIL; return h->e->neighborlist_copy(rec, cap, on_device); }
int clm_get_stats(clm_handle* h, clm_stats* o) { H_OR_FAIL; return h->e->get_stats(o); }
int clm_set_option(clm_handle* h, const char* name, int64_t v) { H_OR_FAIL; return h->e->set_option(name, v); }
int clm_custom_compile(clm_handle* h, const char* source, const char* name, int32_t* id, clm_custom_info* info) { H_OR_FAIL; return h->e->custom_compile(source, name, id, info); }
const char* clm_custom_log(clm_handle* h) { return h ? h->e->custom_log() : ""; }
int clm_map_custom(clm_handle* h, int32_t id, const void* params, int nparams, const void* aux_x, const void* aux_y, int nbins, int flags,
                   void* scalars_out, void* part_out, int64_t* hist_counts, void* hist_sums) {
    H_OR_FAIL;
    return h->e->map_custom(id, params, nparams, aux_x, aux_y, nbins, flags, scalars_out, part_out, hist_counts, hist_sums);
}
int clm_comm_unique_id(void* id128) { std::string e; const int rc = clm::comm_unique_id(id128, e); if (rc) clm::g_create_error = e; return rc; }
int clm_comm_init(clm_handle* h, const void* id128, int rank, int world) { H_OR_FAIL; return h->e->comm_init(id128, rank, world); }
int clm_comm_destroy(clm_handle* h) { H_OR_FAIL; return h->e->comm_destroy(); }
int clm_slab_range(clm_handle* h, int32_t* lo, int32_t* hi) { H_OR_FAIL; return h->e->slab_range(lo, hi); }
int clm_slab_update(clm_handle* h, const void* xyz, int64_t n, int on_device) { H_OR_FAIL; return h->e->slab_update(xyz, n, on_device); }
int clm_comm_allreduce_sum(clm_handle* h, void* buf, int64_t count, int kind, int on_device) { H_OR_FAIL; return h->e->comm_allreduce_sum(buf, count, kind, on_device); }
int clm_slab_info(clm_handle* h, int64_t* n_owned, int64_t* n_foreign, int32_t* rank, int32_t* world) { H_OR_FAIL; return h->e->slab_info(n_owned, n_foreign, rank, world); }
int clm_custom_check(const char* source, const char* name, int dtype, char* log, int64_t log_capacity) { return clm::custom_check(source, name, dtype, log, log_capacity); }
}

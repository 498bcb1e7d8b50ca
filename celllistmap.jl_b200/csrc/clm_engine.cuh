// Engine<T>: one particle system on one B200 (the object behind a clm_handle).
// Host-side orchestration only: buffers, launch configuration, result staging.  No CPU compute path.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <type_traits>
#include "clm_common.cuh"
#include "clm_geometry.hpp"
#include "clm_build.cuh"
#include "clm_sweep.cuh"
#include "clm_sweep_n3.cuh"

namespace clm {

constexpr int CLM_RETRY_INTERNAL = -1;   // never crosses the ABI: build + map must be repeated with the grown capacity
constexpr int DS_SET_STRIDE = DS_SET_STRIDE_DEV;   // dscal block of set y starts at dscal + 6

struct EngineBase {
    std::string err;
    int dim = 3, dtype = 0, device = 0, n_sm = 148;
    clm_stats stats{};
    virtual ~EngineBase() {}
    int fail(int code, const std::string& m) { err = m; return code; }
    virtual int set_stream(void* s) = 0;
    virtual int synchronize() = 0;
    virtual int set_box(int cell_type, const void* uc, int is_matrix, const void* cutoff, int lcell) = 0;
    virtual int get_box(clm_box_info* out) = 0;
    virtual int set_positions(int set, const void* xyz, int64_t n, int on_device) = 0;
    virtual int set_positions_async(int set, const void* xyz, int64_t n) = 0;
    virtual int set_foreign(int set, const void* xyz, int64_t n, int on_device) = 0;
    virtual int set_foreign_mask(int set, const uint8_t* mask, int64_t n, int on_device) = 0;
    virtual int read_ints(const int32_t* dev, int32_t n, int32_t* host_out) = 0;
    virtual int cell_coords(const void* xyz, int64_t n, int on_device, int axis, int32_t* out) = 0;
    virtual int comm_init(const void* id, int rank, int world) = 0;
    virtual int comm_destroy() = 0;
    virtual int slab_range(int32_t* lo, int32_t* hi) = 0;
    virtual int slab_update(const void* xyz, int64_t n, int on_device) = 0;
    virtual int comm_allreduce_sum(void* buf, int64_t count, int kind, int on_device) = 0;
    virtual int slab_info(int64_t* n_owned, int64_t* n_foreign, int32_t* rank, int32_t* world) = 0;
    virtual int select_layers(const void* xyz, int64_t n, int axis, const int32_t* ranges, int merge, void* out_a, void* out_b, int64_t capacity, int32_t* counts, int32_t* idx_a, int32_t* idx_b) = 0;
    virtual int build() = 0;
    virtual int map_lj(const void* p, int flags, void* e, void* f) = 0;
    virtual int map_coulomb(const void* wx, const void* wy, const void* k, int flags, void* e, void* f) = 0;
    virtual int map_dist_hist(const void* width, int nbins, int flags, int64_t* counts) = 0;
    virtual int map_pairvel(const void* vx, const void* vy, const void* rbins, int nbins, int flags, int64_t* counts, void* sums) = 0;
    virtual int map_mindist(int flags, int64_t* i, int64_t* j, void* d) = 0;
    virtual int map_sum(int flags, void* sd, void* sd2, int64_t* np) = 0;
    virtual int neighborlist(int flags, int64_t* n) = 0;
    virtual int neighborlist_copy(void* rec, int64_t cap, int on_device) = 0;
    virtual int get_stats(clm_stats* out) = 0;
    virtual int set_option(const char* name, int64_t v) = 0;
    virtual int custom_compile(const char* source, const char* name, int32_t* id_out, clm_custom_info* info) = 0;
    virtual const char* custom_log() = 0;
    virtual int map_custom(int32_t id, const void* params, int nparams, const void* aux_x, const void* aux_y, int nbins, int flags,
                           void* scalars_out, void* part_out, int64_t* hist_counts, void* hist_sums) = 0;
};
void custom_store_free(void* store);   // clm_rtc.cu
int custom_check(const char* source, const char* name, int dtype, char* log_out, int64_t log_cap);

#define CLM_CK(call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return this->fail(CLM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <class U> struct DBuf {
    U* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 4);
        U* q = nullptr;
        cudaError_t e = cudaMalloc((void**)&q, std::max<size_t>(ncap, 1) * sizeof(U));
        if (e != cudaSuccess) return e;
        if (keep && p && cap) cudaMemcpyAsync(q, p, cap * sizeof(U), cudaMemcpyDeviceToDevice, st);
        if (p) { cudaStreamSynchronize(st); cudaFree(p); }
        p = q; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T> struct DevSet {
    DBuf<T> pos;             // caller's coordinates, AoS n x dim (owning copy, like ParticleSystemPositions)
    DBuf<RecT<T>> rec_n3;    // slot-tagged twin of rec for the Newton's-third-law force sweep (k_twin), written on request
    DBuf<int> slot_of;       // particle -> slot of its real record (scatter pass of the build)
    DBuf<T> pos_alt;         // pipelined frames: the buffer the NEXT frame's coordinates are copied into while this one is binned
    int64_t n = 0;
    DBuf<T> fpos;            // foreign particles of a slab-decomposed system (owned by other ranks), AoS
    int64_t n_foreign = 0;
    DBuf<uint8_t> fmask;     // rows of pos that belong to other ranks (clm_set_foreign_mask: local numbering = global order)
    int64_t n_mask = 0;      // 0: no mask
    DBuf<RecT<T>> rec;       // cell-sorted records, real + image particles
    int64_t n_tot = 0, n_cells_real = 0;
    DBuf<int> cell_start;    // row pitch nfast + 1: [row * pitch + x] = first record of cell x, entry nfast = end of the row (after the scatter pass)
    DBuf<int> counters;      // one memset: [cell_count | cell_nact | ref_real]
    int *cell_count = nullptr, *cell_nact = nullptr, *ref_real = nullptr;
    DBuf<T> aux;             // per-record auxiliary data gathered for the map in flight
};

template <class T> struct Engine : EngineBase {
    cudaStream_t stream = nullptr, own_stream = nullptr, pub_stream = nullptr;
    bool published = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev_built = nullptr, ev_b0 = nullptr, ev_b1 = nullptr;
    bool validate_pending = false;
    int build_retries = 0;
    bool profile_sweep = false, profile_pending = false;
    int profile_collect();
    HostBox<T> box;
    GeomT<T> geom;
    bool box_set = false, nonperiodic = false, two_sets = false, dirty = true;
    T np_cutoff = 0;
    int np_lcell = 1;
    // non-periodic: the limits the current box was made from; the next build reuses the box while every particle stays inside
    // them (checked on the device, no host round trip) -- the reference's _limits_fit_in_box
    bool np_have_limits = false, np_force_limits = false;
    T np_lo[3] = {T(0), T(0), T(0)}, np_hi[3] = {T(0), T(0), T(0)};
    DevSet<T> sets[2];
    DBuf<int> dscal;
    DBuf<Tile> tiles;
    int* h_dscal = nullptr;          // pinned, mapped: written by k_dscal_publish
    int* h_ints = nullptr; int* h_ints_dev = nullptr; cudaEvent_t ev_ints = nullptr;   // clm_read_ints: mapped pinned scratch (READ_INTS_MAX ints)
    static constexpr int READ_INTS_MAX = 64;
    int* h_dscal_dev = nullptr;      // its device-side address
    DBuf<ResultBlock> d_res;
    ResultBlock* h_res = nullptr;    // pinned
    DBuf<unsigned long long> d_hcount, nl;
    DBuf<double> d_hsum;
    DBuf<T> d_rbins, d_forces, d_minmax;
    DBuf<MinPartial> d_minpart;
    DBuf<MinResult> d_minres;
    std::vector<unsigned char> h_stage;   // host staging for accumulate-on-host paths
    int64_t nl_count = 0;
    int64_t ncells = 0, nrows = 0, tiles_upper = 0;
    int nfast = 1, nmid = 1, nslow = 1;   // device cell grid: lin = fast + nfast*(mid + nmid*slow); fast = last reference dim
    int64_t nref = 0;                     // cells of the reference grid
    signed char row_hw[(2 * LF_MAX + 1) * (2 * LF_MAX + 1)];   // per stencil row: half-width along the row in device cells, -1 = skip
    int opt_sub = 0;                      // 0 = choose the sub-cell split from the particle density
    int tile_i = TILE_I, opt_bps = 0;
    int bin_blocks_per_sm = 0;    // grid cap of k_bin in blocks per SM (it strides over the particles); clm_set_option "bin_blocks_per_sm", 0 = one block per 256 particles (measured fastest: profiles/r2_tune_bin_grid.txt)
    // self-set force maps: 1 = Newton's-third-law sweep (k_sweep_n3: every pair once, from the same periodic images as the
    // reference, so Float32 forces match the reference's Float32 arithmetic to ~1e-6), 0 = full-shell k_sweep<MODE_ALL>
    // (faster, but a pair that crosses the periodic boundary is evaluated from two different image pairs: 5.6e-5 in
    // Float32 on the 1M-particle C2 system), -1 = by precision: Float32 -> 1, Float64 -> 0 (full shell: 1e-13, 10 % faster)
    int opt_n3 = -1;
    bool facc_clean = false;                 // the accumulator rows are zero (set by the build, cleared by a sweep)
    bool want_n3 = false, have_n3 = false;   // slot-tagged records requested / present in the current cell list
    DBuf<T> d_facc;                       // record-ordered force accumulator rows (4 x T per record slot), all zero between maps

    int init(int dim_, int device_);
    ~Engine() override;
    int set_stream(void* s) override { stream = s ? (cudaStream_t)s : own_stream; return CLM_OK; }
    int synchronize() override {
        if (copy_in) CLM_CK(cudaStreamSynchronize(copy_in));
        if (int rc = flush_pending_out(nullptr)) return rc;
        CLM_CK(cudaStreamSynchronize(stream));
        if (copy_out) { CLM_CK(cudaStreamSynchronize(copy_out)); out_pending[0] = out_pending[1] = false; }
        const int v = build_validate();
        return (v == CLM_RETRY_INTERNAL) ? fail(CLM_ERR_CAPACITY, "the record capacity of the last enqueued cell-list build was too small: repeat the call") : v;
    }
    int set_box(int cell_type, const void* uc, int is_matrix, const void* cutoff, int lcell) override;
    int get_box(clm_box_info* out) override;
    int set_positions(int set, const void* xyz, int64_t n, int on_device) override;
    int set_positions_async(int set, const void* xyz, int64_t n) override;
    int set_foreign(int set, const void* xyz, int64_t n, int on_device) override;
    int set_foreign_mask(int set, const uint8_t* mask, int64_t n, int on_device) override;
    int read_ints(const int32_t* dev, int32_t n, int32_t* host_out) override;
    int cell_coords(const void* xyz, int64_t n, int on_device, int axis, int32_t* out) override;
    int select_layers(const void* xyz, int64_t n, int axis, const int32_t* ranges, int merge, void* out_a, void* out_b, int64_t capacity, int32_t* counts, int32_t* idx_a, int32_t* idx_b) override;
    // slab decomposition over NCCL (clm_comm.cu)
    void* comm_state = nullptr;
    int comm_init(const void* id, int rank, int world) override;
    int comm_destroy() override;
    int slab_range(int32_t* lo, int32_t* hi) override;
    int slab_update(const void* xyz, int64_t n, int on_device) override;
    int comm_allreduce_sum(void* buf, int64_t count, int kind, int on_device) override;
    int slab_info(int64_t* n_owned, int64_t* n_foreign, int32_t* rank, int32_t* world) override;
    int build() override;
    int build_enqueue();
    int build_validate();
    int map_lj(const void* p, int flags, void* e, void* f) override;
    int map_coulomb(const void* wx, const void* wy, const void* k, int flags, void* e, void* f) override;
    int map_dist_hist(const void* width, int nbins, int flags, int64_t* counts) override;
    int map_pairvel(const void* vx, const void* vy, const void* rbins, int nbins, int flags, int64_t* counts, void* sums) override;
    int map_mindist(int flags, int64_t* i, int64_t* j, void* d) override;
    int map_sum(int flags, void* sd, void* sd2, int64_t* np) override;
    int neighborlist(int flags, int64_t* n) override;
    int neighborlist_copy(void* rec, int64_t cap, int on_device) override;
    int get_stats(clm_stats* out) override;
    int set_option(const char* name, int64_t v) override;
    // run-time compiled user pair functions (clm_rtc.cu)
    void* custom_store = nullptr;
    int custom_compile(const char* source, const char* name, int32_t* id_out, clm_custom_info* info) override;
    const char* custom_log() override;
    int map_custom(int32_t id, const void* params, int nparams, const void* aux_x, const void* aux_y, int nbins, int flags,
                   void* scalars_out, void* part_out, int64_t* hist_counts, void* hist_sums) override;

    // ---- helpers shared by the map translation units ----
    int sweep_mode() const { return two_sets ? MODE_ALL : (box.cell_type == CLM_TRICLINIC ? MODE_TRI : MODE_HALF); }
    int prepare_map(int flags);                       // build if needed, zero the result block and the work counter
    int finish_map(int flags);                        // CLM_PROFILE timing
    int fetch_results();                              // result block -> h_res (synchronises)
    int gather_aux(int set, const T* aux_host_or_dev, int ncomp, bool rotate, bool on_device);
    int store_real(void* out, const double* dev_src, const double* host_src, int n, double scale, int flags);
    int store_i64(int64_t* out, const unsigned long long* dev_src, const unsigned long long* host_src, int n, int flags, int shift = 0);   // shift: counts of a full-shell self sweep are halved
    int forces_begin(void* forces_out, int flags, ForceOut<T>& fo);
    int forces_end(void* forces_out, int flags);
    int part_end(void* out, int flags, int ncomp);     // per-particle output of `ncomp` components: staging buffer -> caller's host array

    SweepArgs<T> make_args() const {
        SweepArgs<T> a;
        const DevSet<T>& tg = sets[two_sets ? 1 : 0];
        a.rec_i = sets[0].rec.p; a.rec_j = tg.rec.p; a.cell_start_i = sets[0].cell_start.p; a.cell_start_j = tg.cell_start.p;
        a.tiles = tiles.p; a.dscal = dscal.p; a.res = d_res.p;
        a.nx = nfast; a.ny = nmid; a.nz = nslow; a.lf = geom.lcell * geom.sub; a.sub = geom.sub; a.sub_magic = (unsigned)((0x100000000ull + (unsigned)geom.sub - 1) / (unsigned)geom.sub); a.self = two_sets ? 0 : 1;
        a.rc2 = geom.cutoff_sqr;
        a.rec_cap_i = (int)std::min<size_t>(sets[0].rec.cap, 0x7fffffff); a.rec_cap_j = (int)std::min<size_t>(tg.rec.cap, 0x7fffffff);
        std::memcpy(a.hw, row_hw, sizeof(a.hw));
        {   // stencil row r -> (dslow, dmid)
            const int lf = a.lf, hww = 2 * lf + 1;
            for (int r = 0; r < hww * hww; ++r) { a.rdz[r] = (signed char)((nslow == 1) ? 0 : r / hww - lf); a.rdy[r] = (signed char)((nslow == 1) ? ((r < hww) ? r - lf : 0) : r % hww - lf); }
        }
        return a;
    }
    template <int MODE, class F> int launch(const F& f, size_t functor_smem) {
        auto kern = k_sweep<T, MODE, F>;
        const size_t smem = (size_t)StageTotal<T, F::AUX>::value + functor_smem;   // per-warp staging buffers + mbarriers, then the functor's bins
        static size_t smem_set[64] = {0};                          // per instantiation and device (the attribute is per device)
        size_t& set = smem_set[device & 63];
        if (smem > set) { CLM_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = smem; }
        int bps = 0;
        CLM_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SWEEP_THREADS, smem));
        if (bps < 1) return fail(CLM_ERR_CUDA, "sweep kernel does not fit on an SM");
        if (opt_bps > 0) bps = std::min(bps, opt_bps);
        else if (opt_bps < 0) bps = std::max(1, bps + opt_bps);   // relative: leave room for |opt_bps| CTAs per SM (FramePipeline)
        int64_t grid = (int64_t)n_sm * bps;
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, (tiles_upper + (SWEEP_THREADS / 32) - 1) / (SWEEP_THREADS / 32)));
        if (profile_sweep) CLM_CK(cudaEventRecord(ev2, stream));
        kern<<<(unsigned)grid, SWEEP_THREADS, smem, stream>>>(make_args(), f);
        CLM_CK(cudaGetLastError());
        if (profile_sweep) CLM_CK(cudaEventRecord(ev3, stream));
        stats.launches += 1;
        last_grid = (int)grid;
        return CLM_OK;
    }
    // reduction-type functors run in the reference's own exactly-once mode for the system type
    template <class F> int launch_reduce(const F& f, size_t smem) {
        if (sweep_mode() == MODE_TRI && (sets[0].n_foreign > 0))
            return fail(CLM_ERR_UNSUPPORTED, "slab-decomposed triclinic self-set systems: the index_i < index_j rule needs the global numbering -- pass owned + halo particles as one array sorted by global index and flag the halo rows with clm_set_foreign_mask");
        switch (sweep_mode()) {
            case MODE_HALF: return launch<MODE_HALF>(f, smem);
            case MODE_TRI: return launch<MODE_TRI>(f, smem);
            default: return launch<MODE_ALL>(f, smem);
        }
    }
    // Newton's-third-law force sweep of a self-set system (clm_sweep_n3.cuh): sweep into the record-ordered accumulator,
    // then gather into the caller's particle order
    bool n3_usable() const { return (opt_n3 < 0 ? sizeof(T) == 4 : opt_n3 != 0) && !two_sets && sets[0].n_foreign == 0 && sets[0].n_mask == 0; }
    int n3_request() { want_n3 = true; if (!have_n3) dirty = true; return CLM_OK; }   // the list in place has no slot-tagged records: rebuild
    template <int MODE, class F> int launch_n3(const F& f, T* out, int accumulate, T scale) {
        auto kern = k_sweep_n3<T, MODE, F>;
        const size_t smem = (size_t)N3Smem<T, F::AUX>::value;
        static size_t smem_set[64] = {0};
        size_t& set = smem_set[device & 63];
        if (smem > set) { CLM_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = smem; }
        int bps = 0;
        CLM_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SWEEP_THREADS, smem));
        if (bps < 1) return fail(CLM_ERR_CUDA, "sweep kernel does not fit on an SM");
        if (opt_bps > 0) bps = std::min(bps, opt_bps);
        else if (opt_bps < 0) bps = std::max(1, bps + opt_bps);   // relative: leave room for |opt_bps| CTAs per SM (FramePipeline)
        int64_t grid = (int64_t)n_sm * bps;
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, (tiles_upper + (SWEEP_THREADS / 32) - 1) / (SWEEP_THREADS / 32)));
        DevSet<T>& S = sets[0];
        const int nrows_cap = (int)std::min<size_t>(S.rec.cap, 0x7fffffff);
        if (!facc_clean) {   // a second map on the same cell list: the rows of the previous sweep are cleared by a coalesced fill
            k_zero_rows<T><<<n_sm * 4, 256, 0, stream>>>(d_facc.p, dscal.p, nrows_cap);
            CLM_CK(cudaGetLastError());
            stats.launches += 1;
        }
        facc_clean = false;
        SweepArgs<T> a = make_args();
        a.rec_j = S.rec_n3.p;
        if (profile_sweep) CLM_CK(cudaEventRecord(ev2, stream));
        kern<<<(unsigned)grid, SWEEP_THREADS, smem, stream>>>(a, f, d_facc.p);
        CLM_CK(cudaGetLastError());
        if (profile_sweep) CLM_CK(cudaEventRecord(ev3, stream));
        k_force_finish<T><<<(int)((S.n + 255) / 256), 256, 0, stream>>>((MODE == MODE_TRI) ? nullptr : S.slot_of.p, d_facc.p, dscal.p, (int)std::min<size_t>(S.rec.cap, 0x7fffffff), (int)S.n, out, dim, scale, accumulate, geom.rotated, geom);
        CLM_CK(cudaGetLastError());
        stats.launches += 2;
        last_grid = (int)grid;
        return CLM_OK;
    }
    // energy-only functors (F::FORCES == false) on the same lean sweep: every pair once from the reference's own image, no
    // accumulator rows, nothing to gather afterwards.  Default for Float32 only, like the force maps: in Float64 the FP64 pipe
    // bounds the pair loop and the half shell's looser cull (31 % against 39 % lane hit rate) makes the lean sweep slower
    bool n3_scalar_usable() const { return n3_usable(); }
    template <int MODE, class F> int launch_n3_scalar(const F& f) {
        static_assert(!F::FORCES, "launch_n3_scalar is for functors without force outputs");
        auto kern = k_sweep_n3<T, MODE, F>;
        const size_t smem = (size_t)N3Smem<T, F::AUX, true>::value;
        static size_t smem_set[64] = {0};
        size_t& set = smem_set[device & 63];
        if (smem > set) { CLM_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = smem; }
        int bps = 0;
        CLM_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SWEEP_THREADS, smem));
        if (bps < 1) return fail(CLM_ERR_CUDA, "sweep kernel does not fit on an SM");
        if (opt_bps > 0) bps = std::min(bps, opt_bps);
        else if (opt_bps < 0) bps = std::max(1, bps + opt_bps);
        int64_t grid = (int64_t)n_sm * bps;
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, (tiles_upper + (SWEEP_THREADS / 32) - 1) / (SWEEP_THREADS / 32)));
        SweepArgs<T> a = make_args();
        a.rec_j = sets[0].rec_n3.p;
        if (profile_sweep) CLM_CK(cudaEventRecord(ev2, stream));
        kern<<<(unsigned)grid, SWEEP_THREADS, smem, stream>>>(a, f, (T*)nullptr);
        CLM_CK(cudaGetLastError());
        if (profile_sweep) CLM_CK(cudaEventRecord(ev3, stream));
        stats.launches += 1;
        last_grid = (int)grid;
        return CLM_OK;
    }
    int last_grid = 0;

    // ---- pipelined frames (clm_set_positions_async + CLM_ASYNC maps): independent frames of a trajectory overlap their
    //      host->device copy, compute and device->host copy on three streams; buffers are double-buffered by frame parity
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_posfree = nullptr, ev_done = nullptr, ev_out[2] = {nullptr, nullptr};
    bool pending_h2d = false, out_pending[2] = {false, false};
    int frame = 0;
    int dbg = 0;   // debugging switches of the pipelined path (clm_set_option "dbg"): 1 no wait for the copy-in, 2 no force copy-out, 4 no copy-in, 8 copy out at once
    DBuf<T> d_forces_alt, d_eout;
    int pipeline_init() {
        if (copy_in) return CLM_OK;
        CLM_CK(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
        CLM_CK(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
        CLM_CK(cudaEventCreateWithFlags(&ev_h2d, cudaEventDisableTiming));
        CLM_CK(cudaEventCreateWithFlags(&ev_posfree, cudaEventDisableTiming));
        CLM_CK(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
        CLM_CK(cudaEventCreateWithFlags(&ev_out[0], cudaEventDisableTiming));
        CLM_CK(cudaEventCreateWithFlags(&ev_out[1], cudaEventDisableTiming));
        CLM_CK(d_eout.ensure(2));
        return CLM_OK;
    }
    // a map called with CLM_ASYNC: the previous frame's build is validated here (its scalars arrived long ago), this
    // frame's outputs go to the buffers of its parity once the copy-out of the frame before the previous one has drained
    int async_begin(int flags) {
        if ((flags & CLM_OUT_DEVICE) || !(flags & CLM_RESET)) return fail(CLM_ERR_ARGUMENT, "CLM_ASYNC needs host outputs and CLM_RESET");
        if (int rc = pipeline_init()) return rc;
        const int v = build_validate();
        if (v == CLM_RETRY_INTERNAL) return fail(CLM_ERR_CAPACITY, "the record capacity of the previous asynchronous frame was too small (it has been grown): repeat that frame");
        if (v) return v;
        const int p = frame & 1;
        if (out_pending[p]) CLM_CK(cudaStreamWaitEvent(stream, ev_out[p], 0));
        return CLM_OK;
    }
    // energy (scaled) -> the frame's device scalar; forces staging + scalar -> caller's PINNED host memory on the copy-out stream
    int async_end(void* e_host, void* f_host, size_t force_count, double escale) {
        const int p = frame & 1;
        if (e_host) {
            k_store_real<T><<<1, 32, 0, stream>>>(d_eout.p + p, &d_res.p->f[RB_ENERGY], 1, escale, 0);
            CLM_CK(cudaGetLastError());
            stats.launches += 1;
        }
        CLM_CK(cudaEventRecord(ev_done, stream));
        CLM_CK(cudaStreamWaitEvent(copy_out, ev_done, 0));
        // The copies themselves are enqueued by the NEXT map call, gated on the end of its cell-list build (or by
        // clm_synchronize): a 12 MB device->host transfer that runs next to the build slows it by ~50 % -- the build is a
        // chain of short kernels, and while the copy saturates the upstream PCIe direction every kernel launch waits longer
        // for its commands (tools/diag_e2e.py: build 0.096 -> 0.145 ms, map tail +0.05 ms).  Next to the one long sweep
        // kernel the copy is free.
        pend_e = e_host; pend_f = (f_host && force_count && !(dbg & 2)) ? f_host : nullptr; pend_count = force_count; pend_p = p; pend_valid = true;
        if (dbg & 8) { if (int rc = flush_pending_out(nullptr)) return rc; }   // debugging: copy out at once (the round-2 behaviour)
        frame += 1;
        return CLM_OK;
    }
    void* pend_e = nullptr; void* pend_f = nullptr; size_t pend_count = 0; int pend_p = 0; bool pend_valid = false;
    int flush_pending_out(cudaEvent_t gate) {
        if (!pend_valid) return CLM_OK;
        pend_valid = false;
        const int p = pend_p;
        if (gate) CLM_CK(cudaStreamWaitEvent(copy_out, gate, 0));
        if (pend_f) CLM_CK(cudaMemcpyAsync(pend_f, (p ? d_forces_alt.p : d_forces.p), pend_count * sizeof(T), cudaMemcpyDeviceToHost, copy_out));
        if (pend_e) CLM_CK(cudaMemcpyAsync(pend_e, d_eout.p + p, sizeof(T), cudaMemcpyDeviceToHost, copy_out));
        CLM_CK(cudaEventRecord(ev_out[p], copy_out));
        out_pending[p] = true;
        return CLM_OK;
    }
};

int comm_unique_id(void* out, std::string& err);

}  // namespace clm

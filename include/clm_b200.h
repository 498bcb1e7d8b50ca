/*
 * clm_b200.h -- C ABI of the B200-native cutoff-pair engine (libclm_b200.so).
 *
 * This is the drop-in boundary for the hot path of m3g/CellListMap.jl v0.10.4-DEV.  The
 * reference has no FFI of its own (it is pure Julia): the seam these entry points replace is
 * the pair of internal generic functions reached from the two public entry points
 *
 *   UpdateCellList!(x, [y,] box, cl, aux; parallel, validate_coordinates)
 *        src/internals/CellLists.jl:727-734, :1180-1188; src/internals/NonPeriodicCells.jl:93-100, :253-261
 *        (called from UpdateParticleSystem!, src/internals/ParticleSystem.jl:158-164, :209-224)
 *   _pairwise!(f, output, box, cl; parallel, output_threaded, show_progress)
 *        src/internals/self.jl:28-45, src/internals/cross.jl:8-25
 *        (called from pairwise!, src/API/pairwise.jl:48-63, and neighborlist!, src/API/neighborlist.jl:217-231)
 *
 * A Julia host package binds these with `ccall` (see INTEGRATION.md and
 * celllistmap.jl_b200/julia/CellListMapB200.jl); Python tests bind them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, else a clm_status error class; the message is
 *     available from clm_last_error().  No exception crosses the ABI.
 *   - `const void*` scalars/arrays are of the handle's dtype T (float for CLM_F32, double for
 *     CLM_F64).  Matrices are column-major N x N with COLUMNS = lattice vectors, exactly the
 *     memory of Julia's SMatrix{N,N,T}.  Positions are AoS n x N of T, exactly the memory of
 *     Julia's Vector{SVector{N,T}} (== an (N,n) Matrix{T}).
 *   - particle indices crossing the ABI are 1-based int64 (NeighborPair.i/j, src/API/NeighborPair.jl:19-33).
 *   - pointers are caller-owned HOST memory unless the call's `on_device` / CLM_OUT_DEVICE says
 *     they are device pointers on the handle's device.
 *   - a handle is single-threaded, like one `pairwise!` at a time per ParticleSystem.
 *   - there is NO CPU fallback: without a CUDA device clm_create fails with CLM_ERR_CUDA.
 */
#ifndef CLM_B200_H
#define CLM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLM_API __attribute__((visibility("default")))
#else
#define CLM_API
#endif

typedef struct clm_handle clm_handle;

enum clm_dtype { CLM_F32 = 0, CLM_F64 = 1 };
/* unit-cell types: src/internals/Box.jl:5-9 */
enum clm_cell_type { CLM_ORTHORHOMBIC = 0, CLM_TRICLINIC = 1, CLM_NONPERIODIC = 2 };

enum clm_status {
    CLM_OK = 0,
    CLM_ERR_INVALID_COORDINATES = 1, /* ArgumentError "Invalid coordinates found"  CellOperations.jl:9-17 */
    CLM_ERR_UNIT_CELL = 2,           /* ArgumentError "Unit cell matrix does not satisfy..."  Box.jl:243 */
    CLM_ERR_ARGUMENT = 3,            /* ArgumentError (lcell < 1 Box.jl:192, bad enum, null pointer, ...) */
    CLM_ERR_STATE = 4,               /* call order: box / positions missing */
    CLM_ERR_DIMENSION = 5,           /* DimensionMismatch  CellLists.jl:740-751 */
    CLM_ERR_CAPACITY = 6,            /* caller buffer too small (clm_neighborlist_copy) */
    CLM_ERR_CUDA = 7,                /* ErrorException: CUDA runtime failure, or no device */
    CLM_ERR_COMM = 8,                /* ErrorException: NCCL failure */
    CLM_ERR_UNSUPPORTED = 9
};

/* flags of the map entry points */
enum clm_flags {
    CLM_RESET = 1,      /* reset=true of pairwise! (API/pairwise.jl:52-54): outputs start from zero;
                           without it results are accumulated on the values found in the output buffers */
    CLM_OUT_DEVICE = 2, /* ARRAY arguments of the call (outputs, and per-particle inputs such as weights and
                           velocities) are device pointers; scalars and bin edges stay host pointers.  The call only
                           enqueues work on the handle's stream (no host synchronisation) */
    CLM_PROFILE = 4,    /* record CUDA-event timings of this call into clm_stats */
    CLM_ASYNC = 8       /* pipelined frames (clm_map_lj; needs CLM_RESET and PINNED host outputs): the call only enqueues; the
                           outputs are written by a device->host copy on a separate stream and are valid after
                           clm_synchronize().  With clm_set_positions_async the copy-in of frame k+1, the compute of frame
                           k and the copy-out of frame k-1 overlap (frames of a trajectory are independent).  The copy-out of
                           a frame is issued by the NEXT map call (behind that frame's cell-list build, so that it travels next
                           to the long sweep kernel instead of slowing the build's chain of short kernels) or by
                           clm_synchronize(), whichever comes first: the outputs of frame k are complete after
                           clm_synchronize(), never merely because later frames were enqueued */
};

/* Box record (src/internals/Box.jl:84-96); values widened to double (exact for float). */
typedef struct clm_box_info {
    int32_t dim, dtype, cell_type, lcell;
    int64_t nc[3];                 /* number of computing cells per dimension (Box.jl:209-220) */
    double cutoff, cutoff_sqr;
    double input_unit_cell[9];     /* column-major dim x dim */
    double aligned_unit_cell[9];
    double rotation[9];
    double inv_rotation[9];
    double computing_box_min[3], computing_box_max[3];
    double cell_size[3];
    double origin[3];
} clm_box_info;

typedef struct clm_stats {
    int64_t n_real[2];        /* particles of set x / y */
    int64_t n_total[2];       /* real + image particles in the computing box (CellList.n_particles) */
    int64_t n_cells;          /* prod(nc) */
    int64_t n_cells_real[2];  /* cells containing at least one real particle */
    int64_t n_tiles;          /* work items of the last build (row tiles of the reference set) */
    int64_t n_pairs;          /* in-cutoff pairs of the last map that counts them (sum_d_d2, neighborlist) */
    int64_t n_cutoff_band;    /* pairs with |d2 - cutoff^2| <= 1 ulp(cutoff^2) seen by the last clm_map_sum_d_d2 / clm_neighborlist:
                                 the at-cutoff band of north_star, "reported separately" (docs/src/neighborlists.md:12) */
    double build_ms;          /* device time of the last clm_build (CUDA events) */
    double map_ms;            /* device time of the last map / neighborlist call run with CLM_PROFILE */
    double sweep_ms;          /* device time of the pair-sweep kernel alone inside that call (CUDA events around the launch) */
    int32_t n_sm;             /* SM count of the device */
    int32_t launches;         /* kernels launched by this handle since creation */
} clm_stats;

/* ---- lifetime ---------------------------------------------------------------------------- */
/* dim = 2|3; dtype = clm_dtype; device = CUDA ordinal; ngpus = 1 (multi-GPU slabs: clm_comm_*). */
CLM_API int clm_create(clm_handle** h, int dim, int dtype, int device, int ngpus);
CLM_API int clm_destroy(clm_handle* h);
CLM_API const char* clm_last_error(clm_handle* h); /* h may be NULL: error of the last failed clm_create */
/* run on a caller-provided cudaStream_t (e.g. torch's current stream); NULL = the handle's own stream */
CLM_API int clm_set_stream(clm_handle* h, void* cuda_stream);
CLM_API int clm_synchronize(clm_handle* h);

/* ---- Box(...)  src/internals/Box.jl:191-203, :330-335, :374-377 ; update_box :395-423 ------- */
/* unitcell: N sides (is_matrix = 0) or N x N column-major matrix (is_matrix = 1); ignored for
 * CLM_NONPERIODIC, where the box is derived from limits(x[,y]) at clm_build time. */
CLM_API int clm_set_box(clm_handle* h, int cell_type, const void* unitcell, int is_matrix, const void* cutoff, int lcell);
CLM_API int clm_get_box(clm_handle* h, clm_box_info* out);

/* ---- positions: ParticleSystemPositions copy semantics (API/ParticleSystemPositions.jl:19-22) */
/* set: 0 = x (reference set), 1 = y (target set; giving it makes the system a two-set system);
 * (set = 1, aos_xyz = NULL, n = 0) removes the second set; a non-NULL pointer with n = 0 is an EMPTY
 * second set (no pairs). */
CLM_API int clm_set_positions(clm_handle* h, int set, const void* aos_xyz, int64_t n, int on_device);
/* pipelined frames: the same update from PINNED host memory, enqueued on the handle's copy-in stream into the buffer the
 * frame in flight is not using; returns at once (the caller must not touch the array until the next clm_synchronize,
 * or until two more frames have been enqueued).  The reference's counterpart is the per-frame
 * `sys.xpositions .= frame; pairwise!(f, sys)` loop of a trajectory analysis (docs/src/ParticleSystem/updating.md). */
CLM_API int clm_set_positions_async(clm_handle* h, int set, const void* aos_xyz_pinned, int64_t n);

/* ---- slab decomposition across GPUs (no counterpart in the reference, which is single-process) ----
 * One handle per rank.  A rank owns the particles whose reference cell along dimension 1 falls in its slab and
 * passes them with clm_set_positions; the particles of the neighbouring slabs within the stencil reach (lcell cells)
 * are passed with clm_set_foreign: they are binned (with their periodic images) like any other particle and serve as
 * partners j, but never act as particle i, so every pair evaluation of the single-GPU sweep happens on exactly one rank.
 * Indices: owned particles 1..n, foreign ones n+1..n+n_foreign; per-particle INPUT side arrays of a map (weights,
 * velocities, aux_x / aux_y) then have n + n_foreign rows in that order, per-particle OUTPUTS n rows.  clm_cell_coords returns the 0-based reference-cell
 * index along `axis` of arbitrary coordinates with exactly the arithmetic of the build (ownership / halo selection).
 * clm_set_foreign: orthorhombic and non-periodic cells.  Triclinic self-set systems evaluate a pair from the particle
 * with the smaller index (index_i < index_j, src/internals/self.jl:164-184), so the LOCAL numbering of a rank must follow
 * the global one: there the owned and the halo particles are passed as ONE array with clm_set_positions, sorted by
 * global index, and clm_set_foreign_mask flags the rows that belong to other ranks (mask[k] != 0: partner only, never
 * particle i; NULL removes the mask).  Works for every cell type; the mask is n bytes (n of clm_set_positions). */
CLM_API int clm_set_foreign(clm_handle* h, int set, const void* aos_xyz, int64_t n, int on_device);
CLM_API int clm_set_foreign_mask(clm_handle* h, int set, const uint8_t* mask, int64_t n, int on_device);
/* n <= 64 DEVICE ints back to the host, ordered behind everything enqueued on the handle's stream: written into mapped
 * pinned memory by a one-warp kernel + an event wait.  For the fill counts of a halo exchange: a cudaMemcpy of the same
 * bytes queues on the device->host DMA engine behind the bulk copy-out of the previous pipelined frame. */
CLM_API int clm_read_ints(clm_handle* h, const int32_t* dev_ints, int32_t n, int32_t* host_out);
CLM_API int clm_cell_coords(clm_handle* h, const void* aos_xyz, int64_t n, int on_device, int axis, int32_t* cell_out);
/* one-pass face selection for the halo exchange; every pointer except `ranges` is a DEVICE pointer and the call only
 * enqueues.  Particles whose cell layer along `axis` is in [ranges[0], ranges[1]) are appended (AoS rows of T) to out_a,
 * those in [ranges[2], ranges[3]) to out_b; merge != 0: either range -> out_a, each particle once.  counts_dev[0..1] must be
 * zeroed by the caller and receive the list lengths; rows beyond `capacity` are counted but not written.  idx_a / idx_b
 * (device, `capacity` entries each, or NULL) receive the 0-based source row of every appended particle, so that the caller can
 * gather the particles' side data (global ids, weights, velocities) into messages of the same order. */
CLM_API int clm_select_layers(clm_handle* h, const void* aos_xyz, int64_t n, int axis, const int32_t ranges[4], int merge,
                              void* out_a, void* out_b, int64_t capacity, int32_t* counts_dev, int32_t* idx_a, int32_t* idx_b);

/* ---- the same decomposition driven by the library itself over NCCL (csrc/clm_comm.cu) ----------------------------------------
 * For hosts without their own communication layer (a Julia or C driver with one process / task per GPU): the slab plan, the halo
 * exchange (ncclSend / ncclRecv on the handle's stream) and the reduction of scalar / histogram results (ncclAllReduce) sit
 * behind these entry points; libnccl.so.2 is opened at run time (CLM_ERR_COMM when it cannot be).  Orthorhombic self-set
 * systems.  Sequence: rank 0 calls clm_comm_unique_id and hands the 128 bytes to every rank by its own means; every rank:
 * clm_create (its device) -> clm_set_box (the GLOBAL box) -> clm_comm_init -> [clm_slab_range: which reference-cell layers
 * along dimension 1 it owns, to partition a global array with clm_cell_coords] -> per step: clm_slab_update(owned particles)
 * -> any clm_map_* (per-particle outputs cover the owned particles; a rank's neighbour list holds the pairs it evaluated,
 * partner indices > n_owned are halo particles) -> clm_comm_allreduce_sum on the scalar / histogram results.
 * clm_slab_update costs one host synchronisation (the received counts); when a halo message outgrows its fixed capacity,
 * EVERY rank learns it from an all-reduced maximum, grows its buffers and repeats the exchange (no rank-local error path). */
CLM_API int clm_comm_unique_id(void* id_out_128_bytes);
CLM_API int clm_comm_init(clm_handle* h, const void* id_128_bytes, int rank, int world);
CLM_API int clm_comm_destroy(clm_handle* h);
CLM_API int clm_slab_range(clm_handle* h, int32_t* layer_lo, int32_t* layer_hi);     /* this rank owns layers [lo, hi) */
CLM_API int clm_slab_update(clm_handle* h, const void* aos_xyz_owned, int64_t n, int on_device);
/* in-place sum over the ranks; kind 0 = the handle's real type, 1 = int64, 2 = double */
CLM_API int clm_comm_allreduce_sum(clm_handle* h, void* buf, int64_t count, int kind, int on_device);
CLM_API int clm_slab_info(clm_handle* h, int64_t* n_owned, int64_t* n_foreign, int32_t* rank, int32_t* world);

/* ---- UpdateCellList!  src/internals/CellLists.jl:727-927 ---------------------------------- */
/* validates coordinates (NaN -> CLM_ERR_INVALID_COORDINATES with the 1-based index in the message),
 * wraps, bins real + image particles, counting-sorts them by cell.  No-op if nothing changed. */
CLM_API int clm_build(clm_handle* h);

/* ---- _pairwise! with the compiled-in functor catalogue ------------------ */
/* LJ: u += c12/d2^6 - c6/d2^3 (test/applications/gromacs/compare_with_gromacs.jl:9-13);
 * params = {c6, c12}.  forces (n_x x N, may be NULL): F_i = -dU/dx_i; self-set systems update both
 * particles of a pair (f[i] += df; f[j] -= df, docs/src/ParticleSystem/examples.md:41-47), two-set
 * systems return the forces on the x set. */
CLM_API int clm_map_lj(clm_handle* h, const void* c6_c12, int flags, void* energy_out, void* forces_out);
/* Coulomb-like: u += k*w_i*w_j/d, F_i = k*w_i*w_j*(x_i-x_j)/d^3
 * (test/examples/gravitational_potential.jl:30-34, gravitational_force.jl:38-44 with k = -9.8). */
CLM_API int clm_map_coulomb(clm_handle* h, const void* weights_x, const void* weights_y, const void* k, int flags,
                    void* energy_out, void* forces_out);
/* distance histogram: counts[floor(d/width)] += 1 (test/examples/distance_histogram.jl:22-26) */
CLM_API int clm_map_dist_hist(clm_handle* h, const void* width, int nbins, int flags, int64_t* counts);
/* halotools-style mean pairwise velocity (test/examples/pairwise_velocities.jl:17-24):
 * b = searchsortedfirst(rbins, r) - 1;  counts[b] += 1;  sums[b] += dot(v_i - v_j, x_i - x_j)/r;
 * rbins has nbins+1 ascending edges, bins are right-closed. */
CLM_API int clm_map_pairvel(clm_handle* h, const void* vel_x, const void* vel_y, const void* rbins, int nbins, int flags,
                    int64_t* counts, void* sums);
/* minimum distance (i, j, d) (test/examples/nearest_neighbor.jl:9-16; docs/src/ParticleSystem/examples.md:117-132);
 * i = j = 0 and d = +Inf when no pair is within the cutoff. */
CLM_API int clm_map_mindist(clm_handle* h, int flags, int64_t* i_out, int64_t* j_out, void* d_out);
/* test functor f1/f2 (test/modules/Testing.jl:23-26): sum of d, sum of d2, number of pairs */
CLM_API int clm_map_sum_d_d2(clm_handle* h, int flags, void* sum_d, void* sum_d2, int64_t* npairs);

/* ---- _pairwise! with a USER pair function ---------------------------- */
/* The reference's pairwise!(f, sys) (src/API/pairwise.jl:48-63) takes an arbitrary Julia closure f(pair, output) and
 * reduces per-task output copies with copy_output / reset_output! / reducer! = `+` for numbers, static vectors and
 * arrays of them (src/API/parallel_custom.jl:53-54, :116-123, :213).  Here the closure body is CUDA C++ source text,
 * compiled at run time with NVRTC (libnvrtc is opened on first use) into the same sm_100a sweep kernel the catalogue
 * uses.  `source` defines a stateless struct `functor_name`:
 *
 *     struct MyPair {
 *         static constexpr int NSCALAR = 1;   // scalar outputs summed over the pairs                    (0..8)
 *         static constexpr int NPART   = 3;   // per-particle output components, output[i] += ...        (0..4)
 *         static constexpr int NAUX    = 1;   // per-particle input components (masses, charges, ...)    (0..4)
 *         static constexpr int HIST    = 0;   // 1: histogram output, counts[bin] += 1; sums[bin] += v
 *         static constexpr unsigned SCALAR_MIN = 0, SCALAR_MAX = 0;   // optional bit masks over the scalar outputs 0..3: reduced
 *                                             // with min / max instead of + (a custom reducer!, src/API/parallel_custom.jl:196-214);
 *                                             // written with out.min_scalar(k, v) / out.max_scalar(k, v); +-Inf when no pair was seen
 *         template <class T, class Out>
 *         __device__ void operator()(const clm::NeighborPair<T>& p, const T* par, Out& out) const {
 *             // p.i, p.j (1-based), p.x[3], p.y[3] (y - x = minimum-image vector), p.d2, p.d(), p.ai[], p.aj[]
 *             out.add_scalar(0, ...); out.add_i(k, ...); out.add_hist(bin, v);
 *         }
 *     };
 *
 * NeighborPair mirrors src/API/NeighborPair.jl:19-33.  Functors without per-particle outputs visit every pair once, in
 * the reference's own mode; functors with per-particle outputs are called once per ORDERED pair (full shell, add_i adds
 * to particle p.i only) and their scalar / histogram outputs are halved for self-set systems: the functor must be
 * symmetric under the exchange of the two particles, as in the reference, where the orientation of (i, j) is unspecified.
 * Errors: CLM_ERR_ARGUMENT with the NVRTC log in clm_last_error() / clm_custom_log() when the source does not compile,
 * CLM_ERR_UNSUPPORTED when libnvrtc cannot be opened. */
typedef struct clm_custom_info { int32_t nscalar, npart, naux, hist, scalar_min_mask, scalar_max_mask; } clm_custom_info;
CLM_API int clm_custom_compile(clm_handle* h, const char* source, const char* functor_name, int32_t* functor_id, clm_custom_info* info_out);
CLM_API const char* clm_custom_log(clm_handle* h); /* NVRTC log of the last compilation of this handle */
/* params: nparams (<= 16) values of T handed to the functor as par[]; aux_x / aux_y: n x NAUX side arrays of the two sets
 * (aux_y only for two-set systems); scalars_out: NSCALAR values of T; per_particle_out: n_x x NPART of T;
 * hist_counts[nbins] / hist_sums[nbins] (T).  CLM_RESET / CLM_OUT_DEVICE as for the catalogue maps. */
CLM_API int clm_map_custom(clm_handle* h, int32_t functor_id, const void* params, int nparams, const void* aux_x, const void* aux_y,
                           int nbins, int flags, void* scalars_out, void* per_particle_out, int64_t* hist_counts, void* hist_sums);
/* compile-only check of a user pair function for every sweep mode (needs libnvrtc, no device); the NVRTC log is
 * copied into log (NUL-terminated, truncated to log_capacity) */
CLM_API int clm_custom_check(const char* source, const char* functor_name, int dtype, char* log, int64_t log_capacity);

/* ---- neighborlist!  src/API/neighborlist.jl:217-231, push_pair! internals/neighborlist.jl:67-76 */
/* Two phases so the caller can resize!() its own Vector{Tuple{Int,Int,T}}: clm_neighborlist builds the
 * list on the device and returns its length; clm_neighborlist_copy writes the 24-byte records
 * {int64 i; int64 j; T d (+padding)} into `records` (host, or device with on_device = 1). */
CLM_API int clm_neighborlist(clm_handle* h, int flags, int64_t* n_out);
CLM_API int clm_neighborlist_copy(clm_handle* h, void* records24, int64_t capacity, int on_device);

CLM_API int clm_get_stats(clm_handle* h, clm_stats* out);
/* tuning knobs (do not change results):
 * "sub" = sub-cells per reference cell and dimension of the device grid (1..7, 0 = from the density),
 * "blocks_per_sm" = resident CTAs per SM of the persistent sweep kernels (0 = as many as fit; k > 0 = at most k; k < 0 = |k| fewer than
 *   fit: two handles that process independent frames in turn leave each other room this way, so that the cell-list build of one
 *   frame runs next to the sweep of the other -- celllistmap.jl_b200/api.py FramePipeline),
 * "n3" = Newton's-third-law force sweep for self-set force maps (1 / 0; -1 = default: Float32 yes, Float64 no),
 * "bin_blocks_per_sm" = grid cap of the binning kernel of the cell-list build in blocks per SM (0 = default: one block per 256
 *   particles; k > 0: at most k blocks per SM striding over the particles). */
CLM_API int clm_set_option(clm_handle* h, const char* name, int64_t value);
CLM_API int clm_version(void);
/* measurement helper: best-of-4 TFLOP/s of a register-resident FMA loop (8 independent chains per thread, all SMs
 * resident) in FP32 (CLM_F32) or FP64 -- the SIMT roofline denominator bench.py reports against */
CLM_API int clm_measure_fma_peak(int device, int dtype, double* tflops);

/* Page-lock a host buffer the caller reuses across calls (the record array of an InPlaceNeighborList, the positions /
 * forces of a trajectory loop), so that the library's host<->device copies of it are direct DMA transfers instead of
 * staged copies through the driver's bounce buffers (5 MB of neighbour-list records: 0.25 -> 0.1 ms).  Thin wrappers of
 * cudaHostRegister / cudaHostUnregister; unregister before the memory is freed.  No reference counterpart (host arrays
 * of the reference never leave the CPU). */
CLM_API int clm_host_register(void* ptr, int64_t bytes);
CLM_API int clm_host_unregister(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* CLM_B200_H */

// Multi-GPU slab decomposition behind the C ABI: one handle per rank / GPU, NCCL for the one exchange step of the path
// (the halo) and for the reduction of scalar / histogram results.  The reference is single-process; the
// entry points are what a Julia / C host calls instead of the torch.distributed plumbing of slab.py:
//   clm_comm_unique_id -> clm_comm_init (every rank) -> per step: clm_slab_update (owned particles; halo exchanged with
//   ncclSend / ncclRecv on the handle's stream and handed to the engine as FOREIGN particles) -> any clm_map_* (forces stay
//   with their owners) -> clm_comm_allreduce_sum on the scalar / histogram results.
// libnccl.so.2 is opened at run time (the copy the process already holds, e.g. torch's, else the system one): the library
// links no NCCL, and a host that never calls clm_comm_* needs none.
#include <dlfcn.h>
#include <cmath>
#include <cstring>
#include "clm_engine.cuh"

namespace clm {

// ---- the few NCCL entry points, resolved at run time (nccl.h is not needed to build) -------------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };   // ncclDataType_t
enum { NCCL_SUM = 0, NCCL_MAX = 2 };                                            // ncclRedOp_t
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
            if (!lib) lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { why = std::string("libnccl.so.2 cannot be opened: ") + (dlerror() ? dlerror() : "?"); return false; }
        bool ok = true;
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) { ok = false; why = std::string("missing NCCL symbol ") + n; } return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) lib = nullptr;
        return ok;
    }
};
static NcclApi g_nccl;

#define CLM_NCCL(call)                                                                                              \
    do {                                                                                                            \
        const int r_ = (call);                                                                                      \
        if (r_ != 0) return this->fail(CLM_ERR_COMM, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error")); \
    } while (0)

template <class T> struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int lo = 0, hi = 0;               // this rank's reference-cell layers along dimension 1: [lo, hi)
    int64_t cap = 0;                  // rows per halo message
    DBuf<T> own, send[2], recv[2];
    DBuf<int> cnt;                    // [0..1] my face counts, [2..3] received counts (upper, lower), [4] global max face count
    int* h_cnt = nullptr;             // pinned mirror
    int64_t n_foreign = 0;
};

template <class T> static CommState<T>*& comm_of(Engine<T>* e) { return reinterpret_cast<CommState<T>*&>(e->comm_state); }

template <class T> int Engine<T>::comm_init(const void* id, int rank, int world) {
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(CLM_ERR_ARGUMENT, "clm_comm_init: bad id / rank / world");
    if (!g_nccl.load()) return fail(CLM_ERR_COMM, g_nccl.why);
    CLM_CK(cudaSetDevice(device));
    comm_destroy();
    auto* c = new CommState<T>();
    comm_of(this) = c;
    c->rank = rank; c->world = world;
    NcclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    CLM_NCCL(g_nccl.CommInitRank(&c->comm, world, uid, rank));
    CLM_CK(c->cnt.ensure(8));
    CLM_CK(cudaMallocHost((void**)&c->h_cnt, 8 * sizeof(int)));
    return CLM_OK;
}
template <class T> int Engine<T>::comm_destroy() {
    auto*& c = comm_of(this);
    if (!c) return CLM_OK;
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    c->own.release(); c->cnt.release();
    for (int k = 0; k < 2; ++k) { c->send[k].release(); c->recv[k].release(); }
    delete c;
    c = nullptr;
    return CLM_OK;
}

// slabs of whole reference cells along dimension 1 (the same plan as slab.py: rank r owns layers
// [lcell + n_inner r / world, lcell + n_inner (r + 1) / world), n_inner = nc[0] - 2 lcell - 1 layers of real particles)
template <class T> int Engine<T>::slab_range(int32_t* lo, int32_t* hi) {
    auto* c = comm_of(this);
    if (!c) return fail(CLM_ERR_STATE, "clm_comm_init must be called first");
    if (!box_set || nonperiodic) return fail(CLM_ERR_STATE, "clm_set_box (periodic) must be called first");
    if (box.cell_type != CLM_ORTHORHOMBIC) return fail(CLM_ERR_UNSUPPORTED, "slab decomposition supports orthorhombic cells only (triclinic lattice shifts move images across slabs)");
    const int lcell = box.lcell, n_inner = (int)box.nc[0] - 2 * lcell - 1;
    if (c->world > 1 && n_inner / c->world < lcell) return fail(CLM_ERR_ARGUMENT, "slabs would be thinner than the stencil reach");
    c->lo = lcell + (int)(((int64_t)n_inner * c->rank) / c->world);
    c->hi = lcell + (int)(((int64_t)n_inner * (c->rank + 1)) / c->world);
    if (lo) *lo = c->lo;
    if (hi) *hi = c->hi;
    return CLM_OK;
}

// owned particles -> the engine; the lcell outermost layers of both neighbours (periodic) -> foreign particles.  The face
// selection runs in one kernel into fixed-capacity messages; messages and their fill counts travel in one NCCL group; an
// all-reduced maximum of the face counts makes the (rare) "message too small" decision collective: every rank grows its
// buffers and repeats the exchange.  One host synchronisation per call (the received counts).
template <class T> int Engine<T>::slab_update(const void* xyz, int64_t n, int on_device) {
    auto* c = comm_of(this);
    if (!c) return fail(CLM_ERR_STATE, "clm_comm_init must be called first");
    if (n < 0 || (n > 0 && !xyz)) return fail(CLM_ERR_ARGUMENT, "bad particle array");
    if (n == 0) return fail(CLM_ERR_ARGUMENT, "a rank without particles is not supported");
    if (int rc = slab_range(nullptr, nullptr)) return rc;
    if (int rc = set_positions(0, xyz, n, on_device)) return rc;
    if (c->world == 1) { c->n_foreign = 0; return set_foreign(0, nullptr, 0, 1); }
    CLM_CK(cudaSetDevice(device));
    const int lcell = box.lcell, world = c->world, lower = (c->rank + world - 1) % world, upper = (c->rank + 1) % world;
    const bool merge = world == 2;      // both faces go to the one peer: one message, every particle once
    if (c->cap == 0) {
        // first call, one collective: the global particle count.  (1) The engine sizes its device grid from the particle
        // density; a rank only sees its slab, so it is told the global density (same rule as Engine::build_enqueue: ~4
        // particles per device cell).  (2) The capacity of the halo messages must be the SAME on every rank (a send and its
        // matching receive carry cap rows): it follows from the global count -- a face holds about lcell / (layers per
        // rank) of a rank's share; 50 % slack -- and grows collectively (below) when a face outgrows it.
        CLM_CK(c->own.ensure(4));
        long long hn = (long long)n;
        CLM_CK(cudaMemcpyAsync(c->own.p, &hn, sizeof(hn), cudaMemcpyHostToDevice, stream));
        CLM_NCCL(g_nccl.AllReduce(c->own.p, c->own.p, 1, NCCL_INT64, NCCL_SUM, c->comm, stream));
        CLM_CK(cudaMemcpyAsync(&hn, c->own.p, sizeof(hn), cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
        if (opt_sub == 0) {
            double inner = 1;
            for (int k = 0; k < dim; ++k) inner *= (double)std::max<int64_t>(1, box.nc[k] - 2 * box.lcell - 1);
            const int sub = (int)std::floor(std::pow(std::max((double)hn / inner / 4.0, 1.0), 1.0 / dim) + 0.35);
            opt_sub = std::max(1, std::min(sub, SUB_MAX / box.lcell));
        }
        const int layers = std::max(1, ((int)box.nc[0] - 2 * lcell - 1) / world);
        const double frac = std::min(1.0, (double)lcell / layers * (merge ? 2.0 : 1.0));
        c->cap = (int64_t)((double)hn / world * frac * 1.5) + 1024;
    }
    const T* x_dev = sets[0].pos.p;     // the engine's own copy of the owned particles
    const int dt = sizeof(T) == 4 ? NCCL_FLOAT32 : NCCL_FLOAT64;
    for (int attempt = 0; attempt < 4; ++attempt) {
        for (int k = 0; k < 2; ++k) { CLM_CK(c->send[k].ensure((size_t)c->cap * dim)); CLM_CK(c->recv[k].ensure((size_t)c->cap * dim)); }
        CLM_CK(cudaMemsetAsync(c->cnt.p, 0, 8 * sizeof(int), stream));
        // the last rank's upper face is closed at its top layer: a real particle can round into layer lcell + n_inner
        const int32_t ranges[4] = {c->lo, c->lo + lcell, c->hi - lcell, c->hi + (c->rank == world - 1 ? 1 : 0)};
        if (int rc = select_layers(x_dev, n, 0, ranges, merge ? 1 : 0, c->send[0].p, c->send[1].p, c->cap, c->cnt.p, nullptr, nullptr)) return rc;
        const size_t msg = (size_t)c->cap * dim;
        CLM_NCCL(g_nccl.GroupStart());
        if (merge) {
            CLM_NCCL(g_nccl.Send(c->send[0].p, msg, dt, upper, c->comm, stream));
            CLM_NCCL(g_nccl.Send(c->cnt.p, 1, NCCL_INT32, upper, c->comm, stream));
            CLM_NCCL(g_nccl.Recv(c->recv[0].p, msg, dt, upper, c->comm, stream));
            CLM_NCCL(g_nccl.Recv(c->cnt.p + 2, 1, NCCL_INT32, upper, c->comm, stream));
        } else {
            CLM_NCCL(g_nccl.Send(c->send[0].p, msg, dt, lower, c->comm, stream));       // my lower face -> lower neighbour
            CLM_NCCL(g_nccl.Send(c->cnt.p, 1, NCCL_INT32, lower, c->comm, stream));
            CLM_NCCL(g_nccl.Send(c->send[1].p, msg, dt, upper, c->comm, stream));       // my upper face -> upper neighbour
            CLM_NCCL(g_nccl.Send(c->cnt.p + 1, 1, NCCL_INT32, upper, c->comm, stream));
            CLM_NCCL(g_nccl.Recv(c->recv[0].p, msg, dt, upper, c->comm, stream));       // the upper neighbour's lower face
            CLM_NCCL(g_nccl.Recv(c->cnt.p + 2, 1, NCCL_INT32, upper, c->comm, stream));
            CLM_NCCL(g_nccl.Recv(c->recv[1].p, msg, dt, lower, c->comm, stream));       // the lower neighbour's upper face
            CLM_NCCL(g_nccl.Recv(c->cnt.p + 3, 1, NCCL_INT32, lower, c->comm, stream));
        }
        CLM_NCCL(g_nccl.GroupEnd());
        // global maximum of the face counts: the overflow decision is the same on every rank
        k_max2<<<1, 32, 0, stream>>>(c->cnt.p, c->cnt.p + 4);
        CLM_CK(cudaGetLastError());
        CLM_NCCL(g_nccl.AllReduce(c->cnt.p + 4, c->cnt.p + 4, 1, NCCL_INT32, NCCL_MAX, c->comm, stream));
        if (int rc = read_ints(c->cnt.p, 8, c->h_cnt)) return rc;      // the one synchronisation (mapped pinned memory, no DMA copy)
        stats.launches += 2;
        if ((int64_t)c->h_cnt[4] <= c->cap) break;
        if (attempt == 3) return fail(CLM_ERR_COMM, "halo messages did not converge on a capacity");
        c->cap = (int64_t)((double)c->h_cnt[4] * 1.5) + 1024;     // every rank computes the same new capacity
    }
    const int64_t n_up = c->h_cnt[2], n_lo = merge ? 0 : c->h_cnt[3];
    DevSet<T>& s = sets[0];
    CLM_CK(s.fpos.ensure((size_t)std::max<int64_t>(n_up + n_lo, 1) * dim));
    if (n_up) CLM_CK(cudaMemcpyAsync(s.fpos.p, c->recv[0].p, (size_t)n_up * dim * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    if (n_lo) CLM_CK(cudaMemcpyAsync(s.fpos.p + (size_t)n_up * dim, c->recv[1].p, (size_t)n_lo * dim * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    s.n_foreign = n_up + n_lo;
    c->n_foreign = s.n_foreign;
    dirty = true;
    return CLM_OK;
}

// in-place sum over the ranks: kind 0 = the handle's real type, 1 = int64, 2 = double
template <class T> int Engine<T>::comm_allreduce_sum(void* buf, int64_t count, int kind, int on_device) {
    auto* c = comm_of(this);
    if (!c) return fail(CLM_ERR_STATE, "clm_comm_init must be called first");
    if (count <= 0) return CLM_OK;
    if (!buf || kind < 0 || kind > 2) return fail(CLM_ERR_ARGUMENT, "bad buffer / kind");
    if (c->world == 1) return CLM_OK;
    CLM_CK(cudaSetDevice(device));
    const size_t esz = kind == 0 ? sizeof(T) : 8;
    const int dt = kind == 0 ? (sizeof(T) == 4 ? NCCL_FLOAT32 : NCCL_FLOAT64) : (kind == 1 ? NCCL_INT64 : NCCL_FLOAT64);
    void* d = buf;
    if (!on_device) {
        CLM_CK(c->own.ensure((size_t)((count * esz + sizeof(T) - 1) / sizeof(T))));
        d = c->own.p;
        CLM_CK(cudaMemcpyAsync(d, buf, (size_t)count * esz, cudaMemcpyHostToDevice, stream));
    }
    CLM_NCCL(g_nccl.AllReduce(d, d, (size_t)count, dt, NCCL_SUM, c->comm, stream));
    if (!on_device) {
        CLM_CK(cudaMemcpyAsync(buf, d, (size_t)count * esz, cudaMemcpyDeviceToHost, stream));
        CLM_CK(cudaStreamSynchronize(stream));
    }
    return CLM_OK;
}
template <class T> int Engine<T>::slab_info(int64_t* n_owned, int64_t* n_foreign, int32_t* rank, int32_t* world) {
    auto* c = comm_of(this);
    if (!c) return fail(CLM_ERR_STATE, "clm_comm_init must be called first");
    if (n_owned) *n_owned = sets[0].n;
    if (n_foreign) *n_foreign = sets[0].n_foreign;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return CLM_OK;
}

int comm_unique_id(void* out, std::string& err) {
    if (!out) { err = "NULL pointer"; return CLM_ERR_ARGUMENT; }
    if (!g_nccl.load()) { err = g_nccl.why; return CLM_ERR_COMM; }
    NcclUniqueId uid;
    const int r = g_nccl.GetUniqueId(&uid);
    if (r != 0) { err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return CLM_ERR_COMM; }
    std::memcpy(out, &uid, sizeof(uid));
    return CLM_OK;
}

#define INST(T)                                                                     \
    template int Engine<T>::comm_init(const void*, int, int);                       \
    template int Engine<T>::comm_destroy();                                         \
    template int Engine<T>::slab_range(int32_t*, int32_t*);                         \
    template int Engine<T>::slab_update(const void*, int64_t, int);                 \
    template int Engine<T>::comm_allreduce_sum(void*, int64_t, int, int);           \
    template int Engine<T>::slab_info(int64_t*, int64_t*, int32_t*, int32_t*);
INST(float)
INST(double)

}  // namespace clm

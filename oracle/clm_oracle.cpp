// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY (see clm_oracle.hpp header).
// C entry points (ctypes-friendly) over the templated restatement, the compiled-in
// functor catalogue as the reference's tests/docs define it,
// and the reference's batched task parallelism (self.jl:81-96, cross.jl:62-76):
// round-robin chunks of home cells, one private output copy per batch, reduced in
// batch order (internals/ParticleSystem.jl:106-123).
// =============================================================================
#include "clm_oracle.hpp"
#include <omp.h>
#include <memory>

using namespace ora;

namespace {

struct Rec24 {  // Tuple{Int,Int,T}: 24 bytes for both Float32 and Float64
    int64_t i, j;
    double dbits;  // T stored in the first sizeof(T) bytes
};

struct Base {
    std::string err;
    virtual ~Base() {}
    virtual int set_box(int cell_type, const void* cell, int is_matrix, const void* cutoff, int lcell) = 0;
    virtual int set_positions(int set, const void* xyz, int64_t n) = 0;
    virtual int build() = 0;
    virtual int get_box(double* out) = 0;
    virtual int get_stats(int64_t* out) = 0;
    virtual int map_sum(int algo, int nb, void* sum_d, void* sum_d2, int64_t* npairs) = 0;
    virtual int map_lj(int algo, int nb, const void* c6c12, void* energy, void* forces) = 0;
    virtual int map_coulomb(int algo, int nb, const void* wx, const void* wy, const void* k, void* energy, void* forces) = 0;
    virtual int map_dist_hist(int algo, int nb, const void* width, int nbins, int64_t* counts) = 0;
    virtual int map_pairvel(int algo, int nb, const void* vx, const void* vy, const void* rbins, int nbins, int64_t* counts,
                            void* sums) = 0;
    virtual int map_mindist(int algo, int nb, int64_t* i, int64_t* j, void* d) = 0;
    virtual int neighborlist(int algo, int nb, int64_t* n) = 0;
    virtual int neighborlist_copy(void* rec, int64_t cap) = 0;
    virtual int candidates(int64_t* out) = 0;
};

template <class T, int N> struct Handle : Base {
    Box<T, N> box;
    bool have_box = false, nonperiodic = false, built = false;
    T np_cutoff = 0;
    int np_lcell = 1;
    std::vector<T> x, y;
    int64_t nx = 0, ny = 0;
    bool two_sets = false;
    CellList<T, N> clx, cly;
    std::vector<Rec24> nl;

    int set_box(int cell_type, const void* cell, int is_matrix, const void* cutoff, int lcell) override {
        if (lcell < 1) { err = "lcell must be greater or equal to 1"; return 3; }
        T rc = *(const T*)cutoff;
        built = false;
        if (cell_type == NONPERIODIC) {  // Box(limits(x[,y]), cutoff): deferred to build()
            nonperiodic = true; np_cutoff = rc; np_lcell = lcell; have_box = false;
            return 0;
        }
        nonperiodic = false;
        Mat<T, N> M;
        const T* c = (const T*)cell;
        if (is_matrix) { for (int k = 0; k < N * N; ++k) M.m[k] = c[k]; }
        else { for (int k = 0; k < N * N; ++k) M.m[k] = T(0); for (int k = 0; k < N; ++k) M(k, k) = c[k]; }
        Vec<T, N> origin;
        for (int k = 0; k < N; ++k) origin[k] = T(0);
        if (!construct_box(box, cell_type, M, rc, lcell, origin)) { err = "Unit cell matrix does not satisfy required conditions."; return 2; }
        have_box = true;
        return 0;
    }
    int set_positions(int set, const void* xyz, int64_t n) override {
        std::vector<T>& v = set ? y : x;
        v.assign((const T*)xyz, (const T*)xyz + n * N);
        (set ? ny : nx) = n;
        if (set) two_sets = true;
        built = false;
        return 0;
    }
    // _validate_coordinates (CellOperations.jl:6-21)
    int validate(const std::vector<T>& v, int64_t n) {
        for (int64_t i = 0; i < n; ++i)
            for (int k = 0; k < N; ++k)
                if (std::isnan(v[i * N + k])) { err = "Invalid coordinates found for particle of index " + std::to_string(i + 1); return 1; }
        return 0;
    }
    static void minmax(const std::vector<T>& v, int64_t n, Vec<T, N>& lo, Vec<T, N>& hi) {  // _minmax CellOperations.jl:262-275
        if (n == 0) { for (int k = 0; k < N; ++k) lo[k] = hi[k] = T(0); return; }
        for (int k = 0; k < N; ++k) { lo[k] = std::numeric_limits<T>::max(); hi[k] = std::numeric_limits<T>::lowest(); }
        for (int64_t i = 0; i < n; ++i)
            for (int k = 0; k < N; ++k) { lo[k] = std::min(lo[k], v[i * N + k]); hi[k] = std::max(hi[k], v[i * N + k]); }
    }
    int build() override {
        if (int e = validate(x, nx)) return e;
        if (two_sets) if (int e = validate(y, ny)) return e;
        if (nonperiodic) {  // limits (CellOperations.jl:290-324) + Box(::Limits) (Box.jl:38, :374-377)
            Vec<T, N> lo, hi;
            minmax(x, nx, lo, hi);
            if (two_sets) {
                Vec<T, N> lo2, hi2;
                minmax(y, ny, lo2, hi2);
                for (int k = 0; k < N; ++k) { lo[k] = std::min(lo[k], lo2[k]); hi[k] = std::max(hi[k], hi2[k]); }
            }
            Mat<T, N> M;
            for (int k = 0; k < N * N; ++k) M.m[k] = T(0);
            T pad = T(210) * np_cutoff / T(100);
            for (int k = 0; k < N; ++k) M(k, k) = (hi[k] - lo[k]) + pad;
            if (!construct_box(box, NONPERIODIC, M, np_cutoff, np_lcell, lo)) { err = "Unit cell matrix does not satisfy required conditions."; return 2; }
            have_box = true;
        }
        if (!have_box) { err = "box not set"; return 4; }
        try {
            build_cell_list(x.data(), nx, box, clx);
            if (two_sets) build_cell_list(y.data(), ny, box, cly);
        } catch (std::exception& e) { err = e.what(); return 5; }
        built = true;
        return 0;
    }
    int get_box(double* o) override {  // 9+9+9+9 matrices, then nc[3], cutoff, cutoff_sqr, cbmin[3], cbmax[3], cs[3], origin[3], lcell, type
        if (!have_box) { err = "box not set"; return 4; }
        for (int k = 0; k < 60; ++k) o[k] = 0;
        for (int k = 0; k < N * N; ++k) { o[k] = box.input_unit_cell.m[k]; o[9 + k] = box.aligned_unit_cell.m[k]; o[18 + k] = box.rotation.m[k]; o[27 + k] = box.inv_rotation.m[k]; }
        for (int k = 0; k < N; ++k) { o[36 + k] = (double)box.nc[k]; o[41 + k] = box.cb_min[k]; o[44 + k] = box.cb_max[k]; o[47 + k] = box.cell_size[k]; o[50 + k] = box.origin[k]; }
        o[39] = box.cutoff; o[40] = box.cutoff_sqr; o[53] = box.lcell; o[54] = box.cell_type;
        return 0;
    }
    int get_stats(int64_t* o) override {
        if (!built) { err = "not built"; return 4; }
        o[0] = clx.n_real_particles; o[1] = clx.n_particles; o[2] = (int64_t)clx.cell_indices_real.size(); o[3] = (int64_t)clx.cells.size();
        o[4] = two_sets ? cly.n_real_particles : 0; o[5] = two_sets ? cly.n_particles : 0;
        o[6] = two_sets ? (int64_t)cly.cell_indices_real.size() : 0; o[7] = two_sets ? (int64_t)cly.cells.size() : 0;
        return 0;
    }
    // number of candidate pairs of the reference's own stencil on the reference's own grid
    // (SURVEY.md §8(d) C_st), real/ghost rules ignored: out[0] = same-cell, out[1] = vicinal
    int candidates(int64_t* out) override {
        if (!built) { err = "not built"; return 4; }
        const int kind = two_sets ? CROSS : SELF;
        const bool forward = (kind == SELF) && (box.cell_type != TRICLINIC);
        auto st = make_stencil<N>(box.lcell, forward);
        const CellList<T, N>& tg = two_sets ? cly : clx;
        int64_t same = 0, vic = 0;
        for (int64_t s : clx.cell_indices_real) {
            const Cell<T, N>& ci = clx.cells[s];
            int64_t ni = (int64_t)ci.particles.size();
            if (kind == SELF) same += forward ? ni * (ni - 1) / 2 : ni * ni;
            else { int64_t sl = tg.cell_indices[ci.linear_index - 1]; if (sl) same += ni * (int64_t)tg.cells[sl - 1].particles.size(); }
            for (auto& off : st) {
                int64_t c[N];
                for (int k = 0; k < N; ++k) c[k] = ci.cart[k] + off[k];
                int64_t sl = tg.cell_indices[linear_index<N>(box.nc, c) - 1];
                if (sl) vic += ni * (int64_t)tg.cells[sl - 1].particles.size();
            }
        }
        out[0] = same; out[1] = vic;
        return 0;
    }

    // generic runner: algo 0 = cell lists + projection filter (reference default),
    // 1 = cell lists without the filter, 2 = naive O(N^2) twin.  nb = 0: serial path;
    // nb >= 1: reference's batched parallel path with nb batches.
    template <class Fn> int run(int algo, int nb, Fn& out) {
        if (!built) { err = "cell lists not built"; return 4; }
        const int kind = two_sets ? CROSS : SELF;
        try {
            if (algo == 2) {
                if (kind == SELF) map_naive_self<T, N>(x.data(), nx, box, [&](const Pair<T, N>& p) { out(p); });
                else map_naive_cross<T, N>(x.data(), nx, y.data(), ny, box, [&](const Pair<T, N>& p) { out(p); });
                return 0;
            }
            const bool proj = (algo == 0);
            const CellList<T, N>& tg = two_sets ? cly : clx;
            if (nb <= 0) {
                map_pairwise_serial(box, clx, tg, kind, proj, [&](const Pair<T, N>& p) { out(p); });
                return 0;
            }
            std::vector<Fn> copies((size_t)nb, out.fresh());
            const bool forward = (kind == SELF) && (box.cell_type != TRICLINIC);
            auto stencil = make_stencil<N>(box.lcell, forward);
            const int64_t nhome = (int64_t)clx.cell_indices_real.size();
            std::string ferr;
#pragma omp parallel for schedule(static, 1)
            for (int b = 0; b < nb; ++b) {
                std::vector<Projected<T, N>> scratch;
                Fn& o = copies[(size_t)b];
                try {
                    for (int64_t c = b; c < nhome; c += nb)  // index_chunks(...; split = RoundRobin())
                        inner_loop(box, clx.cells[clx.cell_indices_real[c]], tg, kind, stencil, scratch, proj,
                                   [&](const Pair<T, N>& p) { o(p); });
                } catch (std::exception& e) {
#pragma omp critical
                    ferr = e.what();
                }
            }
            if (!ferr.empty()) { err = ferr; return 5; }
            for (int b = 0; b < nb; ++b) out.merge(copies[(size_t)b]);
        } catch (std::exception& e) { err = e.what(); return 5; }
        return 0;
    }

    // ---- catalogue -------------------------------------------------------
    struct FSum {  // f1/f2 of test/modules/Testing.jl:23-26 plus a pair count
        T sd = 0, sd2 = 0; int64_t n = 0;
        FSum fresh() const { return FSum(); }
        void operator()(const Pair<T, N>& p) { sd += std::sqrt(p.d2); sd2 += p.d2; ++n; }
        void merge(const FSum& o) { sd += o.sd; sd2 += o.sd2; n += o.n; }
    };
    struct FLJ {  // test/applications/gromacs/compare_with_gromacs.jl:9-13; force pattern docs/src/ParticleSystem/examples.md:41-47
        T c6, c12, u = 0; bool want_f; int64_t n; std::vector<T> f;
        FLJ(T c6_, T c12_, bool wf, int64_t n_) : c6(c6_), c12(c12_), want_f(wf), n(n_) { if (wf) f.assign((size_t)n_ * N, T(0)); }
        FLJ fresh() const { FLJ r(c6, c12, want_f, n); r.two = two; return r; }
        void operator()(const Pair<T, N>& p) {
            T d2 = p.d2, d6 = d2 * d2 * d2;
            u += c12 / std::pow(d2, T(6)) - c6 / d6;
            if (want_f) {  // F_i = -dU/dx_i = (12 c12/d2^7 - 6 c6/d2^4) (x_i - x_j)
                T fs = (T(12) * c12 / (d6 * d6) - T(6) * c6 / d6) / d2;
                for (int k = 0; k < N; ++k) {
                    T df = fs * (p.x[k] - p.y[k]);
                    f[(size_t)(p.i - 1) * N + k] += df;
                    if (two) f[(size_t)(p.j - 1) * N + k] -= df;
                }
            }
        }
        bool two = true;  // self-set: Newton's third law updates both; cross-set: forces on the first set only
        void merge(const FLJ& o) { u += o.u; for (size_t k = 0; k < f.size(); ++k) f[k] += o.f[k]; }
    };
    struct FCoul {  // test/examples/gravitational_potential.jl:30-34, gravitational_force.jl:38-44 (k carries the sign)
        const T *wx, *wy; T k, u = 0; bool want_f, two; int64_t n; std::vector<T> f;
        FCoul(const T* wx_, const T* wy_, T k_, bool wf, bool two_, int64_t n_) : wx(wx_), wy(wy_), k(k_), want_f(wf), two(two_), n(n_) { if (wf) f.assign((size_t)n_ * N, T(0)); }
        FCoul fresh() const { return FCoul(wx, wy, k, want_f, two, n); }
        void operator()(const Pair<T, N>& p) {
            T d = std::sqrt(p.d2);
            T q = k * wx[p.i - 1] * wy[p.j - 1];
            u += q / d;
            if (want_f) {  // F_i = q (x_i - x_j) / d^3
                T g = q / p.d2 / d;
                for (int c = 0; c < N; ++c) {
                    T df = g * (p.x[c] - p.y[c]);
                    f[(size_t)(p.i - 1) * N + c] += df;
                    if (two) f[(size_t)(p.j - 1) * N + c] -= df;
                }
            }
        }
        void merge(const FCoul& o) { u += o.u; for (size_t c = 0; c < f.size(); ++c) f[c] += o.f[c]; }
    };
    struct FHist {  // test/examples/distance_histogram.jl:22-26 (ibin = floor(Int, d/width) + 1; out-of-range pairs dropped)
        T width; std::vector<int64_t> h;
        FHist(T w, int nb) : width(w), h((size_t)nb, 0) {}
        FHist fresh() const { return FHist(width, (int)h.size()); }
        void operator()(const Pair<T, N>& p) {
            T d = std::sqrt(p.d2);
            int64_t b = (int64_t)std::floor(d / width);
            if (b >= 0 && b < (int64_t)h.size()) h[(size_t)b] += 1;
        }
        void merge(const FHist& o) { for (size_t k = 0; k < h.size(); ++k) h[k] += o.h[k]; }
    };
    struct FVel {  // test/examples/pairwise_velocities.jl:17-24
        const T *vx, *vy, *rbins; int nb; std::vector<int64_t> cnt; std::vector<T> sum;
        FVel(const T* vx_, const T* vy_, const T* rb, int nb_) : vx(vx_), vy(vy_), rbins(rb), nb(nb_), cnt((size_t)nb_, 0), sum((size_t)nb_, T(0)) {}
        FVel fresh() const { return FVel(vx, vy, rbins, nb); }
        void operator()(const Pair<T, N>& p) {
            T r = std::sqrt(p.d2);
            int first = 0;  // searchsortedfirst(rbins, r) - 1 (1-based) -> 0-based bin = first - 1
            while (first < nb + 1 && rbins[first] < r) ++first;
            int b = first - 1;
            if (b < 0 || b >= nb) return;  // BoundsError in the reference; dropped here
            T dv[N], d[N];
            for (int k = 0; k < N; ++k) { dv[k] = vx[(size_t)(p.i - 1) * N + k] - vy[(size_t)(p.j - 1) * N + k]; d[k] = p.x[k] - p.y[k]; }
            cnt[(size_t)b] += 1;
            sum[(size_t)b] += dotn<T, N>(dv, d) / r;
        }
        void merge(const FVel& o) { for (int k = 0; k < nb; ++k) { cnt[(size_t)k] += o.cnt[(size_t)k]; sum[(size_t)k] += o.sum[(size_t)k]; } }
    };
    struct FMin {  // test/examples/nearest_neighbor.jl:9-16, :43 ; docs/src/ParticleSystem/examples.md:117-132
        int64_t i = 0, j = 0; T d = std::numeric_limits<T>::infinity();
        FMin fresh() const { return FMin(); }
        void operator()(const Pair<T, N>& p) { T dd = std::sqrt(p.d2); if (dd < d) { i = p.i; j = p.j; d = dd; } }
        void merge(const FMin& o) { if (!(d <= o.d)) { i = o.i; j = o.j; d = o.d; } }
    };
    struct FList {  // push_pair! (internals/neighborlist.jl:67-76), concat reduce (:25-40)
        std::vector<Rec24> l;
        FList fresh() const { return FList(); }
        void operator()(const Pair<T, N>& p) {
            Rec24 r; r.i = p.i; r.j = p.j; r.dbits = 0;
            T d = std::sqrt(p.d2);
            std::memcpy(&r.dbits, &d, sizeof(T));
            l.push_back(r);
        }
        void merge(const FList& o) { l.insert(l.end(), o.l.begin(), o.l.end()); }
    };

    int map_sum(int algo, int nb, void* sum_d, void* sum_d2, int64_t* npairs) override {
        FSum f;
        if (int e = run(algo, nb, f)) return e;
        *(T*)sum_d = f.sd; *(T*)sum_d2 = f.sd2; *npairs = f.n;
        return 0;
    }
    int map_lj(int algo, int nb, const void* c6c12, void* energy, void* forces) override {
        const T* c = (const T*)c6c12;
        FLJ f(c[0], c[1], forces != nullptr, nx);
        f.two = !two_sets;
        if (int e = run(algo, nb, f)) return e;
        *(T*)energy = f.u;
        if (forces) std::memcpy(forces, f.f.data(), f.f.size() * sizeof(T));
        return 0;
    }
    int map_coulomb(int algo, int nb, const void* wx, const void* wy, const void* k, void* energy, void* forces) override {
        const T* wyp = two_sets ? (const T*)wy : (const T*)wx;
        FCoul f((const T*)wx, wyp, *(const T*)k, forces != nullptr, !two_sets, nx);
        if (int e = run(algo, nb, f)) return e;
        *(T*)energy = f.u;
        if (forces) std::memcpy(forces, f.f.data(), f.f.size() * sizeof(T));
        return 0;
    }
    int map_dist_hist(int algo, int nb, const void* width, int nbins, int64_t* counts) override {
        FHist f(*(const T*)width, nbins);
        if (int e = run(algo, nb, f)) return e;
        std::memcpy(counts, f.h.data(), (size_t)nbins * sizeof(int64_t));
        return 0;
    }
    int map_pairvel(int algo, int nb, const void* vx, const void* vy, const void* rbins, int nbins, int64_t* counts, void* sums) override {
        const T* vyp = two_sets ? (const T*)vy : (const T*)vx;
        FVel f((const T*)vx, vyp, (const T*)rbins, nbins);
        if (int e = run(algo, nb, f)) return e;
        std::memcpy(counts, f.cnt.data(), (size_t)nbins * sizeof(int64_t));
        std::memcpy(sums, f.sum.data(), (size_t)nbins * sizeof(T));
        return 0;
    }
    int map_mindist(int algo, int nb, int64_t* i, int64_t* j, void* d) override {
        FMin f;
        if (int e = run(algo, nb, f)) return e;
        *i = f.i; *j = f.j; *(T*)d = f.d;
        return 0;
    }
    int neighborlist(int algo, int nb, int64_t* n) override {
        FList f;
        if (int e = run(algo, nb, f)) return e;
        nl.swap(f.l);
        *n = (int64_t)nl.size();
        return 0;
    }
    int neighborlist_copy(void* rec, int64_t cap) override {
        if ((int64_t)nl.size() > cap) { err = "capacity too small"; return 6; }
        std::memcpy(rec, nl.data(), nl.size() * sizeof(Rec24));
        return 0;
    }
};

}  // namespace

extern "C" {
void* ora_create(int dim, int dtype) {
    if (dim == 2 && dtype == 0) return new Handle<float, 2>();
    if (dim == 3 && dtype == 0) return new Handle<float, 3>();
    if (dim == 2 && dtype == 1) return new Handle<double, 2>();
    if (dim == 3 && dtype == 1) return new Handle<double, 3>();
    return nullptr;
}
void ora_destroy(void* h) { delete (Base*)h; }
const char* ora_last_error(void* h) { return ((Base*)h)->err.c_str(); }
int ora_set_box(void* h, int cell_type, const void* cell, int is_matrix, const void* cutoff, int lcell) { return ((Base*)h)->set_box(cell_type, cell, is_matrix, cutoff, lcell); }
int ora_set_positions(void* h, int set, const void* xyz, int64_t n) { return ((Base*)h)->set_positions(set, xyz, n); }
int ora_build(void* h) { return ((Base*)h)->build(); }
int ora_get_box(void* h, double* out60) { return ((Base*)h)->get_box(out60); }
int ora_get_stats(void* h, int64_t* out8) { return ((Base*)h)->get_stats(out8); }
int ora_candidates(void* h, int64_t* out2) { return ((Base*)h)->candidates(out2); }
int ora_map_sum_d_d2(void* h, int algo, int nb, void* sd, void* sd2, int64_t* np) { return ((Base*)h)->map_sum(algo, nb, sd, sd2, np); }
int ora_map_lj(void* h, int algo, int nb, const void* c6c12, void* e, void* f) { return ((Base*)h)->map_lj(algo, nb, c6c12, e, f); }
int ora_map_coulomb(void* h, int algo, int nb, const void* wx, const void* wy, const void* k, void* e, void* f) { return ((Base*)h)->map_coulomb(algo, nb, wx, wy, k, e, f); }
int ora_map_dist_hist(void* h, int algo, int nb, const void* width, int nbins, int64_t* counts) { return ((Base*)h)->map_dist_hist(algo, nb, width, nbins, counts); }
int ora_map_pairvel(void* h, int algo, int nb, const void* vx, const void* vy, const void* rbins, int nbins, int64_t* counts, void* sums) { return ((Base*)h)->map_pairvel(algo, nb, vx, vy, rbins, nbins, counts, sums); }
int ora_map_mindist(void* h, int algo, int nb, int64_t* i, int64_t* j, void* d) { return ((Base*)h)->map_mindist(algo, nb, i, j, d); }
int ora_neighborlist(void* h, int algo, int nb, int64_t* n) { return ((Base*)h)->neighborlist(algo, nb, n); }
int ora_neighborlist_copy(void* h, void* rec, int64_t cap) { return ((Base*)h)->neighborlist_copy(rec, cap); }
int ora_num_threads(void) { return omp_get_max_threads(); }
}

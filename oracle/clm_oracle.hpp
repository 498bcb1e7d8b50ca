// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// CPU restatement (C++17) of the cutoff-pair hot path of m3g/CellListMap.jl
// v0.10.4-DEV.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// Parity status: the reference is pure Julia and no Julia toolchain exists in
// this environment, so the reference itself cannot be executed here.  This
// restatement is pinned against the reference's own golden vectors (NAMD LJ
// energies of 8 DCD frames, argon doctest sums / neighbour-list sizes, grid
// KATs, nextfloat/prevfloat boundary KATs; see tests/test_oracle_golden.py)
// and against its own independent O(N^2) twin of the reference's test oracle
// map_naive!.  Bit-level parity of StaticArrays' closed-form `M \ x` is NOT
// pinned by the reference (its tests use isapprox); the op order chosen here is
// stated next to each function.
//
// Every function cites the reference file:line (relative to /root/reference/src)
// that it follows.  All arithmetic is carried out in T (float or double) exactly
// as the reference does for Float32 / Float64; compile with -ffp-contract=off.
// =============================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>

namespace ora {

enum CellType { ORTHO = 0, TRICLINIC = 1, NONPERIODIC = 2 };

template <class T, int N> struct Vec {
    T v[N];
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
};
// column-major N x N, element (r,c) at m[r + N*c]; columns = lattice vectors
template <class T, int N> struct Mat {
    T m[N * N];
    T& operator()(int r, int c) { return m[r + N * c]; }
    const T& operator()(int r, int c) const { return m[r + N * c]; }
};

template <class T, int N> static inline Mat<T, N> identity() {
    Mat<T, N> I;
    for (int i = 0; i < N * N; ++i) I.m[i] = T(0);
    for (int i = 0; i < N; ++i) I(i, i) = T(1);
    return I;
}

// StaticArrays matrix*vector: each row is a left fold  (a1*b1 + a2*b2) + a3*b3
template <class T, int N> static inline Vec<T, N> matvec(const Mat<T, N>& A, const Vec<T, N>& b) {
    Vec<T, N> r;
    for (int i = 0; i < N; ++i) {
        T s = A(i, 0) * b[0];
        for (int k = 1; k < N; ++k) s = s + A(i, k) * b[k];
        r[i] = s;
    }
    return r;
}
// StaticArrays matrix*matrix (unrolled, left fold over k)
template <class T, int N> static inline Mat<T, N> matmul(const Mat<T, N>& A, const Mat<T, N>& B) {
    Mat<T, N> C;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
            T s = A(i, 0) * B(0, j);
            for (int k = 1; k < N; ++k) s = s + A(i, k) * B(k, j);
            C(i, j) = s;
        }
    return C;
}
template <class T> static inline void cross3(const T* a, const T* b, T* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
template <class T, int N> static inline T dotn(const T* a, const T* b) {
    T s = a[0] * b[0];
    for (int k = 1; k < N; ++k) s = s + a[k] * b[k];
    return s;
}
template <class T, int N> static inline T norm2n(const T* a) { return dotn<T, N>(a, a); }

// det / solve / inv: StaticArrays closed forms for 2x2 and 3x3 (third party, compat
// "1.9.15", Project.toml:37; not vendored).  3x3 det = x0 . (x1 x x2) over columns.
template <class T> static inline T det(const Mat<T, 2>& A) { return A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0); }
template <class T> static inline T det(const Mat<T, 3>& A) {
    T c[3];
    cross3(&A.m[3], &A.m[6], c);
    return dotn<T, 3>(&A.m[0], c);
}
// M \ x  (call site: internals/CellOperations.jl:62)
template <class T> static inline Vec<T, 2> solve(const Mat<T, 2>& a, const Vec<T, 2>& b) {
    T d = det(a);
    Vec<T, 2> r;
    r[0] = (a(1, 1) * b[0] - a(0, 1) * b[1]) / d;
    r[1] = (a(0, 0) * b[1] - a(1, 0) * b[0]) / d;
    return r;
}
template <class T> static inline Vec<T, 3> solve(const Mat<T, 3>& a, const Vec<T, 3>& b) {
    T d = det(a);
    Vec<T, 3> r;
    r[0] = ((a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * b[0] + (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * b[1] +
            (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * b[2]) / d;
    r[1] = ((a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * b[0] + (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * b[1] +
            (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * b[2]) / d;
    r[2] = ((a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * b[0] + (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * b[1] +
            (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * b[2]) / d;
    return r;
}
// inv(rotation)  (call site: internals/Box.jl:261)
template <class T> static inline Mat<T, 2> inverse(const Mat<T, 2>& A) {
    T idet = T(1) / det(A);
    Mat<T, 2> R;
    R.m[0] = A.m[3] * idet;
    R.m[1] = -(A.m[1] * idet);
    R.m[2] = -(A.m[2] * idet);
    R.m[3] = A.m[0] * idet;
    return R;
}
template <class T> static inline Mat<T, 3> inverse(const Mat<T, 3>& A) {
    T x0[3] = {A.m[0], A.m[1], A.m[2]}, x1[3] = {A.m[3], A.m[4], A.m[5]}, x2[3] = {A.m[6], A.m[7], A.m[8]};
    T y0[3], y1[3], y2[3];
    cross3(x1, x2, y0);
    T d = dotn<T, 3>(x0, y0);
    for (int k = 0; k < 3; ++k) { x0[k] = x0[k] / d; y0[k] = y0[k] / d; }
    cross3(x2, x0, y1);
    cross3(x0, x1, y2);
    Mat<T, 3> R;
    R.m[0] = y0[0]; R.m[1] = y1[0]; R.m[2] = y2[0];
    R.m[3] = y0[1]; R.m[4] = y1[1]; R.m[5] = y2[1];
    R.m[6] = y0[2]; R.m[7] = y1[2]; R.m[8] = y2[2];
    return R;
}

// ---------------------------------------------------------------------------
// Box  (internals/Box.jl:84-96)
// ---------------------------------------------------------------------------
template <class T, int N> struct Box {
    int cell_type = ORTHO;
    Mat<T, N> input_unit_cell, aligned_unit_cell, rotation, inv_rotation;
    int lcell = 1;
    int64_t nc[N];
    T cutoff, cutoff_sqr;
    Vec<T, N> cb_min, cb_max, cell_size, origin;
};

// align_cell, 2-D  (internals/CellOperations.jl:353-375)
template <class T> static inline void align_cell(const Mat<T, 2>& min_, Mat<T, 2>& mout, Mat<T, 2>& R) {
    Mat<T, 2> m = min_;
    const T* a = &m.m[0];
    const T* b = &m.m[2];
    if (std::sqrt(norm2n<T, 2>(b)) > std::sqrt(norm2n<T, 2>(a))) a = b;
    if (a[1] == T(0)) {  // `a[y] ≈ zero(T)` is an exact-zero test (isapprox with atol = 0)
        R = identity<T, 2>();
    } else {
        T na = std::sqrt(norm2n<T, 2>(a));
        T sint = -na / (a[0] * a[0] / a[1] + a[1]);
        T cost = -a[0] * sint / a[1];
        R(0, 0) = cost; R(0, 1) = -sint;
        R(1, 0) = sint; R(1, 1) = cost;
        m = matmul(R, m);
    }
    mout = m;
}
// align_cell, 3-D  (internals/CellOperations.jl:377-423)
template <class T> static inline void align_cell(const Mat<T, 3>& min_, Mat<T, 3>& mout, Mat<T, 3>& R) {
    Mat<T, 3> m = min_;
    T n[3] = {norm2n<T, 3>(&m.m[0]), norm2n<T, 3>(&m.m[3]), norm2n<T, 3>(&m.m[6])};
    static const int comb[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    int ia = 0;
    for (int c = 0; c < 6; ++c) {
        if (n[comb[c][0]] >= n[comb[c][1]] && n[comb[c][1]] >= n[comb[c][2]]) { ia = comb[c][0]; break; }
    }
    const T* a = &m.m[3 * ia];
    T na = std::sqrt(norm2n<T, 3>(a));
    T a1[3] = {a[0] / na, a[1] / na, a[2] / na};
    T v[3] = {T(0), a1[2], -a1[1]};  // a1 x i
    Mat<T, 3> R1;
    if (norm2n<T, 3>(v) == T(0)) {
        R1 = identity<T, 3>();
    } else {
        Mat<T, 3> vs;
        vs(0, 0) = T(0);  vs(0, 1) = -v[2]; vs(0, 2) = v[1];
        vs(1, 0) = v[2];  vs(1, 1) = T(0);  vs(1, 2) = -v[0];
        vs(2, 0) = -v[1]; vs(2, 1) = v[0];  vs(2, 2) = T(0);
        Mat<T, 3> vs2 = matmul(vs, vs);
        T s = T(1) / (T(1) + a1[0]);
        Mat<T, 3> I = identity<T, 3>();
        for (int k = 0; k < 9; ++k) R1.m[k] = (I.m[k] + vs.m[k]) + vs2.m[k] * s;
    }
    m = matmul(R1, m);
    // rotation about x computed from column 2 of the rotated matrix (follow the code, :407)
    T x = m(0, 1), y = m(1, 1), z = m(2, 1);
    Mat<T, 3> R2;
    if ((y * y + z * z) == T(0)) {
        R2 = identity<T, 3>();
    } else {
        T b = std::sqrt(norm2n<T, 3>(&m.m[3]) - x * x);
        T sint = -z * b / (y * y + z * z);
        T cost = std::sqrt(T(1) - sint * sint);
        R2 = identity<T, 3>();
        R2(1, 1) = cost; R2(1, 2) = -sint;
        R2(2, 1) = sint; R2(2, 2) = cost;
    }
    m = matmul(R2, m);
    mout = m;
    R = matmul(R2, R1);
}

// cell_vertices / cell_limits  (internals/CellOperations.jl:431-479)
template <class T> static inline void cell_limits(const Mat<T, 2>& m, Vec<T, 2>& lo, Vec<T, 2>& hi) {
    T vtx[4][2] = {{T(0), T(0)}, {m.m[0], m.m[1]}, {m.m[0] + m.m[2], m.m[1] + m.m[3]}, {m.m[2], m.m[3]}};
    for (int j = 0; j < 2; ++j) lo[j] = hi[j] = vtx[0][j];
    for (int k = 1; k < 4; ++k)
        for (int j = 0; j < 2; ++j) { lo[j] = std::min(lo[j], vtx[k][j]); hi[j] = std::max(hi[j], vtx[k][j]); }
}
template <class T> static inline void cell_limits(const Mat<T, 3>& m, Vec<T, 3>& lo, Vec<T, 3>& hi) {
    T vtx[8][3];
    for (int j = 0; j < 3; ++j) {
        T c1 = m.m[j], c2 = m.m[3 + j], c3 = m.m[6 + j];
        vtx[0][j] = T(0); vtx[1][j] = c1; vtx[2][j] = c1 + c2; vtx[3][j] = c2;
        vtx[4][j] = c1 + c3; vtx[5][j] = c3; vtx[6][j] = c2 + c3; vtx[7][j] = (c1 + c2) + c3;
    }
    for (int j = 0; j < 3; ++j) lo[j] = hi[j] = vtx[0][j];
    for (int k = 1; k < 8; ++k)
        for (int j = 0; j < 3; ++j) { lo[j] = std::min(lo[j], vtx[k][j]); hi[j] = std::max(hi[j], vtx[k][j]); }
}

// check_unit_cell  (internals/Box.jl:579-636)
template <class T> static inline bool check_unit_cell(const Mat<T, 3>& M, T cutoff) {
    const T *a = &M.m[0], *b = &M.m[3], *c = &M.m[6];
    T bc[3], ab[3], ca[3];
    cross3(b, c, bc);
    T nb = std::sqrt(norm2n<T, 3>(bc));
    for (int k = 0; k < 3; ++k) bc[k] = bc[k] / nb;
    T aproj = dotn<T, 3>(a, bc);
    cross3(a, b, ab);
    nb = std::sqrt(norm2n<T, 3>(ab));
    for (int k = 0; k < 3; ++k) ab[k] = ab[k] / nb;
    T cproj = dotn<T, 3>(c, ab);
    cross3(c, a, ca);
    nb = std::sqrt(norm2n<T, 3>(ca));
    for (int k = 0; k < 3; ++k) ca[k] = ca[k] / nb;
    T bproj = dotn<T, 3>(b, ca);
    T two = T(2) * cutoff;
    return !((aproj <= two) || (bproj <= two) || (cproj <= two));
}
template <class T> static inline bool check_unit_cell(const Mat<T, 2>& M, T cutoff) {
    const T *a = &M.m[0], *b = &M.m[2];
    T na = std::sqrt(norm2n<T, 2>(a)), nb = std::sqrt(norm2n<T, 2>(b));
    T i[2] = {a[0] / na, a[1] / na};
    T di = dotn<T, 2>(b, i);
    T bproj = std::sqrt(norm2n<T, 2>(b) - di * di);
    T j[2] = {b[0] / nb, b[1] / nb};
    T dj = dotn<T, 2>(a, j);
    T aproj = std::sqrt(norm2n<T, 2>(a) - dj * dj);
    T two = T(2) * cutoff;
    return !((aproj <= two) || (bproj <= two));
}

// _construct_box  (internals/Box.jl:240-270) with _compute_nc_and_cell_size (:209-220)
// returns false when check_unit_cell fails (reference throws ArgumentError, :243)
template <class T, int N>
static inline bool construct_box(Box<T, N>& box, int cell_type, const Mat<T, N>& input, T cutoff, int lcell,
                                 const Vec<T, N>& origin) {
    box.cell_type = cell_type;
    box.input_unit_cell = input;
    box.lcell = lcell;
    box.cutoff = cutoff;
    box.origin = origin;
    if (cell_type == TRICLINIC) {
        align_cell(input, box.aligned_unit_cell, box.rotation);
    } else {
        box.aligned_unit_cell = input;
        box.rotation = identity<T, N>();
    }
    if (!check_unit_cell(box.aligned_unit_cell, cutoff)) return false;
    Vec<T, N> lo, hi;
    cell_limits(box.aligned_unit_cell, lo, hi);
    T side = cutoff / T(lcell);
    for (int i = 0; i < N; ++i) {
        int64_t nci;
        if (cell_type == TRICLINIC) {
            nci = (int64_t)std::ceil((hi[i] - lo[i]) / side);
            box.cell_size[i] = side;
        } else {
            nci = (int64_t)std::floor((hi[i] - lo[i]) / side);
            box.cell_size[i] = (hi[i] - lo[i]) / T(nci);
        }
        box.nc[i] = nci + 2 * lcell + 1;
    }
    for (int i = 0; i < N; ++i) {
        T t = T(lcell) * box.cell_size[i];
        box.cb_min[i] = (lo[i] + origin[i]) - t;
        box.cb_max[i] = (hi[i] + origin[i]) + t;
    }
    box.cutoff_sqr = cutoff * cutoff;
    box.inv_rotation = inverse(box.rotation);
    return true;
}

// ---------------------------------------------------------------------------
// Wrapping  (internals/CellOperations.jl:30, :56-66, :91-94, :102-127)
// ---------------------------------------------------------------------------
template <class T, int N> static inline Vec<T, N> wrap_cell_fraction(const Vec<T, N>& x, const Mat<T, N>& M) {
    Vec<T, N> p = solve(M, x);
    for (int i = 0; i < N; ++i) {
        p[i] = p[i] - std::floor(p[i]);  // fastmod1
        if (p[i] == T(1)) p[i] = T(0);
    }
    return p;
}
template <class T, int N> static inline Vec<T, N> wrap_to_first(const Vec<T, N>& x, const Mat<T, N>& M) {
    return matvec(M, wrap_cell_fraction(x, M));
}
// Julia mod(x, y) for floats with y > 0
template <class T> static inline T jl_mod(T x, T y) {
    T r = std::fmod(x, y);
    if (r == T(0)) return std::copysign(r, y);
    if ((r > 0) != (y > 0)) r = r + y;  // literal: may round to y itself; the caller's `>= 1/2` branch absorbs it
    return r;
}
// wrap_relative_to with a matrix (CellOperations.jl:102-111) -- the naive oracle's PBC
template <class T, int N>
static inline Vec<T, N> wrap_relative_to(const Vec<T, N>& x, const Vec<T, N>& xref, const Mat<T, N>& M) {
    Vec<T, N> xf = wrap_cell_fraction(x, M), rf = wrap_cell_fraction(xref, M), xw;
    for (int i = 0; i < N; ++i) {  // sides-version with unit sides (:120-127)
        T w = jl_mod(xf[i] - rf[i], T(1));
        if (w >= T(1) / T(2)) w = w - T(1);
        xw[i] = (w + rf[i]) - rf[i];
    }
    Vec<T, N> r = matvec(M, xw);
    for (int i = 0; i < N; ++i) r[i] = r[i] + xref[i];
    return r;
}

// ---------------------------------------------------------------------------
// Cell lists  (internals/CellLists.jl:19-23, :67-82, :115-137)
// ---------------------------------------------------------------------------
template <class T, int N> struct Particle {
    int64_t index;  // 1-based original index
    bool real;
    Vec<T, N> x;
};
template <class T, int N> struct Cell {
    int64_t linear_index = 0;  // 1-based
    int64_t cart[N];
    Vec<T, N> center;
    bool contains_real = false;
    std::vector<Particle<T, N>> particles;
};
template <class T, int N> struct CellList {
    int64_t n_real_particles = 0, number_of_cells = 0, n_particles = 0;
    std::vector<int64_t> cell_indices;       // linear (1-based) -> slot+1 in cells (0 = empty)
    std::vector<int64_t> cell_indices_real;  // slots (0-based) of cells with real particles, creation order
    std::vector<Cell<T, N>> cells;
};

// cell_linear_index (CellOperations.jl:256-257): column major, 1-based
template <int N> static inline int64_t linear_index(const int64_t* nc, const int64_t* c) {
    int64_t li = 0, stride = 1;
    for (int i = 0; i < N; ++i) { li += (c[i] - 1) * stride; stride *= nc[i]; }
    return li + 1;
}

template <class T, int N>
static inline void add_particle_to_celllist(int64_t ip, const Vec<T, N>& x, const Box<T, N>& box, CellList<T, N>& cl,
                                            bool real_particle) {
    // CellLists.jl:983-1057
    cl.n_particles += 1;
    int64_t c[N];
    for (int i = 0; i < N; ++i) {  // particle_cell, Box.jl:499-508
        T xi = (x[i] - box.cb_min[i]) / box.cell_size[i];
        c[i] = (int64_t)std::floor(xi) + 1;
    }
    if (real_particle) {  // real_particle_border_case, CellLists.jl:956-967
        for (int i = 0; i < N; ++i) {
            if (c[i] == box.lcell) c[i] += 1;
            if (c[i] == box.nc[i] - box.lcell + 1) c[i] -= 1;
        }
    }
    for (int i = 0; i < N; ++i)
        if (c[i] < 1 || c[i] > box.nc[i]) throw std::out_of_range("particle outside the computing grid (BoundsError in the reference)");
    int64_t li = linear_index<N>(box.nc, c);
    int64_t slot = cl.cell_indices[li - 1];
    if (slot == 0) {
        cl.cells.emplace_back();
        slot = (int64_t)cl.cells.size();
        cl.cell_indices[li - 1] = slot;
        Cell<T, N>& cell = cl.cells.back();
        cell.linear_index = li;
        for (int i = 0; i < N; ++i) {
            cell.cart[i] = c[i];
            // cell_center, Box.jl:517-525
            cell.center[i] = (box.cb_min[i] + box.cell_size[i] * T(c[i])) - box.cell_size[i] / T(2);
        }
    }
    Cell<T, N>& cell = cl.cells[slot - 1];
    if (real_particle && !cell.contains_real) {
        cell.contains_real = true;
        cl.cell_indices_real.push_back(slot - 1);
    }
    cell.particles.push_back(Particle<T, N>{ip, real_particle, x});
}

// UpdateCellList! serial path = the semantic definition
// (CellLists.jl:755-757, add_particles! :940-950, replicate_particle! Box.jl:556-566,
//  non-periodic add_particles! NonPeriodicCells.jl:63-70)
template <class T, int N>
static inline void build_cell_list(const T* xyz, int64_t n, const Box<T, N>& box, CellList<T, N>& cl) {
    cl = CellList<T, N>();
    cl.n_real_particles = n;
    int64_t ncells = 1;
    for (int i = 0; i < N; ++i) ncells *= box.nc[i];
    cl.number_of_cells = ncells;
    cl.cell_indices.assign((size_t)ncells, 0);
    const int nimg = (N == 2) ? 9 : 27;
    for (int64_t ip = 0; ip < n; ++ip) {
        Vec<T, N> p;
        for (int i = 0; i < N; ++i) p[i] = xyz[ip * N + i];
        if (box.cell_type == NONPERIODIC) {
            add_particle_to_celllist(ip + 1, p, box, cl, true);
            continue;
        }
        p = matvec(box.rotation, wrap_to_first(p, box.input_unit_cell));
        add_particle_to_celllist(ip + 1, p, box, cl, true);
        for (int img = 0; img < nimg; ++img) {  // Iterators.product(-1:1, ...): first index fastest
            int idx[3] = {img % 3 - 1, (img / 3) % 3 - 1, (img / 9) % 3 - 1};
            bool zero = true;
            for (int i = 0; i < N; ++i) zero = zero && (idx[i] == 0);
            if (zero) continue;
            Vec<T, N> tv;  // translation_image: x + M * SVector{N,Int}(indices)  (CellOperations.jl:138-139)
            for (int i = 0; i < N; ++i) tv[i] = T(idx[i]);
            Vec<T, N> sh = matvec(box.aligned_unit_cell, tv), q;
            bool inbox = true;
            for (int i = 0; i < N; ++i) {
                q[i] = p[i] + sh[i];
                if (!(box.cb_min[i] <= q[i] && q[i] < box.cb_max[i])) inbox = false;  // in_computing_box, Box.jl:536-546
            }
            if (inbox) add_particle_to_celllist(ip + 1, q, box, cl, false);
        }
    }
}

// ---------------------------------------------------------------------------
// Pair sweep  (internals/self.jl:106-184, cross.jl:81-129, vicinal_cells.jl:4-75,
//              NonPeriodicCells.jl:281-352, auxiliary_functions.jl:66-111)
// ---------------------------------------------------------------------------
template <class T, int N> struct Pair {  // NeighborPair (API/NeighborPair.jl:19-33)
    int64_t i, j;
    Vec<T, N> x, y;
    T d2;
};
template <class T, int N> struct Projected {
    int64_t index;
    T xproj;
    Vec<T, N> x;
    bool real;
};

template <class T, int N> static inline T dist2(const Vec<T, N>& a, const Vec<T, N>& b) {
    // sum(abs2, a - b): left fold, unfused
    T d = a[0] - b[0];
    T s = d * d;
    for (int k = 1; k < N; ++k) { d = a[k] - b[k]; s = s + d * d; }
    return s;
}

// stencils (Box.jl:436-474): forward for ortho / non-periodic self, full otherwise
template <int N> static inline std::vector<std::vector<int>> make_stencil(int lcell, bool forward) {
    std::vector<std::vector<int>> st;
    const int l = lcell;
    if (N == 3) {
        if (forward) {
            for (int k = -l; k <= l; ++k) for (int j = -l; j <= l; ++j) for (int i = 1; i <= l; ++i) st.push_back({i, j, k});
            for (int k = -l; k <= l; ++k) for (int j = 1; j <= l; ++j) st.push_back({0, j, k});
            for (int k = 1; k <= l; ++k) st.push_back({0, 0, k});
        } else {
            for (int k = -l; k <= l; ++k) for (int j = -l; j <= l; ++j) for (int i = -l; i <= l; ++i)
                if (i || j || k) st.push_back({i, j, k});
        }
    } else {
        if (forward) {
            for (int j = -l; j <= l; ++j) for (int i = 1; i <= l; ++i) st.push_back({i, j});
            for (int j = 1; j <= l; ++j) st.push_back({0, j});
        } else {
            for (int j = -l; j <= l; ++j) for (int i = -l; i <= l; ++i) if (i || j) st.push_back({i, j});
        }
    }
    return st;
}

enum SweepKind { SELF = 0, CROSS = 1 };

// One home cell against everything its stencil reaches.  `F` is callable as f(const Pair&).
// `use_projection` switches the reference's projection/partition pre-filter (pure pruning).
template <class T, int N, class F>
static inline void inner_loop(const Box<T, N>& box, const Cell<T, N>& ci, const CellList<T, N>& target, int kind,
                              const std::vector<std::vector<int>>& stencil, std::vector<Projected<T, N>>& scratch,
                              bool use_projection, F&& f) {
    const int ct = box.cell_type;
    const T rc2 = box.cutoff_sqr;
    auto emit = [&](const Particle<T, N>& pi, const Vec<T, N>& xj, int64_t jidx, T d2) {
        Pair<T, N> pr;
        pr.i = pi.index; pr.j = jidx; pr.d2 = d2;
        if (ct == ORTHO && kind == SELF) { pr.x = pi.x; pr.y = xj; }  // inv_rotation is the identity (self.jl:155)
        else { pr.x = matvec(box.inv_rotation, pi.x); pr.y = matvec(box.inv_rotation, xj); }
        f(pr);
    };
    // ---- current cell ----
    if (kind == SELF) {
        const auto& P = ci.particles;
        const int64_t np = (int64_t)P.size();
        if (ct == TRICLINIC) {  // self.jl:164-184
            for (int64_t i = 0; i < np; ++i) {
                if (!P[i].real) continue;
                for (int64_t j = 0; j < np; ++j) {
                    if (P[i].index >= P[j].index) continue;
                    T d2 = dist2(P[i].x, P[j].x);
                    if (d2 <= rc2) emit(P[i], P[j].x, P[j].index, d2);
                }
            }
        } else {  // ortho self.jl:143-161; non-periodic NonPeriodicCells.jl:287-303
            for (int64_t i = 0; i + 1 < np; ++i)
                for (int64_t j = i + 1; j < np; ++j) {
                    if (ct == ORTHO && !(P[i].real | P[j].real)) continue;
                    T d2 = dist2(P[i].x, P[j].x);
                    if (d2 <= rc2) emit(P[i], P[j].x, P[j].index, d2);
                }
        }
    } else {  // cross.jl:87-91, :111-129 ; NonPeriodicCells.jl:308-326
        int64_t slot = target.cell_indices[ci.linear_index - 1];
        if (slot != 0) {
            const auto& Q = target.cells[slot - 1].particles;
            for (const auto& pi : ci.particles) {
                if (ct != NONPERIODIC && !pi.real) continue;
                for (const auto& pj : Q) {
                    T d2 = dist2(pi.x, pj.x);
                    if (d2 <= rc2) emit(pi, pj.x, pj.index, d2);
                }
            }
        }
    }
    // ---- vicinal cells ----
    for (const auto& off : stencil) {
        int64_t c[N];
        bool inside = true;
        for (int i = 0; i < N; ++i) { c[i] = ci.cart[i] + off[i]; if (c[i] < 1 || c[i] > box.nc[i]) inside = false; }
        if (!inside) throw std::out_of_range("stencil leaves the grid (BoundsError in the reference)");
        int64_t slot = target.cell_indices[linear_index<N>(box.nc, c) - 1];
        if (slot == 0) continue;
        const Cell<T, N>& cj = target.cells[slot - 1];
        const bool skip = (kind == SELF);  // Val(true) only from self.jl:130
        auto test_pair = [&](const Particle<T, N>& pi, const Vec<T, N>& xj, int64_t jidx, bool jreal) {
            if (ct == ORTHO) { if (!(pi.real | jreal)) return; }                       // vicinal_cells.jl:33
            else if (ct == TRICLINIC) { if (skip && pi.index >= jidx) return; }        // vicinal_cells.jl:63-65
            T d2 = dist2(pi.x, xj);
            if (d2 <= rc2) emit(pi, xj, jidx, d2);
        };
        if (!use_projection) {
            for (const auto& pi : ci.particles) {
                if (ct == TRICLINIC && !pi.real) continue;  // vicinal_cells.jl:53
                for (const auto& pj : cj.particles) test_pair(pi, pj.x, pj.index, pj.real);
            }
            continue;
        }
        // _vicinal_cell_interactions! (vicinal_cells.jl:4-14) + project_particles! (auxiliary_functions.jl:91-111)
        Vec<T, N> dc;
        T s2 = T(0);
        for (int i = 0; i < N; ++i) { dc[i] = cj.center[i] - ci.center[i]; }
        s2 = norm2n<T, N>(dc.v);
        T dcn = std::sqrt(s2);
        for (int i = 0; i < N; ++i) dc[i] = dc[i] / dcn;
        T margin = (box.lcell == 1) ? (box.cutoff + dcn / T(2)) : (box.cutoff * (T(1) + std::sqrt(T(N)) / T(2)));
        if (scratch.size() < cj.particles.size()) scratch.resize(cj.particles.size());
        int64_t npp = 0;
        for (const auto& pj : cj.particles) {
            T d[N];
            for (int i = 0; i < N; ++i) d[i] = pj.x[i] - ci.center[i];
            T xproj = dotn<T, N>(d, dc.v);
            if (std::fabs(xproj) <= margin) scratch[npp++] = Projected<T, N>{pj.index, xproj, pj.x, pj.real};
        }
        if (npp == 0) continue;
        for (const auto& pi : ci.particles) {
            if (ct == TRICLINIC && !pi.real) continue;
            T d[N];
            for (int i = 0; i < N; ++i) d[i] = pi.x[i] - ci.center[i];
            T xproj = dotn<T, N>(d, dc.v);
            // partition! (auxiliary_functions.jl:66-77)
            int64_t iswap = 0;
            for (int64_t k = 0; k < npp; ++k) {
                if (std::fabs(scratch[k].xproj - xproj) <= box.cutoff) {
                    if (iswap != k) std::swap(scratch[iswap], scratch[k]);
                    ++iswap;
                }
            }
            for (int64_t k = 0; k < iswap; ++k) test_pair(pi, scratch[k].x, scratch[k].index, scratch[k].real);
        }
    }
}

// _pairwise! serial (self.jl:50-63, cross.jl:30-45)
template <class T, int N, class F>
static inline void map_pairwise_serial(const Box<T, N>& box, const CellList<T, N>& ref, const CellList<T, N>& target,
                                       int kind, bool use_projection, F&& f) {
    const bool forward = (kind == SELF) && (box.cell_type != TRICLINIC);
    auto stencil = make_stencil<N>(box.lcell, forward);
    std::vector<Projected<T, N>> scratch;
    for (int64_t s : ref.cell_indices_real) inner_loop(box, ref.cells[s], target, kind, stencil, scratch, use_projection, f);
}

// map_naive! (test/modules/Testing.jl:76-104): the reference's own O(N^2) test oracle.
template <class T, int N, class F>
static inline void map_naive_self(const T* x, int64_t n, const Box<T, N>& box, F&& f) {
    for (int64_t i = 0; i + 1 < n; ++i) {
        Vec<T, N> xi;
        for (int k = 0; k < N; ++k) xi[k] = x[i * N + k];
        for (int64_t j = i + 1; j < n; ++j) {
            Vec<T, N> xj;
            for (int k = 0; k < N; ++k) xj[k] = x[j * N + k];
            if (box.cell_type != NONPERIODIC) xj = wrap_relative_to(xj, xi, box.input_unit_cell);
            T d2 = dist2(xi, xj);
            if (d2 <= box.cutoff_sqr) f(Pair<T, N>{i + 1, j + 1, xi, xj, d2});
        }
    }
}
template <class T, int N, class F>
static inline void map_naive_cross(const T* x, int64_t nx, const T* y, int64_t ny, const Box<T, N>& box, F&& f) {
    for (int64_t i = 0; i < nx; ++i) {
        Vec<T, N> xi;
        for (int k = 0; k < N; ++k) xi[k] = x[i * N + k];
        for (int64_t j = 0; j < ny; ++j) {
            Vec<T, N> yj;
            for (int k = 0; k < N; ++k) yj[k] = y[j * N + k];
            if (box.cell_type != NONPERIODIC) yj = wrap_relative_to(yj, xi, box.input_unit_cell);
            T d2 = dist2(xi, yj);
            if (d2 <= box.cutoff_sqr) f(Pair<T, N>{i + 1, j + 1, xi, yj, d2});
        }
    }
}

}  // namespace ora

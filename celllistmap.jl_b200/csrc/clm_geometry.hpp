// Host-side box geometry of the B200 engine: what `Box(...)` computes in the reference
// (src/internals/Box.jl:191-270, :330-335, :374-377; src/internals/CellOperations.jl:337-479),
// evaluated in the handle's precision T so that every number handed to the kernels carries the same
// roundings the reference's Float32 / Float64 code would produce.  Matrices are stored row-major
// here (a[r][c]); the ABI converts from/to Julia's column-major layout.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include "clm_common.cuh"
#include "../../include/clm_b200.h"

namespace clm {

template <class T> struct HostBox {
    int dim = 3, cell_type = CLM_ORTHORHOMBIC, lcell = 1;
    T in[3][3], al[3][3], rot[3][3], irot[3][3];
    T cutoff = 0, cutoff_sqr = 0;
    T cb_min[3], cb_max[3], cs[3], origin[3];
    int64_t nc[3];
    bool valid = false;
};

namespace geo {
template <class T> inline void eye(T a[3][3]) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = (r == c) ? T(1) : T(0); }
template <class T> inline void copy(const T a[3][3], T b[3][3]) { std::memcpy(b, a, sizeof(T) * 9); }
// product of n x n blocks, each element a left fold a_i1*b_1j + a_i2*b_2j (+ a_i3*b_3j)
template <class T> inline void mul(int n, const T a[3][3], const T b[3][3], T out[3][3]) {
    T t[3][3];
    eye(t);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            T s = a[r][0] * b[0][c];
            for (int k = 1; k < n; ++k) s = s + a[r][k] * b[k][c];
            t[r][c] = s;
        }
    copy(t, out);
}
template <class T> inline T sq3(T x, T y, T z) { return (x * x + y * y) + z * z; }
template <class T> inline void cross(const T a[3], const T b[3], T o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
template <class T> inline T dot3(const T a[3], const T b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
template <class T> inline void column(const T a[3][3], int c, T o[3]) { for (int r = 0; r < 3; ++r) o[r] = a[r][c]; }

// rotation that puts the longest lattice vector on +x (2-D), CellOperations.jl:353-375
template <class T> inline void align2(T m[3][3], T R[3][3]) {
    eye(R);
    T a[2] = {m[0][0], m[1][0]}, b[2] = {m[0][1], m[1][1]};
    if (std::sqrt(b[0] * b[0] + b[1] * b[1]) > std::sqrt(a[0] * a[0] + a[1] * a[1])) { a[0] = b[0]; a[1] = b[1]; }
    if (a[1] == T(0)) return;
    const T na = std::sqrt(a[0] * a[0] + a[1] * a[1]);
    const T s = -na / (a[0] * a[0] / a[1] + a[1]);
    const T c = -a[0] * s / a[1];
    R[0][0] = c; R[0][1] = -s; R[1][0] = s; R[1][1] = c;
    mul(2, R, m, m);
}
// 3-D alignment, CellOperations.jl:377-423: Rodrigues rotation of the longest vector onto x, then a
// rotation about x derived from the SECOND COLUMN of the rotated matrix (as the code does).
template <class T> inline void align3(T m[3][3], T R[3][3]) {
    T col[3][3], n2[3];
    for (int c = 0; c < 3; ++c) { column(m, c, col[c]); n2[c] = sq3(col[c][0], col[c][1], col[c][2]); }
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    int first = 0;
    for (int p = 0; p < 6; ++p)
        if (n2[perm[p][0]] >= n2[perm[p][1]] && n2[perm[p][1]] >= n2[perm[p][2]]) { first = perm[p][0]; break; }
    const T na = std::sqrt(n2[first]);
    const T u[3] = {col[first][0] / na, col[first][1] / na, col[first][2] / na};
    const T v[3] = {T(0), u[2], -u[1]};
    T R1[3][3];
    eye(R1);
    if (sq3(v[0], v[1], v[2]) != T(0)) {
        T K[3][3] = {{T(0), -v[2], v[1]}, {v[2], T(0), -v[0]}, {-v[1], v[0], T(0)}}, K2[3][3];
        mul(3, K, K, K2);
        const T f = T(1) / (T(1) + u[0]);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R1[r][c] = (((r == c) ? T(1) : T(0)) + K[r][c]) + K2[r][c] * f;
    }
    mul(3, R1, m, m);
    const T x = m[0][1], y = m[1][1], z = m[2][1];
    T R2[3][3];
    eye(R2);
    if ((y * y + z * z) != T(0)) {
        const T b = std::sqrt(sq3(x, y, z) - x * x);
        const T s = -z * b / (y * y + z * z);
        const T c = std::sqrt(T(1) - s * s);
        R2[1][1] = c; R2[1][2] = -s; R2[2][1] = s; R2[2][2] = c;
    }
    mul(3, R2, m, m);
    mul(3, R2, R1, R);
}
// inverse of the rotation with the closed forms StaticArrays uses for 2x2 / 3x3
template <class T> inline void inverse(int n, const T a[3][3], T o[3][3]) {
    eye(o);
    if (n == 2) {
        const T idet = T(1) / (a[0][0] * a[1][1] - a[0][1] * a[1][0]);
        o[0][0] = a[1][1] * idet; o[1][0] = -(a[1][0] * idet);
        o[0][1] = -(a[0][1] * idet); o[1][1] = a[0][0] * idet;
        return;
    }
    T x0[3], x1[3], x2[3], y0[3], y1[3], y2[3];
    column(a, 0, x0); column(a, 1, x1); column(a, 2, x2);
    cross(x1, x2, y0);
    const T d = dot3(x0, y0);
    for (int k = 0; k < 3; ++k) { x0[k] = x0[k] / d; y0[k] = y0[k] / d; }
    cross(x2, x0, y1);
    cross(x0, x1, y2);
    for (int c = 0; c < 3; ++c) { o[0][c] = y0[c]; o[1][c] = y1[c]; o[2][c] = y2[c]; }
}
// minimum-image validity: every cell height must exceed 2*cutoff (Box.jl:579-636)
template <class T> inline bool heights_ok(int n, const T m[3][3], T cutoff) {
    const T lim = T(2) * cutoff;
    if (n == 2) {
        const T a[2] = {m[0][0], m[1][0]}, b[2] = {m[0][1], m[1][1]};
        const T na = std::sqrt(a[0] * a[0] + a[1] * a[1]), nb = std::sqrt(b[0] * b[0] + b[1] * b[1]);
        const T ba = b[0] * (a[0] / na) + b[1] * (a[1] / na);
        const T hb = std::sqrt((b[0] * b[0] + b[1] * b[1]) - ba * ba);
        const T ab = a[0] * (b[0] / nb) + a[1] * (b[1] / nb);
        const T ha = std::sqrt((a[0] * a[0] + a[1] * a[1]) - ab * ab);
        return !(ha <= lim || hb <= lim);
    }
    T a[3], b[3], c[3], nrm[3];
    column(m, 0, a); column(m, 1, b); column(m, 2, c);
    auto height = [&](const T* p, const T* q, const T* w) {
        cross(p, q, nrm);
        const T len = std::sqrt(sq3(nrm[0], nrm[1], nrm[2]));
        const T un[3] = {nrm[0] / len, nrm[1] / len, nrm[2] / len};
        return dot3(w, un);
    };
    const T ha = height(b, c, a), hc = height(a, b, c), hb = height(c, a, b);
    return !(ha <= lim || hb <= lim || hc <= lim);
}
}  // namespace geo

// _construct_box (Box.jl:240-270).  `cell` row-major dim x dim.  Returns a clm_status.
template <class T>
inline int make_box(HostBox<T>& B, int dim, int cell_type, const T cell[3][3], T cutoff, int lcell, const T origin[3],
                    std::string& err) {
    if (lcell < 1) { err = "lcell must be greater or equal to 1"; return CLM_ERR_ARGUMENT; }
    B.dim = dim; B.cell_type = cell_type; B.lcell = lcell; B.cutoff = cutoff; B.valid = false;
    geo::copy(cell, B.in);
    geo::copy(cell, B.al);
    geo::eye(B.rot);
    if (cell_type == CLM_TRICLINIC) { if (dim == 2) geo::align2(B.al, B.rot); else geo::align3(B.al, B.rot); }
    if (!geo::heights_ok(dim, B.al, cutoff)) {
        err = "Unit cell matrix does not satisfy required conditions. (UNIT CELL CHECK FAILED: distance between cell planes "
              "too small relative to cutoff: must be greater than 2*cutoff)";
        return CLM_ERR_UNIT_CELL;
    }
    // bounding box of the cell vertices (CellOperations.jl:431-479)
    T lo[3] = {T(0), T(0), T(0)}, hi[3] = {T(0), T(0), T(0)};
    for (int j = 0; j < dim; ++j) {
        const T c1 = B.al[j][0], c2 = B.al[j][1], c3 = (dim == 3) ? B.al[j][2] : T(0);
        T vt[8] = {T(0), c1, c1 + c2, c2, c1 + c3, c3, c2 + c3, (c1 + c2) + c3};
        const int nv = (dim == 3) ? 8 : 4;
        for (int k = 0; k < nv; ++k) { lo[j] = std::min(lo[j], vt[k]); hi[j] = std::max(hi[j], vt[k]); }
    }
    const T side = cutoff / T(lcell);
    for (int j = 0; j < 3; ++j) { B.nc[j] = 1; B.cs[j] = T(1); B.cb_min[j] = T(0); B.cb_max[j] = T(0); B.origin[j] = (j < dim) ? origin[j] : T(0); }
    for (int j = 0; j < dim; ++j) {
        int64_t inner;
        if (cell_type == CLM_TRICLINIC) { inner = (int64_t)std::ceil((hi[j] - lo[j]) / side); B.cs[j] = side; }
        else { inner = (int64_t)std::floor((hi[j] - lo[j]) / side); B.cs[j] = (hi[j] - lo[j]) / T(inner); }
        B.nc[j] = inner + 2 * lcell + 1;
        const T pad = T(lcell) * B.cs[j];
        B.cb_min[j] = (lo[j] + B.origin[j]) - pad;
        B.cb_max[j] = (hi[j] + B.origin[j]) + pad;
    }
    B.cutoff_sqr = cutoff * cutoff;
    geo::inverse(dim, B.rot, B.irot);
    B.valid = true;
    return CLM_OK;
}

// everything the kernels need, with the wrap's cofactors and the 3^N image shifts pre-evaluated in T
template <class T> inline void fill_geom(const HostBox<T>& B, GeomT<T>& G) {
    std::memset(&G, 0, sizeof(G));
    const int n = B.dim;
    const T(*a)[3] = B.in;
    if (n == 3) {
        // StaticArrays' 3x3 `\`: numerators built from 2x2 cofactors, det = col0 . (col1 x col2)
        G.cof[0] = a[1][1] * a[2][2] - a[1][2] * a[2][1]; G.cof[1] = a[0][2] * a[2][1] - a[0][1] * a[2][2]; G.cof[2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
        G.cof[3] = a[1][2] * a[2][0] - a[1][0] * a[2][2]; G.cof[4] = a[0][0] * a[2][2] - a[0][2] * a[2][0]; G.cof[5] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
        G.cof[6] = a[1][0] * a[2][1] - a[1][1] * a[2][0]; G.cof[7] = a[0][1] * a[2][0] - a[0][0] * a[2][1]; G.cof[8] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
        T c0[3], c1[3], c2[3], cr[3];
        geo::column(a, 0, c0); geo::column(a, 1, c1); geo::column(a, 2, c2);
        geo::cross(c1, c2, cr);
        G.det = geo::dot3(c0, cr);
    } else {
        // 2x2: ((a22*x1 - a12*x2)/d, (a11*x2 - a21*x1)/d); stored so that frac_k = (cof[3k]*x0 - cof[3k+1]*x1)/det
        G.cof[0] = a[1][1]; G.cof[1] = a[0][1];
        G.cof[3] = a[0][0]; G.cof[4] = a[1][0];  // frac_1 = (cof[3]*x1 - cof[4]*x0)/det
        G.det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    }
    bool rotated = false;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            G.m[3 * r + c] = B.in[r][c]; G.rot[3 * r + c] = B.rot[r][c]; G.inv_rot[3 * r + c] = B.irot[r][c];
            if (B.rot[r][c] != ((r == c) ? T(1) : T(0))) rotated = true;
        }
    const int nimg = (n == 3) ? 27 : 9;
    for (int img = 0; img < nimg; ++img) {
        const T idx[3] = {T(img % 3 - 1), T((img / 3) % 3 - 1), T((n == 3) ? (img / 9) % 3 - 1 : 0)};
        for (int r = 0; r < n; ++r) {
            T s = B.al[r][0] * idx[0];
            for (int k = 1; k < n; ++k) s = s + B.al[r][k] * idx[k];
            G.shift[img][r] = s;
        }
    }
    for (int j = 0; j < 3; ++j) { G.cb_min[j] = B.cb_min[j]; G.cb_max[j] = B.cb_max[j]; G.cs[j] = B.cs[j]; G.nc[j] = (int)B.nc[j]; }
    G.cutoff = B.cutoff; G.cutoff_sqr = B.cutoff_sqr; G.sub = 1; G.lcell = B.lcell; G.dim = n; G.cell_type = B.cell_type; G.rotated = rotated ? 1 : 0;
}

}  // namespace clm

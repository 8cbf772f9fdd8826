// Hierarchical H(curl) basis of the master tetrahedron, host + device.
//
// Reference: shape3DETet and its helpers (hvfem.py:319-1052): AffineTetrahedron (:1014-1052), the
// ancillary families AncEE / AncETri (:467-580), the scaled orthogonal polynomials PolyLegendre /
// PolyJacobi / PolyIJacobi / HomIJacobi (:583-788) and the orientation handling OrientE / OrientTri
// (:791-878).  The reference evaluates them per element, per Gauss point; here one evaluation per
// (order, entity variant, point) feeds (a) the reference-element contraction tables of the assembly
// kernels (pg_tables_init) and (b) point evaluations for the CSEM right-hand side and the receivers
// (pg_csem_rhs, pg_interpolate_fields).  Independent restatement of the mathematics (Fuentes, Keith,
// Demkowicz, Nagaraj 2015), same function order as the reference:
//   edges 0..5 (p functions each, increasing degree); faces 0..3 (p(p-1) each: for k = i+j ascending,
//   i ascending, the two families interleaved); interior (three families interleaved).
#pragma once
#include "pg_common.cuh"

namespace pg {
namespace fe {

#define PG_HD __host__ __device__ __forceinline__

struct V3 {
    double x, y, z;
};
PG_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
PG_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PG_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PG_HD V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
PG_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

constexpr int kMaxP = PG_MAX_ORDER;

// local topology (hvfem.py:150-161, :976-1011) and the six vertex permutations of a face (OrientTri)
__host__ __device__ inline int edge_vertex(int e, int k) {
    const int t[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    return t[e][k];
}
__host__ __device__ inline int face_vertex(int f, int k) {
    const int t[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};
    return t[f][k];
}
__host__ __device__ inline int face_perm(int o, int k) {
    const int t[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
    return t[o][k];
}
__host__ __device__ inline V3 grad_lambda(int v) {
    const double t[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    return V3{t[v][0], t[v][1], t[v][2]};
}

// shifted scaled Legendre P_0..P_n at (x; t)
__host__ __device__ inline void scaled_legendre(double x, double t, int n, double *P) {
    P[0] = 1.0;
    if (n >= 1) {
        const double y = 2.0 * x - t, tt = t * t;
        P[1] = y;
        for (int i = 1; i < n; ++i) P[i + 1] = ((2 * i + 1) * y * P[i] - i * tt * P[i - 1]) / (i + 1);
    }
}

// shifted scaled Jacobi P^alpha_0..P^alpha_n at (x; t)
__host__ __device__ inline void scaled_jacobi(double x, double t, int n, int alpha, double *P) {
    P[0] = 1.0;
    if (n >= 1) {
        const double y = 2.0 * x - t, tt = t * t, aa = (double)alpha * alpha;
        P[1] = y + alpha * x;
        for (int j = 2; j <= n; ++j) {
            const double a = 2.0 * j * (j + alpha) * (2 * j + alpha - 2);
            const double b = 2.0 * j + alpha - 1;
            const double c = (2.0 * j + alpha) * (2 * j + alpha - 2);
            const double d = 2.0 * (j + alpha - 1) * (j - 1) * (2 * j + alpha);
            P[j] = (b * (c * y + aa * t) * P[j - 1] - d * tt * P[j - 2]) / a;
        }
    }
}

// integrated scaled Jacobi L^alpha_1..L^alpha_n with dL/dx, dL/dt (index 0 <-> order 1)
__host__ __device__ inline void integrated_jacobi(double x, double t, int n, int alpha, double *L, double *dLdx,
                                                  double *dLdt) {
    double P[kMaxP + 1];
    scaled_jacobi(x, t, n, alpha, P);
    L[0] = x, dLdx[0] = P[0], dLdt[0] = 0.0;
    const double tt = t * t;
    for (int j = 2; j <= n; ++j) {
        const double t0 = 2.0 * j + alpha;
        const double a = (j + alpha) / ((t0 - 1) * t0);
        const double b = alpha / ((t0 - 2) * t0);
        const double c = (j - 1) / ((t0 - 2) * (t0 - 1));
        L[j - 1] = a * P[j] + b * t * P[j - 1] - c * tt * P[j - 2];
        dLdx[j - 1] = P[j - 1];
        dLdt[j - 1] = -(j - 1) * (P[j - 1] + t * P[j - 2]) / (t0 - 2);
    }
}

// E_i = P_i(s1; s0+s1) (s0 grad s1 - s1 grad s0), curl E_i = (i+2) P_i g0 x g1, i < nfun
__host__ __device__ inline void edge_family(double s0, double s1, V3 g0, V3 g1, int nfun, V3 *val, V3 *curl) {
    double P[kMaxP + 1];
    scaled_legendre(s1, s0 + s1, nfun > 0 ? nfun - 1 : 0, P);
    const V3 w = s0 * g1 - s1 * g0, cw = cross(g0, g1);
    for (int i = 0; i < nfun; ++i) {
        val[i] = P[i] * w;
        curl[i] = ((i + 2) * P[i]) * cw;
    }
}

// triangle ancillary functions of `order` for the triple (s, g): entry (i, j), i >= 0, j >= 1,
// i + j <= order - 1, stored at [i * kMaxP + j]
__host__ __device__ inline void triangle_family(const double s[3], const V3 g[3], int order, V3 *val, V3 *curl) {
    if (order < 2) return;
    V3 E[kMaxP], cE[kMaxP];
    edge_family(s[0], s[1], g[0], g[1], order - 1, E, cE);
    const double t = s[0] + s[1] + s[2];
    const V3 gsum = g[0] + g[1] + g[2];
    for (int i = 0; i < order - 1; ++i) {
        const int jmax = order - 1 - i;
        double L[kMaxP], dLdx[kMaxP], dLdt[kMaxP];
        integrated_jacobi(s[2], t, jmax, 2 * i + 1, L, dLdx, dLdt);
        for (int j = 1; j <= jmax; ++j) {
            const V3 gradL = dLdx[j - 1] * g[2] + dLdt[j - 1] * gsum;
            val[i * kMaxP + j] = L[j - 1] * E[i];
            curl[i * kMaxP + j] = L[j - 1] * cE[i] + cross(gradL, E[i]);
        }
    }
}

// the p(p-1) functions of local face f under orientation code o, storage order of hvfem.py:402-413
__host__ __device__ inline void face_functions(const double lam[4], int f, int o, int p, V3 *val, V3 *curl) {
    if (p < 2) return;
    int tri[3];
    for (int k = 0; k < 3; ++k) tri[k] = face_vertex(f, face_perm(o, k));
    for (int fam = 0; fam < 2; ++fam) {
        double s[3];
        V3 g[3];
        for (int k = 0; k < 3; ++k) {
            const int v = tri[(k + fam) % 3];
            s[k] = lam[v], g[k] = grad_lambda(v);
        }
        V3 tv[kMaxP * kMaxP], tc[kMaxP * kMaxP];
        triangle_family(s, g, p, tv, tc);
        int slot = fam;
        for (int k = 1; k < p; ++k)
            for (int i = 0; i < k; ++i) {
                val[slot] = tv[i * kMaxP + (k - i)], curl[slot] = tc[i * kMaxP + (k - i)];
                slot += 2;
            }
    }
}

// the p(p-1)(p-2)/2 interior functions, storage order of hvfem.py:429-453
__host__ __device__ inline void bubble_functions(const double lam[4], int p, V3 *val, V3 *curl) {
    if (p < 3) return;
    for (int fam = 0; fam < 3; ++fam) {
        const int a = fam % 4, b = (1 + fam) % 4, c = (2 + fam) % 4, d = (3 + fam) % 4;
        const double s[3] = {lam[a], lam[b], lam[c]};
        const V3 g[3] = {grad_lambda(a), grad_lambda(b), grad_lambda(c)};
        V3 tv[kMaxP * kMaxP], tc[kMaxP * kMaxP];
        triangle_family(s, g, p - 1, tv, tc);
        const V3 gd = grad_lambda(d);
        int slot = fam;
        for (int j = 2; j < p; ++j)
            for (int k = 1; k < j; ++k) {
                double L[kMaxP], dLdx[kMaxP], dLdt[kMaxP];
                integrated_jacobi(lam[d], 1.0, p - 2, 2 * k, L, dLdx, dLdt);
                const int q = j - k;
                for (int r = 0; r < k; ++r) {
                    const V3 v = tv[r * kMaxP + (k - r)], cv = tc[r * kMaxP + (k - r)];
                    const V3 gradL = dLdx[q - 1] * gd;
                    val[slot] = L[q - 1] * v;
                    curl[slot] = L[q - 1] * cv + cross(gradL, v);
                    slot += 3;
                }
            }
    }
}

__host__ __device__ inline void affine(const double xi[3], double lam[4]) {
    lam[0] = 1.0 - xi[0] - xi[1] - xi[2], lam[1] = xi[0], lam[2] = xi[1], lam[3] = xi[2];
}

// every function of the EXPANDED set (6 edges at orientation 0, 4 faces x 6 orientations, interior) at xi
__host__ __device__ inline void evaluate_expanded(int p, const double xi[3], V3 *N, V3 *C) {
    double lam[4];
    affine(xi, lam);
    for (int e = 0; e < 6; ++e) {
        const int a = edge_vertex(e, 0), b = edge_vertex(e, 1);
        edge_family(lam[a], lam[b], grad_lambda(a), grad_lambda(b), p, N + e * p, C + e * p);
    }
    const int nf = p * (p - 1);
    for (int f = 0; f < 4; ++f)
        for (int o = 0; o < 6; ++o) face_functions(lam, f, o, p, N + 6 * p + (f * 6 + o) * nf, C + 6 * p + (f * 6 + o) * nf);
    bubble_functions(lam, p, N + 6 * p + 24 * nf, C + 6 * p + 24 * nf);
}

// the n LOCAL functions of an element with orientation code `code` (pg_element_geometry) at xi:
// what shape3DETet(X, Nord, NoriE, NoriF) returns (hvfem.py:319-464)
__host__ __device__ inline void evaluate_local(int p, uint32_t code, const double xi[3], V3 *N, V3 *C) {
    double lam[4];
    affine(xi, lam);
    for (int e = 0; e < 6; ++e) {
        int a = edge_vertex(e, 0), b = edge_vertex(e, 1);
        if ((code >> e) & 1u) {  // OrientE: swap (s0, s1)
            const int t = a;
            a = b, b = t;
        }
        edge_family(lam[a], lam[b], grad_lambda(a), grad_lambda(b), p, N + e * p, C + e * p);
    }
    const int nf = p * (p - 1);
    for (int f = 0; f < 4; ++f)
        face_functions(lam, f, (int)((code >> (6 + 3 * f)) & 7u), p, N + 6 * p + f * nf, C + 6 * p + f * nf);
    bubble_functions(lam, p, N + 6 * p + 4 * nf, C + 6 * p + 4 * nf);
}

}  // namespace fe
}  // namespace pg

// Per-element kernels: geometry + orientation, element matrices, DOF connectivity.
// Reference call sites: solver.py:193-224, hvfem.py:15-316.
#include <stdarg.h>

#include "pg_common.cuh"

namespace pg {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------
// geometry + orientation: one thread per element
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) geometry_kernel(int64_t T, const double *__restrict__ nodes,
                                                       const int32_t *__restrict__ elemsN,
                                                       const int32_t *__restrict__ elemsE,
                                                       const int32_t *__restrict__ edgesNodes,
                                                       const int32_t *__restrict__ facesEdges,
                                                       const double *__restrict__ sigma, double *__restrict__ geo,
                                                       uint32_t *__restrict__ code) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= T) return;

    // ---- computeJacobian (hvfem.py:101-119): rows of J are x_i - x_0 ----
    double x[12];
    {
        const double2 *src = reinterpret_cast<const double2 *>(nodes + t * 12);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double2 v = __ldg(src + i);
            x[2 * i] = v.x;
            x[2 * i + 1] = v.y;
        }
    }
    double J[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) J[i][c] = x[3 * (i + 1) + c] - x[c];

    // adjugate (cofactor transposed) and signed determinant (hvfem.py:265: no abs())
    double A[3][3];
    A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
    double inv_det = 1.0 / det;

    const double sh = __ldg(sigma + 2 * t), sv = __ldg(sigma + 2 * t + 1);
    const double sg[3] = {sh, sh, sv};
    const int pa[6] = {0, 1, 2, 0, 0, 1}, pb[6] = {0, 1, 2, 1, 2, 2};
    double out[12];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        int a = pa[c], b = pb[c];
        // gK = J J^T / detJ   (curl_real = C_ref J / detJ, times detJ: hvfem.py:304,:314)
        out[c] = (J[a][0] * J[b][0] + J[a][1] * J[b][1] + J[a][2] * J[b][2]) * inv_det;
        // gM = detJ Jinv^T diag(s) Jinv with Jinv = A/det  (hvfem.py:292,:297,:313)
        out[6 + c] = (sg[0] * A[0][a] * A[0][b] + sg[1] * A[1][a] * A[1][b] + sg[2] * A[2][a] * A[2][b]) * inv_det;
    }
    {
        double2 *dst = reinterpret_cast<double2 *>(geo + t * 12);
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[i] = make_double2(out[2 * i], out[2 * i + 1]);
    }

    // ---- computeElementOrientation (hvfem.py:122-220) ----
    int nd[4], ed[6];
    {
        int4 v = __ldg(reinterpret_cast<const int4 *>(elemsN + t * 4));
        nd[0] = v.x, nd[1] = v.y, nd[2] = v.z, nd[3] = v.w;
        const int2 *e2 = reinterpret_cast<const int2 *>(elemsE + t * 6);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int2 w = __ldg(e2 + i);
            ed[2 * i] = w.x;
            ed[2 * i + 1] = w.y;
        }
    }
    const int ea[6] = {0, 1, 0, 0, 1, 2}, eb[6] = {1, 2, 2, 3, 3, 3};
    uint32_t cd = 0;
    {
        const int4 *en4 = reinterpret_cast<const int4 *>(edgesNodes + t * 12);
        int en[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int4 v = __ldg(en4 + i);
            en[4 * i] = v.x, en[4 * i + 1] = v.y, en[4 * i + 2] = v.z, en[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (nd[ea[i]] == en[2 * i + 1] && nd[eb[i]] == en[2 * i]) cd |= (1u << i);
    }
    {
        const int4 *fe4 = reinterpret_cast<const int4 *>(facesEdges + t * 12);
        int fe[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int4 v = __ldg(fe4 + i);
            fe[4 * i] = v.x, fe[4 * i + 1] = v.y, fe[4 * i + 2] = v.z, fe[4 * i + 3] = v.w;
        }
        const int l0[4] = {0, 0, 1, 2}, l1[4] = {1, 4, 5, 5};
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            int g0 = ed[l0[f]], g1 = ed[l1[f]];
            int k1 = 0, k2 = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (g0 == fe[3 * f + k]) k1 = k + 1;
                if (g1 == fe[3 * f + k]) k2 = k + 1;
            }
            int sel = k1 * 10 + k2;
            uint32_t o = sel == 31 ? 1u : sel == 23 ? 2u : sel == 32 ? 3u : sel == 13 ? 4u : sel == 21 ? 5u : 0u;
            cd |= o << (6 + 3 * f);
        }
    }
    code[t] = cd;
}

// ---------------------------------------------------------------------------
// element matrices, warp per element (CUDA-core path)
// ---------------------------------------------------------------------------
template <int P, int MODE>  // MODE 0: Me,Ke real ; MODE 1: Ae = K + i*scale*M complex
__global__ void __launch_bounds__(256) element_matrices_kernel(int64_t T, const double *__restrict__ geo,
                                                               const uint32_t *__restrict__ code,
                                                               const double *__restrict__ table, double scale,
                                                               double *__restrict__ Me, double *__restrict__ Ke,
                                                               double *__restrict__ Ae) {
    using O = Ord<P>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < T; t += nwarps) {
        double g[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) g[i] = __ldg(geo + t * 12 + i);
        const uint32_t cd = __ldg(code + t);
        for (int jk = lane; jk < O::n * O::n; jk += 32) {
            int j = jk / O::n, k = jk - j * O::n;
            int sj, dj, sk, dk;
            slot_of_local<P>(j, sj, dj);
            slot_of_local<P>(k, sk, dk);
            double s1, s2;
            int Jx = expanded_of_slot<P>(sj, dj, cd, s1);
            int Kx = expanded_of_slot<P>(sk, dk, cd, s2);
            double kk, mm;
            contract12(table + ((int64_t)Jx * O::nexp + Kx) * 12, g, kk, mm);
            double s = s1 * s2;
            kk *= s;
            mm *= s;
            int64_t o = t * (int64_t)(O::n * O::n) + jk;
            if (MODE == 0) {
                if (Ke) Ke[o] = kk;
                if (Me) Me[o] = mm;
            } else {
                reinterpret_cast<double2 *>(Ae)[o] = make_double2(kk, scale * mm);
            }
        }
    }
}

// computeConnectivityDOFS (hvfem.py:15-98): one thread per (element, local dof)
template <int P>
__global__ void __launch_bounds__(256) dofs_kernel(int64_t T, const int32_t *__restrict__ elemsE,
                                                   const int32_t *__restrict__ elemsF, int64_t nE, int64_t nF,
                                                   int32_t *__restrict__ dofs) {
    using O = Ord<P>;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= T * O::n) return;
    int64_t t = i / O::n;
    int k = (int)(i - t * O::n);
    int slot, d;
    slot_of_local<P>(k, slot, d);
    int64_t v;
    if (slot < 6)
        v = (int64_t)elemsE[t * 6 + slot] * O::ne + d;
    else if (slot < 10)
        v = nE * O::ne + (int64_t)elemsF[t * 4 + slot - 6] * O::nf + d;
    else
        v = nE * O::ne + nF * O::nf + t * O::nv + d;
    dofs[i] = (int32_t)v;
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_version(void) { return 100; }
const char *pg_last_error(void) { return pg::g_err; }
int pg_ndof_element(int p) { return ndof_element(p); }
int pg_ndof_edge(int p) { return ndof_edge(p); }
int pg_ndof_face(int p) { return ndof_face(p); }
int pg_ndof_volume(int p) { return ndof_volume(p); }
int pg_nexp(int p) { return 6 * p + 24 * ndof_face(p) + ndof_volume(p); }

int pg_element_geometry(int64_t T, const double *nodes, const int32_t *elemsN, const int32_t *elemsE,
                        const int32_t *edgesNodes, const int32_t *facesEdges, const double *sigma, double *geo,
                        uint32_t *code, void *stream) {
    PG_REQUIRE(T >= 0, PG_EINVAL, "pg_element_geometry: T < 0");
    if (T == 0) return PG_OK;
    PG_REQUIRE(nodes && elemsN && elemsE && edgesNodes && facesEdges && sigma && geo && code, PG_EINVAL,
               "pg_element_geometry: null pointer");
    int64_t blocks = (T + 127) / 128;
    geometry_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(T, nodes, elemsN, elemsE, edgesNodes,
                                                                        facesEdges, sigma, geo, code);
    PG_LAUNCH_OK();
    return PG_OK;
}

static int launch_element(int64_t T, int p, const double *geo, const uint32_t *code, const double *table,
                          double scale, double *Me, double *Ke, double *Ae, void *stream) {
    PG_REQUIRE(T >= 0, PG_EINVAL, "element matrices: T < 0");
    if (T == 0) return PG_OK;
    PG_REQUIRE(geo && code && table, PG_EINVAL, "element matrices: null pointer");
    int64_t blocks = (T + 7) / 8;  // 8 warps per block
    const int64_t cap = (int64_t)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    return dispatch_order(p, [&](auto o) {
        constexpr int P = decltype(o)::p;
        if (Ae)
            element_matrices_kernel<P, 1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
                T, geo, code, table, scale, nullptr, nullptr, Ae);
        else
            element_matrices_kernel<P, 0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
                T, geo, code, table, scale, Me, Ke, nullptr);
        PG_LAUNCH_OK();
        return PG_OK;
    });
}

int pg_element_matrices(int64_t T, int p, const double *geo, const uint32_t *code, const double *table,
                        double *Me, double *Ke, void *stream) {
    PG_REQUIRE(Me || Ke, PG_EINVAL, "pg_element_matrices: both outputs null");
    return launch_element(T, p, geo, code, table, 0.0, Me, Ke, nullptr, stream);
}

int pg_element_systems(int64_t T, int p, const double *geo, const uint32_t *code, const double *table,
                       double mass_scale, double *Ae, void *stream) {
    PG_REQUIRE(Ae, PG_EINVAL, "pg_element_systems: null output");
    return launch_element(T, p, geo, code, table, mass_scale, nullptr, nullptr, Ae, stream);
}

int pg_connectivity_dofs(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                         int64_t nFaces, int32_t *dofs, void *stream) {
    PG_REQUIRE(T >= 0, PG_EINVAL, "pg_connectivity_dofs: T < 0");
    if (T == 0) return PG_OK;
    PG_REQUIRE(elemsE && elemsF && dofs, PG_EINVAL, "pg_connectivity_dofs: null pointer");
    return dispatch_order(p, [&](auto o) {
        constexpr int P = decltype(o)::p;
        int64_t total = T * Ord<P>::n;
        int64_t N = nEdges * Ord<P>::ne + nFaces * Ord<P>::nf + T * Ord<P>::nv;
        PG_REQUIRE(N < 2147483647LL, PG_ERANGE, "pg_connectivity_dofs: %lld dofs exceed int32", (long long)N);
        int64_t blocks = (total + 255) / 256;
        dofs_kernel<P><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(T, elemsE, elemsF, nEdges, nFaces, dofs);
        PG_LAUNCH_OK();
        return PG_OK;
    });
}

}  // extern "C"

// Reference-element tables and point evaluations of the basis, on the device.
//
//   pg_tables_init          the contraction tables SK, SM of the assembly kernels (what basis.py builds on the
//                           host): T^{ab}_{JK} = int N_J^a N_K^b over the master tetrahedron, exact conical
//                           Gauss-Jacobi rule of degree 2p+1; replaces the per-element, per-Gauss-point
//                           evaluation of hvfem.py:270-314
//   pg_locate_points        containing element of each point (postprocessing.py:532-539, find_simplex)
//   pg_interpolate_fields   fieldInterpolator (postprocessing.py:479-616): E, H at the receivers
//   pg_csem_rhs             the dipole right-hand side (solver.py:247-316)
#include <math.h>

#include <algorithm>
#include <vector>

#include "pg_basis.cuh"

namespace pg {
namespace {

using fe::V3;

// ---- host: n-point Gauss rule on [0,1] for the weight (1-u)^alpha (Golub-Welsch, cyclic Jacobi sweeps) ----
void gauss_jacobi_01(int n, int alpha, double *x, double *w) {
    const double a = alpha, b = 0.0;
    std::vector<double> T(n * n, 0.0), V(n * n, 0.0);
    for (int k = 0; k < n; ++k) {
        T[k * n + k] = (k == 0) ? (b - a) / (a + b + 2.0) : (b * b - a * a) / ((2 * k + a + b) * (2 * k + a + b + 2.0));
        V[k * n + k] = 1.0;
        if (k >= 1) {
            const double kk = k;
            const double off = 2.0 / (2 * kk + a + b) *
                               sqrt(kk * (kk + a) * (kk + b) * (kk + a + b) / ((2 * kk + a + b - 1.0) * (2 * kk + a + b + 1.0)));
            T[(k - 1) * n + k] = T[k * n + (k - 1)] = off;
        }
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double offn = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) offn += T[i * n + j] * T[i * n + j];
        if (offn < 1e-60) break;
        for (int pi = 0; pi < n; ++pi)
            for (int q = pi + 1; q < n; ++q) {
                const double apq = T[pi * n + q];
                if (apq == 0.0) continue;
                const double theta = (T[q * n + q] - T[pi * n + pi]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {  // columns
                    const double kp = T[k * n + pi], kq = T[k * n + q];
                    T[k * n + pi] = c * kp - s * kq, T[k * n + q] = s * kp + c * kq;
                }
                for (int k = 0; k < n; ++k) {  // rows
                    const double pk = T[pi * n + k], qk = T[q * n + k];
                    T[pi * n + k] = c * pk - s * qk, T[q * n + k] = s * pk + c * qk;
                }
                for (int k = 0; k < n; ++k) {  // eigenvectors (columns of V)
                    const double kp = V[k * n + pi], kq = V[k * n + q];
                    V[k * n + pi] = c * kp - s * kq, V[k * n + q] = s * kp + c * kq;
                }
            }
    }
    const double mu0 = pow(2.0, a + b + 1) / (a + b + 1);
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    for (int i = 0; i < n; ++i)  // ascending nodes
        for (int j = i + 1; j < n; ++j)
            if (T[idx[j] * n + idx[j]] < T[idx[i] * n + idx[i]]) std::swap(idx[i], idx[j]);
    for (int i = 0; i < n; ++i) {
        const int k = idx[i];
        x[i] = (T[k * n + k] + 1.0) / 2.0;
        w[i] = mu0 * V[0 * n + k] * V[0 * n + k] / pow(2.0, a + 1);
    }
}

__global__ void eval_points_kernel(int p, int nexp, int nq, const double *__restrict__ pts, V3 *__restrict__ N,
                                   V3 *__restrict__ C) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nq) return;
    const double xi[3] = {pts[3 * g], pts[3 * g + 1], pts[3 * g + 2]};
    fe::evaluate_expanded(p, xi, N + (size_t)g * nexp, C + (size_t)g * nexp);
}

// one thread per expanded pair (J, K): table[J][K][0..5] = SK, [6..11] = SM, packed (00,11,22,01,02,12)
__global__ void __launch_bounds__(128) contract_kernel(int nexp, int nq, const double *__restrict__ wts,
                                                       const V3 *__restrict__ N, const V3 *__restrict__ C,
                                                       double *__restrict__ table) {
    const int64_t pair = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (pair >= (int64_t)nexp * nexp) return;
    const int J = (int)(pair / nexp), K = (int)(pair - (int64_t)J * nexp);
    double sk[6] = {0, 0, 0, 0, 0, 0}, sm[6] = {0, 0, 0, 0, 0, 0};
    for (int g = 0; g < nq; ++g) {
        const double w = wts[g];
        const V3 a = C[(size_t)g * nexp + J], b = C[(size_t)g * nexp + K];
        const V3 aw = w * a;
        sk[0] = fma(aw.x, b.x, sk[0]), sk[1] = fma(aw.y, b.y, sk[1]), sk[2] = fma(aw.z, b.z, sk[2]);
        sk[3] += aw.x * b.y + aw.y * b.x, sk[4] += aw.x * b.z + aw.z * b.x, sk[5] += aw.y * b.z + aw.z * b.y;
        const V3 c = N[(size_t)g * nexp + J], d = N[(size_t)g * nexp + K];
        const V3 cw = w * c;
        sm[0] = fma(cw.x, d.x, sm[0]), sm[1] = fma(cw.y, d.y, sm[1]), sm[2] = fma(cw.z, d.z, sm[2]);
        sm[3] += cw.x * d.y + cw.y * d.x, sm[4] += cw.x * d.z + cw.z * d.x, sm[5] += cw.y * d.z + cw.z * d.y;
    }
    double *o = table + pair * 12;
#pragma unroll
    for (int c = 0; c < 6; ++c) o[c] = sk[c], o[6 + c] = sm[c];
}

// ---- points ----------------------------------------------------------------------------------------------
struct ElemGeom {
    double x0[3], J[3][3], Ji[3][3], det;  // J rows = edge vectors (hvfem.py:112-114), Ji = J^-1
};

__device__ inline void element_geometry(const double *__restrict__ nodes12, ElemGeom &g) {
    for (int c = 0; c < 3; ++c) g.x0[c] = nodes12[c];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) g.J[i][c] = nodes12[3 * (i + 1) + c] - nodes12[c];
    const double(*J)[3] = g.J;
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
    g.det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) g.Ji[i][c] = A[i][c] / g.det;
}

// master coordinates of a physical point: x = x0 + J^T xi  ->  xi = J^-T (x - x0)  (hvfem.py:2347-2490)
__device__ inline void to_master(const ElemGeom &g, const double *pt, double xi[3]) {
    const double d[3] = {pt[0] - g.x0[0], pt[1] - g.x0[1], pt[2] - g.x0[2]};
    for (int i = 0; i < 3; ++i) xi[i] = g.Ji[0][i] * d[0] + g.Ji[1][i] * d[1] + g.Ji[2][i] * d[2];
}

// lowest-index element whose barycentric coordinates are all >= -tol (the oracle's locate_points)
__global__ void __launch_bounds__(256) locate_kernel(int64_t T, const double *__restrict__ nodes, int64_t npts,
                                                     const double *__restrict__ points, double tol,
                                                     unsigned long long *__restrict__ best) {
    const int64_t ipt = blockIdx.y;
    const double *pt = points + 3 * ipt;
    unsigned long long mine = ~0ull;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
        ElemGeom g;
        element_geometry(nodes + t * 12, g);
        double xi[3];
        to_master(g, pt, xi);
        const double l0 = 1.0 - xi[0] - xi[1] - xi[2];
        if (xi[0] >= -tol && xi[1] >= -tol && xi[2] >= -tol && l0 >= -tol) {
            mine = (unsigned long long)t;
            break;  // ascending t within a thread: first hit is this thread's lowest
        }
    }
    if (mine != ~0ull) atomicMin(best + ipt, mine);
}

__global__ void finish_locate_kernel(int64_t npts, const unsigned long long *__restrict__ best,
                                     int32_t *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < npts) out[i] = best[i] == ~0ull ? -1 : (int32_t)best[i];
}

__device__ inline int64_t dof_of_local(int p, int k, const int32_t *__restrict__ eE, const int32_t *__restrict__ eF,
                                       int64_t t, int64_t nE, int64_t nF) {
    const int ne = p, nf = p * (p - 1), nv = p * (p - 1) * (p - 2) / 2;
    if (k < 6 * ne) return (int64_t)eE[t * 6 + k / ne] * ne + k % ne;
    k -= 6 * ne;
    if (k < 4 * nf) return nE * ne + (int64_t)eF[t * 4 + k / nf] * nf + k % nf;
    k -= 4 * nf;
    return nE * ne + nF * nf + t * nv + k;
}

// one block per point; threads share the n local functions (evaluated by thread 0 into shared memory is
// not needed: npts is tiny, every thread of a warp evaluates nothing -- one thread per point)
__global__ void interpolate_kernel(int64_t npts, const double *__restrict__ points, const int32_t *__restrict__ pt_elem,
                                   int p, const double *__restrict__ nodes, const uint32_t *__restrict__ code,
                                   const int32_t *__restrict__ elemsE, const int32_t *__restrict__ elemsF, int64_t nE,
                                   int64_t nF, const int32_t *__restrict__ perm, const double2 *__restrict__ x,
                                   double omega_mu, double2 *__restrict__ fields, V3 *__restrict__ scratch) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const int n = ndof_element(p);
    double2 *f = fields + 6 * i;
    const int64_t t = pt_elem[i];
    if (t < 0) {
        for (int c = 0; c < 6; ++c) f[c] = make_double2(nan(""), nan(""));
        return;
    }
    ElemGeom g;
    element_geometry(nodes + t * 12, g);
    double xi[3];
    to_master(g, points + 3 * i, xi);
    V3 *N = scratch + (size_t)i * 2 * n, *C = N + n;
    fe::evaluate_local(p, code[t], xi, N, C);
    double2 E[3] = {{0, 0}, {0, 0}, {0, 0}}, H[3] = {{0, 0}, {0, 0}, {0, 0}};
    for (int k = 0; k < n; ++k) {
        int64_t d = dof_of_local(p, k, elemsE, elemsF, t, nE, nF);
        if (perm) d = perm[d];
        const double2 xv = x[d];
        // N_real = J^-1 N_ref (hvfem.py:292), curl_real = J^T C_ref / detJ (hvfem.py:304)
        const double nr[3] = {g.Ji[0][0] * N[k].x + g.Ji[0][1] * N[k].y + g.Ji[0][2] * N[k].z,
                              g.Ji[1][0] * N[k].x + g.Ji[1][1] * N[k].y + g.Ji[1][2] * N[k].z,
                              g.Ji[2][0] * N[k].x + g.Ji[2][1] * N[k].y + g.Ji[2][2] * N[k].z};
        const double cr[3] = {(g.J[0][0] * C[k].x + g.J[1][0] * C[k].y + g.J[2][0] * C[k].z) / g.det,
                              (g.J[0][1] * C[k].x + g.J[1][1] * C[k].y + g.J[2][1] * C[k].z) / g.det,
                              (g.J[0][2] * C[k].x + g.J[1][2] * C[k].y + g.J[2][2] * C[k].z) / g.det};
        for (int c = 0; c < 3; ++c) {
            E[c].x = fma(nr[c], xv.x, E[c].x), E[c].y = fma(nr[c], xv.y, E[c].y);
            H[c].x = fma(cr[c], xv.x, H[c].x), H[c].y = fma(cr[c], xv.y, H[c].y);
        }
    }
    for (int c = 0; c < 3; ++c) {
        f[c] = E[c];
        // H = curl E / (i omega mu):  (a + ib) / (i w) = (b - i a) / w
        f[3 + c] = make_double2(H[c].y / omega_mu, -H[c].x / omega_mu);
    }
}

// b[dof_j] += i omega mu (moment . N_j(x_src)) for the dofs of the source element (solver.py:255-316)
__global__ void csem_rhs_kernel(int p, int64_t t, const double *__restrict__ position, const double *__restrict__ moment,
                                const double *__restrict__ nodes, const uint32_t *__restrict__ code,
                                const int32_t *__restrict__ elemsE, const int32_t *__restrict__ elemsF, int64_t nE,
                                int64_t nF, const int32_t *__restrict__ perm, int64_t row_begin, int64_t local_rows,
                                double omega_mu, double2 *__restrict__ b, V3 *__restrict__ scratch) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int n = ndof_element(p);
    ElemGeom g;
    element_geometry(nodes + t * 12, g);
    double xi[3];
    to_master(g, position, xi);
    V3 *N = scratch, *C = scratch + n;
    fe::evaluate_local(p, code[t], xi, N, C);
    for (int k = 0; k < n; ++k) {
        const double nr[3] = {g.Ji[0][0] * N[k].x + g.Ji[0][1] * N[k].y + g.Ji[0][2] * N[k].z,
                              g.Ji[1][0] * N[k].x + g.Ji[1][1] * N[k].y + g.Ji[1][2] * N[k].z,
                              g.Ji[2][0] * N[k].x + g.Ji[2][1] * N[k].y + g.Ji[2][2] * N[k].z};
        const double s = moment[0] * nr[0] + moment[1] * nr[1] + moment[2] * nr[2];
        int64_t d = dof_of_local(p, k, elemsE, elemsF, t, nE, nF);
        if (perm) d = perm[d];
        d -= row_begin;
        if (d >= 0 && d < local_rows) b[d].y += omega_mu * s;  // i * omega mu * s
    }
}

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int pg_shape_functions_host(int p, uint32_t code, const double *xi_host, double *N_host, double *C_host) {
    PG_REQUIRE(p >= 1 && p <= PG_MAX_ORDER, PG_EINVAL, "pg_shape_functions_host: order %d", p);
    PG_REQUIRE(xi_host && N_host && C_host, PG_EINVAL, "pg_shape_functions_host: null pointer");
    static_assert(sizeof(fe::V3) == 3 * sizeof(double), "V3 must be three packed doubles");
    fe::evaluate_local(p, code, xi_host, reinterpret_cast<fe::V3 *>(N_host), reinterpret_cast<fe::V3 *>(C_host));
    return PG_OK;
}

int64_t pg_table_size(int p) {
    if (p < 1 || p > PG_MAX_ORDER) return 0;
    const int64_t nexp = pg_nexp(p);
    return nexp * nexp * 12;
}

int pg_tables_init(int p, double *table, void *stream) {
    PG_REQUIRE(p >= 1 && p <= PG_MAX_ORDER, PG_EINVAL, "pg_tables_init: order %d outside 1..%d", p, PG_MAX_ORDER);
    PG_REQUIRE(table, PG_EINVAL, "pg_tables_init: null table");
    cudaStream_t st = (cudaStream_t)stream;
    const int nexp = pg_nexp(p);
    // conical product rule exact for degree 2p+1 (the integrands have degree <= 2p): p+1 points per direction
    const int n1 = p + 1, nq = n1 * n1 * n1;
    std::vector<double> u(n1), wu(n1), v(n1), wv(n1), w(n1), ww(n1), pts(3 * (size_t)nq), wts(nq);
    gauss_jacobi_01(n1, 2, u.data(), wu.data());
    gauss_jacobi_01(n1, 1, v.data(), wv.data());
    gauss_jacobi_01(n1, 0, w.data(), ww.data());
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n1; ++j)
            for (int k = 0; k < n1; ++k) {
                const int g = (i * n1 + j) * n1 + k;
                pts[3 * g] = u[i];
                pts[3 * g + 1] = v[j] * (1.0 - u[i]);
                pts[3 * g + 2] = w[k] * (1.0 - u[i]) * (1.0 - v[j]);
                wts[g] = wu[i] * wv[j] * ww[k];
            }
    double *d_pts = nullptr, *d_wts = nullptr;
    fe::V3 *d_N = nullptr, *d_C = nullptr;
    const size_t fbytes = (size_t)nq * nexp * sizeof(fe::V3);
    PG_CUDA_OK(cudaMalloc((void **)&d_pts, pts.size() * 8));
    PG_CUDA_OK(cudaMalloc((void **)&d_wts, wts.size() * 8));
    PG_CUDA_OK(cudaMalloc((void **)&d_N, fbytes));
    PG_CUDA_OK(cudaMalloc((void **)&d_C, fbytes));
    PG_CUDA_OK(cudaMemcpyAsync(d_pts, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice, st));
    PG_CUDA_OK(cudaMemcpyAsync(d_wts, wts.data(), wts.size() * 8, cudaMemcpyHostToDevice, st));
    eval_points_kernel<<<(nq + 31) / 32, 32, 0, st>>>(p, nexp, nq, d_pts, d_N, d_C);
    const int64_t pairs = (int64_t)nexp * nexp;
    contract_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(nexp, nq, d_wts, d_N, d_C, table);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the host vectors and scratch go away below
    cudaFree(d_pts), cudaFree(d_wts), cudaFree(d_N), cudaFree(d_C);
    PG_REQUIRE(e == cudaSuccess, PG_ECUDA, "pg_tables_init: %s", cudaGetErrorString(e));
    return PG_OK;
}

int pg_locate_points(int64_t T, const double *nodes, int64_t npts, const double *points, double tol,
                     int32_t *pt_elem, void *stream) {
    PG_REQUIRE(T >= 0 && npts >= 0 && (npts == 0 || (points && pt_elem)) && (T == 0 || nodes), PG_EINVAL,
               "pg_locate_points: bad argument");
    if (npts == 0) return PG_OK;
    PG_REQUIRE(npts <= 65535, PG_ERANGE, "pg_locate_points: at most 65535 points per call (%lld)", (long long)npts);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *best = nullptr;
    PG_CUDA_OK(cudaMalloc((void **)&best, npts * sizeof(unsigned long long)));
    PG_CUDA_OK(cudaMemsetAsync(best, 0xff, npts * sizeof(unsigned long long), st));
    if (T > 0) {
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((T + 255) / 256, 4 * (int64_t)kNumSMs));
        locate_kernel<<<dim3(gx, (unsigned)npts), 256, 0, st>>>(T, nodes, npts, points, tol, best);
    }
    finish_locate_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(npts, best, pt_elem);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(best);
    PG_REQUIRE(e == cudaSuccess, PG_ECUDA, "pg_locate_points: %s", cudaGetErrorString(e));
    return PG_OK;
}

int pg_interpolate_fields(int64_t npts, const double *points, const int32_t *pt_elem, int p, const double *nodes,
                          const uint32_t *code, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                          int64_t nFaces, const int32_t *perm, const double *x, double omega, double mu, double *fields,
                          void *stream) {
    PG_REQUIRE(p >= 1 && p <= PG_MAX_ORDER, PG_EINVAL, "pg_interpolate_fields: order %d", p);
    PG_REQUIRE(npts >= 0 && (npts == 0 || (points && pt_elem && nodes && code && elemsE && x && fields)), PG_EINVAL,
               "pg_interpolate_fields: null pointer");
    PG_REQUIRE(p < 2 || elemsF, PG_EINVAL, "pg_interpolate_fields: elemsF needed for p >= 2");
    if (npts == 0) return PG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    fe::V3 *scratch = nullptr;
    PG_CUDA_OK(cudaMalloc((void **)&scratch, (size_t)npts * 2 * pg_ndof_element(p) * sizeof(fe::V3)));
    interpolate_kernel<<<(unsigned)((npts + 31) / 32), 32, 0, st>>>(
        npts, points, pt_elem, p, nodes, code, elemsE, elemsF, nEdges, nFaces, perm,
        reinterpret_cast<const double2 *>(x), omega * mu, reinterpret_cast<double2 *>(fields), scratch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    PG_REQUIRE(e == cudaSuccess, PG_ECUDA, "pg_interpolate_fields: %s", cudaGetErrorString(e));
    return PG_OK;
}

int pg_csem_rhs(int p, int64_t source_elem, const double *position_host, const double *moment_host, const double *nodes,
                const uint32_t *code, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges, int64_t nFaces,
                const int32_t *perm, int64_t row_begin, int64_t local_rows, double omega, double mu, double *b,
                void *stream) {
    PG_REQUIRE(p >= 1 && p <= PG_MAX_ORDER, PG_EINVAL, "pg_csem_rhs: order %d", p);
    PG_REQUIRE(source_elem >= 0 && position_host && moment_host && nodes && code && elemsE && b, PG_EINVAL,
               "pg_csem_rhs: bad argument");
    PG_REQUIRE(p < 2 || elemsF, PG_EINVAL, "pg_csem_rhs: elemsF needed for p >= 2");
    cudaStream_t st = (cudaStream_t)stream;
    double *dbuf = nullptr;
    fe::V3 *scratch = nullptr;
    PG_CUDA_OK(cudaMalloc((void **)&dbuf, 6 * sizeof(double)));
    PG_CUDA_OK(cudaMalloc((void **)&scratch, 2 * (size_t)pg_ndof_element(p) * sizeof(fe::V3)));
    double h[6] = {position_host[0], position_host[1], position_host[2], moment_host[0], moment_host[1], moment_host[2]};
    PG_CUDA_OK(cudaMemcpyAsync(dbuf, h, sizeof(h), cudaMemcpyHostToDevice, st));
    csem_rhs_kernel<<<1, 32, 0, st>>>(p, source_elem, dbuf, dbuf + 3, nodes, code, elemsE, elemsF, nEdges, nFaces, perm,
                                      row_begin, local_rows, omega * mu, reinterpret_cast<double2 *>(b), scratch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dbuf), cudaFree(scratch);
    PG_REQUIRE(e == cudaSuccess, PG_ECUDA, "pg_csem_rhs: %s", cudaGetErrorString(e));
    return PG_OK;
}

}  // extern "C"

// The Phi-GEMM form of computeElementalMatrices on the FP64 tensor pipe (DMMA), for the A/B against the
// table contraction that the product uses.
//
// Reference: computeElementalMatrices (hvfem.py:223-316) evaluates, per element and Gauss point, the mapped
// basis J^-1 N_g and J^T C_g / detJ and accumulates Me += detJ W_g (J^-1 N_g)^T diag(sigma) (J^-1 N_g),
// Ke += detJ W_g (J^T C_g/detJ)^T (J^T C_g/detJ).  Written as matrix products that is, per element,
//     Me = sign(detJ) U^T U,  U[(g,i), j] = sqrt(W_g |detJ| sigma_i) s_j (J^-1 N^_{J(j)}(g))_i
//     Ke = sign(detJ) V^T V,  V[(g,m), j] = sqrt(W_g / |detJ|)      s_j (J^T  C^_{J(j)}(g))_m
// two n x n x 3*ngauss SYRKs -- the only GEMM-shaped form of this path (4 n^2 3 ngauss flops per
// element against 24 n^2 for the table contraction with K = 12).  tcgen05 has no FP64 kind; the FP64 tensor
// instruction of sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA), used below.  N^, C^ are the orientation-expanded
// reference functions at the Gauss points (host: basis.evaluate_expanded), J(j), s_j the expanded index and
// sign of local dof j under the element's orientation code.
//
// This path is NOT used by pg_assemble: tools/dmma_ab.py measures it next to pg_element_matrices (same
// output) and profiles/r2_dmma_ab.json records why (flop count, DMMA rate = DFMA rate on B200).
#include <math.h>

#include <algorithm>

#include "pg_common.cuh"

namespace pg {
namespace {

template <int P>
__global__ void __launch_bounds__(256) phi_operands_kernel(int64_t T, const double *__restrict__ nodes,
                                                           const double *__restrict__ sigma,
                                                           const uint32_t *__restrict__ code, int ng,
                                                           const double *__restrict__ phiN,
                                                           const double *__restrict__ phiC,
                                                           const double *__restrict__ wts, int64_t ld, int64_t kpad,
                                                           double *__restrict__ U, double *__restrict__ V,
                                                           double *__restrict__ sgn) {
    using O = Ord<P>;
    const int64_t t = blockIdx.x;
    __shared__ double sJ[9], sA[9], sdet;
    if (threadIdx.x == 0) {
        double x[12];
        for (int i = 0; i < 12; ++i) x[i] = nodes[t * 12 + i];
        double J[3][3];
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) J[i][c] = x[3 * (i + 1) + c] - x[c];
        double A[3][3];  // adjugate: J^-1 = A / det (same expressions as geometry_kernel)
        A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
        A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
        A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
        A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
        A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) {
                sJ[3 * i + c] = J[i][c];
                sA[3 * i + c] = A[i][c] / det;
            }
        sdet = det;
        sgn[t] = det < 0.0 ? -1.0 : 1.0;
    }
    __syncthreads();
    const uint32_t cd = code[t];
    const double adet = fabs(sdet);
    const double sg[3] = {sigma[2 * t], sigma[2 * t], sigma[2 * t + 1]};
    double *Ut = U + t * kpad * ld, *Vt = V + t * kpad * ld;
    const int total = ng * O::n;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int gp = idx / O::n, j = idx - gp * O::n;
        int slot, d;
        slot_of_local<P>(j, slot, d);
        double s;
        const int Jx = expanded_of_slot<P>(slot, d, cd, s);
        const double *pn = phiN + ((int64_t)Jx * ng + gp) * 3, *pc = phiC + ((int64_t)Jx * ng + gp) * 3;
        const double n0 = pn[0], n1 = pn[1], n2 = pn[2], c0 = pc[0], c1 = pc[1], c2 = pc[2];
        const double w = wts[gp];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double u = sA[3 * i] * n0 + sA[3 * i + 1] * n1 + sA[3 * i + 2] * n2;  // (J^-1 N^)_i
            const double v = sJ[i] * c0 + sJ[3 + i] * c1 + sJ[6 + i] * c2;              // (J^T C^)_i
            Ut[(int64_t)(3 * gp + i) * ld + j] = sqrt(w * adet * sg[i]) * s * u;
            Vt[(int64_t)(3 * gp + i) * ld + j] = sqrt(w / adet) * s * v;
        }
    }
    // zero padding: columns n..ld-1 of every row, rows 3 ng..kpad-1
    for (int64_t idx = threadIdx.x; idx < kpad * ld; idx += blockDim.x) {
        const int64_t r = idx / ld, cidx = idx - r * ld;
        if (r >= 3 * ng || cidx >= O::n) {
            Ut[idx] = 0.0;
            Vt[idx] = 0.0;
        }
    }
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[t] = sgn[t] X[t]^T X[t]: one warp per 16 x 16 block of the upper triangle (2 x 2 DMMA tiles, 4 loads per
// 4 DMMAs, operands straight from L1/L2), mirrored on store.
__global__ void __launch_bounds__(128) syrk_dmma_kernel(int n, int64_t ld, int64_t kpad, int nb16,
                                                        const double *__restrict__ X,
                                                        const double *__restrict__ sgn, double *__restrict__ C) {
    const int64_t t = blockIdx.y;
    const int lane = threadIdx.x & 31, gid = lane >> 2, tg = lane & 3;
    int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nb16 * (nb16 + 1) / 2) return;
    int bi = 0;
    while (b >= nb16 - bi) {  // row bi of the upper triangle has nb16 - bi blocks
        b -= nb16 - bi;
        ++bi;
    }
    const int bj = bi + b;
    const int i0 = bi * 16, j0 = bj * 16;
    const double *Xt = X + t * kpad * ld;
    double c[2][2][2] = {};
    for (int64_t k = 0; k < kpad; k += 4) {
        const double *row = Xt + (k + tg) * ld;
        const double a0 = __ldg(row + i0 + gid), a1 = __ldg(row + i0 + 8 + gid);
        const double b0 = __ldg(row + j0 + gid), b1 = __ldg(row + j0 + 8 + gid);
        dmma884(c[0][0][0], c[0][0][1], a0, b0);
        dmma884(c[0][1][0], c[0][1][1], a0, b1);
        dmma884(c[1][0][0], c[1][0][1], a1, b0);
        dmma884(c[1][1][0], c[1][1][1], a1, b1);
    }
    const double s = sgn[t];
    double *Ct = C + t * (int64_t)n * n;
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = i0 + 8 * ti + gid, j = j0 + 8 * tj + 2 * tg + e;
                if (i < n && j < n) {
                    const double v = s * c[ti][tj][e];
                    Ct[(int64_t)i * n + j] = v;
                    if (bi != bj) Ct[(int64_t)j * n + i] = v;
                }
            }
}

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int64_t pg_phi_gemm_workspace_doubles(int64_t T, int p, int ngauss) {
    const int64_t ld = (ndof_element(p) + 15) / 16 * 16, kpad = (3 * (int64_t)ngauss + 3) / 4 * 4;
    return 2 * T * kpad * ld + T;
}

int pg_element_matrices_phi_gemm(int64_t T, int p, const double *nodes, const double *sigma, const uint32_t *code,
                                 int ngauss, const double *phiN, const double *phiC, const double *weights,
                                 double *work, int stage, double *Me, double *Ke, void *stream) {
    PG_REQUIRE(T >= 0 && nodes && sigma && code && phiN && phiC && weights && work && ngauss > 0, PG_EINVAL,
               "pg_element_matrices_phi_gemm: bad argument");
    if (T == 0) return PG_OK;
    const int n = ndof_element(p);
    const int64_t ld = (n + 15) / 16 * 16, kpad = (3 * (int64_t)ngauss + 3) / 4 * 4;
    double *U = work, *V = work + T * kpad * ld, *sgn = work + 2 * T * kpad * ld;
    cudaStream_t st = (cudaStream_t)stream;
    if (stage == 0 || stage == 1) {
        const int rc = dispatch_order(p, [&](auto o) {
            constexpr int P = decltype(o)::p;
            phi_operands_kernel<P><<<(unsigned)T, 256, 0, st>>>(T, nodes, sigma, code, ngauss, phiN, phiC, weights, ld,
                                                                kpad, U, V, sgn);
            PG_LAUNCH_OK();
            return PG_OK;
        });
        if (rc != PG_OK) return rc;
    }
    if (stage == 0 || stage == 2) {
        PG_REQUIRE(Me && Ke, PG_EINVAL, "pg_element_matrices_phi_gemm: null output");
        const int nb16 = (int)(ld / 16);
        const int nblk = nb16 * (nb16 + 1) / 2;
        dim3 grid((nblk + 3) / 4, (unsigned)T);
        syrk_dmma_kernel<<<grid, 128, 0, st>>>(n, ld, kpad, nb16, U, sgn, Me);
        PG_LAUNCH_OK();
        syrk_dmma_kernel<<<grid, 128, 0, st>>>(n, ld, kpad, nb16, V, sgn, Ke);
        PG_LAUNCH_OK();
    }
    return PG_OK;
}

}  // extern "C"

// pg_krylov_solve: the complex symmetric Krylov solvers (COCG, COCR) with Jacobi preconditioning as one
// C-ABI call on one GPU, for callers that bind the library without the Python drivers.
// Reference: ksp.solve(b, x) at solver.py:584-590 with -ksp_type cg -ksp_cg_type symmetric / cr and
// -pc_type jacobi.  Same recurrences, kernels and convergence test (preconditioned residual norm
// relative to ||M^-1 b||) as petgem_b200/krylov.py:cocg_multi with one right-hand side; the host reads
// the residual norm every `check_every` iterations only.
#include <math.h>

#include <algorithm>

#include "pg_common.cuh"

namespace pg {
namespace {

__global__ void __launch_bounds__(256) inv_diag_kernel(int64_t n, double2 *__restrict__ d) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = d[i];
        const double den = v.x * v.x + v.y * v.y;
        d[i] = den == 0.0 ? make_double2(1.0, 0.0) : make_double2(v.x / den, -v.y / den);  // PCJACOBI: 0 -> 1
    }
}

constexpr int kScalars = 16;  // complex device scalars kept in the workspace

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int64_t pg_krylov_workspace_bytes(int64_t n) {
    return (6 * n + kScalars) * 16 + pg_reduce_workspace_bytes(2);
}

int pg_krylov_solve(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *b,
                    double *x, int method, int jacobi, double rtol, int maxit, int check_every, void *work,
                    int *iterations, double *rel_residual, void *stream) {
    PG_REQUIRE(n >= 0 && rowptr && b && x && work && iterations && rel_residual, PG_EINVAL,
               "pg_krylov_solve: bad argument");
    PG_REQUIRE(method == 0 || method == 1, PG_EINVAL, "pg_krylov_solve: method %d (0 = COCG, 1 = COCR)", method);
    PG_REQUIRE(maxit >= 0 && check_every >= 1 && rtol >= 0.0, PG_EINVAL, "pg_krylov_solve: bad control parameter");
    cudaStream_t st = (cudaStream_t)stream;
    *iterations = 0;
    *rel_residual = 0.0;
    if (n == 0) return PG_OK;
    const size_t vb = (size_t)n * 16;
    double *w = static_cast<double *>(work);
    double *dinv = w, *Z = w + 2 * n, *P = w + 4 * n, *Q = w + 6 * n, *AR = w + 8 * n, *R = w + 10 * n;
    double *sc = w + 12 * n;  // [0,1] rho ping-pong  [2] pq  [3,4] alpha,-alpha  [5,6] beta,-beta  [7,8] r^T z, |z|^2
    void *red = sc + 2 * kScalars;
    const double *dp = jacobi ? dinv : nullptr;
#define PG_TRY(call)                 \
    do {                             \
        const int rc_ = (call);      \
        if (rc_ != PG_OK) return rc_; \
    } while (0)

    PG_CUDA_OK(cudaMemsetAsync(x, 0, vb, st));
    PG_CUDA_OK(cudaMemsetAsync(sc, 0, kScalars * 16, st));
    if (jacobi) {
        PG_TRY(pg_csr_diagonal(n, 0, rowptr, colidx, vals, dinv, st));
        inv_diag_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8), 256, 0, st>>>(
            n, reinterpret_cast<double2 *>(dinv));
        PG_LAUNCH_OK();
        PG_TRY(pg_zbscale_rows(n, 1, dinv, b, Z, st));
    } else {
        PG_CUDA_OK(cudaMemcpyAsync(Z, b, vb, cudaMemcpyDeviceToDevice, st));
    }
    double host2[4];
    PG_TRY(pg_zbnrm2sq(n, 1, Z, sc + 16, red, st));
    PG_CUDA_OK(cudaMemcpyAsync(host2, sc + 16, 16, cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    const double bnorm = sqrt(host2[0]);
    if (bnorm == 0.0) return PG_OK;  // zero right-hand side: x = 0
    const double tol = rtol * bnorm;
    PG_CUDA_OK(cudaMemcpyAsync(P, Z, vb, cudaMemcpyDeviceToDevice, st));
    if (method == 0) {
        PG_CUDA_OK(cudaMemcpyAsync(R, b, vb, cudaMemcpyDeviceToDevice, st));
        PG_TRY(pg_zbdotu(n, 1, R, Z, sc, red, st));
    } else {
        PG_TRY(pg_spmv_scaled(n, rowptr, colidx, vals, Z, nullptr, AR, st));
        PG_CUDA_OK(cudaMemcpyAsync(Q, AR, vb, cudaMemcpyDeviceToDevice, st));
        PG_TRY(pg_zbdotu(n, 1, Z, AR, sc, red, st));
    }
    int cur = 0, it = 0;
    double res = bnorm;
    while (it < maxit) {
        const int count = std::min(check_every, maxit - it);
        for (int c = 0; c < count; ++c) {
            double *rho = sc + 2 * cur, *rho_new = sc + 2 * (cur ^ 1);
            if (method == 0) {
                PG_TRY(pg_spmv_scaled(n, rowptr, colidx, vals, P, nullptr, Q, st));
                PG_TRY(pg_zbdotu(n, 1, P, Q, sc + 4, red, st));
                PG_TRY(pg_zbdiv(1, rho, sc + 4, sc + 6, st));
                PG_TRY(pg_cocg_step(n, 1, sc + 6, P, Q, dp, x, R, Z, sc + 14, red, st));
                PG_CUDA_OK(cudaMemcpyAsync(rho_new, sc + 14, 16, cudaMemcpyDeviceToDevice, st));
                PG_TRY(pg_zbdiv(1, rho_new, rho, sc + 10, st));
                PG_TRY(pg_zbaypx(n, 1, sc + 10, Z, P, st));
            } else {
                PG_TRY(pg_zbdotu_w(n, 1, Q, Q, dp, sc + 4, red, st));
                PG_TRY(pg_zbdiv(1, rho, sc + 4, sc + 6, st));
                PG_TRY(pg_cocr_update(n, 1, sc + 6, P, Q, dp, x, Z, st));
                PG_TRY(pg_spmv_scaled(n, rowptr, colidx, vals, Z, nullptr, AR, st));
                PG_TRY(pg_zbdotu(n, 1, Z, AR, rho_new, red, st));
                PG_TRY(pg_zbdiv(1, rho_new, rho, sc + 10, st));
                PG_TRY(pg_cocr_direction(n, 1, sc + 10, Z, AR, P, Q, st));
            }
            cur ^= 1;
        }
        it += count;
        if (method == 1) PG_TRY(pg_zbnrm2sq(n, 1, Z, sc + 16, red, st));
        PG_CUDA_OK(cudaMemcpyAsync(host2, sc + 16, 16, cudaMemcpyDeviceToHost, st));  // |z|^2 of this batch
        PG_CUDA_OK(cudaStreamSynchronize(st));
        if (!(host2[0] == host2[0]) || host2[0] < 0.0) {  // NaN: breakdown
            *iterations = it;
            *rel_residual = host2[0];
            set_error("pg_krylov_solve: breakdown after %d iterations", it);
            return PG_ERANGE;
        }
        res = sqrt(host2[0]);
        if (res <= tol) break;
    }
#undef PG_TRY
    *iterations = it;
    *rel_residual = res / bnorm;
    return PG_OK;
}

}  // extern "C"

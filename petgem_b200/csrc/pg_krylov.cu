// pg_krylov_solve: the Krylov solvers as one C-ABI call on one GPU, for callers that bind the library
// without the Python drivers: COCG / COCR for the complex symmetric system, and the two general solvers the
// reference's option files name, BiCGStab (-ksp_type bcgs) and restarted GMRES (-ksp_type gmres, the PETSc
// default: left preconditioning, classical Gram-Schmidt), each with optional Jacobi preconditioning.
// Reference: ksp.solve(b, x) at solver.py:584-590.  Same recurrences, kernels and convergence test
// (preconditioned residual norm relative to ||M^-1 b||) as petgem_b200/krylov.py.  COCG/COCR keep their
// coefficients on the device and let the host read the residual every `check_every` iterations only;
// BiCGStab and GMRES read their scalars every iteration, like KSPBCGS / KSPGMRES.
#include <math.h>

#include <algorithm>
#include <complex>
#include <vector>

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

int64_t pg_krylov_workspace_bytes(int64_t n, int method, int restart) {
    int64_t vectors = 6;                                            // COCG / COCR
    if (method == PG_KSP_BCGS) vectors = 7;                         // dinv r rhat p v s t
    if (method == PG_KSP_GMRES) vectors = (int64_t)std::max(restart, 1) + 3;  // dinv w V[restart + 1]
    const int nsc = method == PG_KSP_GMRES ? 2 * (std::max(restart, 1) + 4) : kScalars;
    return (vectors * n + nsc) * 16 + pg_reduce_workspace_bytes(method == PG_KSP_GMRES ? std::max(restart, 1) + 2 : 2);
}

static int solve_general(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *b,
                         double *x, int method, int restart, int jacobi, double rtol, int maxit, void *work,
                         int *iterations, double *rel_residual, cudaStream_t st);

int pg_krylov_solve(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *b,
                    double *x, int method, int restart, int jacobi, double rtol, int maxit, int check_every, void *work,
                    int *iterations, double *rel_residual, void *stream) {
    PG_REQUIRE(n >= 0 && rowptr && b && x && work && iterations && rel_residual, PG_EINVAL,
               "pg_krylov_solve: bad argument");
    PG_REQUIRE(method >= PG_KSP_COCG && method <= PG_KSP_GMRES, PG_EINVAL,
               "pg_krylov_solve: method %d (0 = COCG, 1 = COCR, 2 = BiCGStab, 3 = GMRES)", method);
    PG_REQUIRE(maxit >= 0 && check_every >= 1 && rtol >= 0.0, PG_EINVAL, "pg_krylov_solve: bad control parameter");
    PG_REQUIRE(method != PG_KSP_GMRES || (restart >= 1 && restart <= 1024), PG_EINVAL,
               "pg_krylov_solve: GMRES restart %d (1..1024)", restart);
    cudaStream_t st = (cudaStream_t)stream;
    *iterations = 0;
    *rel_residual = 0.0;
    if (n == 0) return PG_OK;
    if (method == PG_KSP_BCGS || method == PG_KSP_GMRES)
        return solve_general(n, rowptr, colidx, vals, b, x, method, restart, jacobi, rtol, maxit, work, iterations,
                             rel_residual, st);
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
    *iterations = it;
    *rel_residual = res / bnorm;
    return PG_OK;
}

// BiCGStab and GMRES(restart): the scalars of every iteration visit the host (KSPBCGS / KSPGMRES do the same)
static int solve_general(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *b,
                         double *x, int method, int restart, int jacobi, double rtol, int maxit, void *work,
                         int *iterations, double *rel_residual, cudaStream_t st) {
    typedef std::complex<double> cplx;
    const size_t vb = (size_t)n * 16;
    double *w0 = static_cast<double *>(work);
    double *dinv = w0;
    const double *dp = jacobi ? dinv : nullptr;
    auto vec = [&](int i) { return w0 + 2 * n * (int64_t)(i + 1); };
    const int nvec = method == PG_KSP_BCGS ? 6 : restart + 2;
    double *sc = vec(nvec);  // device scalars
    const int nsc = method == PG_KSP_GMRES ? 2 * (restart + 4) : kScalars;
    void *red = sc + 2 * nsc;
    auto put = [&](int i, cplx v) -> int {  // host scalar -> device scalar i
        double h[2] = {v.real(), v.imag()};
        PG_CUDA_OK(cudaMemcpyAsync(sc + 2 * i, h, 16, cudaMemcpyHostToDevice, st));
        return PG_OK;
    };
    auto get = [&](int i, int count, cplx *out) -> int {
        PG_CUDA_OK(cudaMemcpyAsync(out, sc + 2 * i, 16 * (size_t)count, cudaMemcpyDeviceToHost, st));
        PG_CUDA_OK(cudaStreamSynchronize(st));
        return PG_OK;
    };
    auto precond = [&](const double *in, double *out) -> int {  // out = M^-1 in
        if (jacobi) return pg_zbscale_rows(n, 1, dinv, in, out, st);
        PG_CUDA_OK(cudaMemcpyAsync(out, in, vb, cudaMemcpyDeviceToDevice, st));
        return PG_OK;
    };
    auto apply = [&](const double *in, double *out) -> int {  // out = M^-1 A in (Jacobi rides in the epilogue)
        return pg_spmv_scaled(n, rowptr, colidx, vals, in, dp, out, st);
    };
    PG_CUDA_OK(cudaMemsetAsync(x, 0, vb, st));
    PG_CUDA_OK(cudaMemsetAsync(sc, 0, (size_t)nsc * 16, st));
    if (jacobi) {
        PG_TRY(pg_csr_diagonal(n, 0, rowptr, colidx, vals, dinv, st));
        inv_diag_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8), 256, 0, st>>>(
            n, reinterpret_cast<double2 *>(dinv));
        PG_LAUNCH_OK();
    }
    cplx h1;
    const int ONE = 0, S1 = 1, S2 = 2, S3 = 3, OUT = 4;  // scalar slots
    PG_TRY(put(ONE, cplx(1.0, 0.0)));

    if (method == PG_KSP_BCGS) {
        double *r = vec(0), *rhat = vec(1), *p = vec(2), *v = vec(3), *s = vec(4), *t = vec(5);
        PG_TRY(precond(b, r));
        PG_TRY(pg_dznrm2sq(n, r, sc + 2 * OUT, red, st));
        PG_TRY(get(OUT, 1, &h1));
        const double bnorm = sqrt(h1.real());
        if (bnorm == 0.0) return PG_OK;
        const double tol = rtol * bnorm;
        PG_CUDA_OK(cudaMemcpyAsync(rhat, r, vb, cudaMemcpyDeviceToDevice, st));
        PG_CUDA_OK(cudaMemsetAsync(p, 0, vb, st));
        PG_CUDA_OK(cudaMemsetAsync(v, 0, vb, st));
        cplx rho(1.0), alpha(1.0), omega(1.0);
        double res = bnorm;
        int it = 0;
        while (it < maxit) {
            cplx rho_new;
            PG_TRY(pg_zdotc(n, rhat, r, sc + 2 * OUT, red, st));
            PG_TRY(get(OUT, 1, &rho_new));
            if (rho_new == cplx(0.0) || !(rho_new == rho_new)) break;  // breakdown
            const cplx beta = (rho_new / rho) * (alpha / omega);
            PG_TRY(put(S1, beta));
            PG_TRY(put(S2, -beta * omega));
            PG_TRY(pg_zaxpbypcz(n, sc + 2 * ONE, r, sc + 2 * S1, p, sc + 2 * S2, v, p, st));  // p = r + beta (p - omega v)
            PG_TRY(apply(p, v));
            cplx den;
            PG_TRY(pg_zdotc(n, rhat, v, sc + 2 * OUT, red, st));
            PG_TRY(get(OUT, 1, &den));
            if (den == cplx(0.0)) break;
            alpha = rho_new / den;
            PG_TRY(put(S1, -alpha));
            PG_TRY(pg_zaxpbypcz(n, sc + 2 * ONE, r, sc + 2 * S1, v, nullptr, nullptr, s, st));  // s = r - alpha v
            PG_TRY(apply(s, t));
            cplx tst[2];
            PG_TRY(pg_zdotc(n, t, s, sc + 2 * OUT, red, st));
            PG_TRY(pg_dznrm2sq(n, t, sc + 2 * (OUT + 1), red, st));
            PG_TRY(get(OUT, 2, tst));
            if (tst[1].real() == 0.0) break;
            omega = tst[0] / tst[1].real();
            PG_TRY(put(S1, alpha));
            PG_TRY(put(S2, omega));
            PG_TRY(put(S3, -omega));
            PG_TRY(pg_zaxpy(n, sc + 2 * S1, p, x, st));
            PG_TRY(pg_zaxpy(n, sc + 2 * S2, s, x, st));
            PG_TRY(pg_zaxpbypcz(n, sc + 2 * ONE, s, sc + 2 * S3, t, nullptr, nullptr, r, st));  // r = s - omega t
            rho = rho_new;
            PG_TRY(pg_dznrm2sq(n, r, sc + 2 * OUT, red, st));
            PG_TRY(get(OUT, 1, &h1));
            res = sqrt(h1.real());
            ++it;
            if (!(res == res)) {
                *iterations = it;
                set_error("pg_krylov_solve: BiCGStab breakdown after %d iterations", it);
                return PG_ERANGE;
            }
            if (res <= tol) break;
        }
        *iterations = it;
        *rel_residual = res / bnorm;
        return PG_OK;
    }

    // GMRES(restart), left preconditioning, classical Gram-Schmidt: one VecMDot + one fused VecMAXPY/norm +
    // one scaled copy per iteration; the Hessenberg column (<= restart + 1 scalars) is the host sync
    double *w = vec(0), *tmp = w;  // w doubles as the scratch of the residual evaluation
    double *V = vec(1);            // V[restart + 1][n]
    const int HC = OUT;            // hcol: restart + 2 scalars; ycoef behind it
    const int YC = HC + restart + 2;
    PG_TRY(precond(b, V));
    PG_TRY(pg_dznrm2sq(n, V, sc + 2 * S1, red, st));
    PG_TRY(get(S1, 1, &h1));
    const double bnorm = sqrt(h1.real());
    if (bnorm == 0.0) return PG_OK;
    const double tol = rtol * bnorm;
    std::vector<cplx> H((size_t)(restart + 1) * restart), g(restart + 1), sn(restart), hcol(restart + 2), y(restart);
    std::vector<double> cs(restart);
    int it = 0;
    double res = bnorm;
    bool x_is_zero = true;
    PG_TRY(put(S2, cplx(-1.0, 0.0)));
    while (true) {
        if (!x_is_zero) {  // V0 = M^-1 (b - A x)
            PG_TRY(pg_spmv_scaled(n, rowptr, colidx, vals, x, nullptr, tmp, st));
            PG_TRY(pg_zaxpbypcz(n, sc + 2 * ONE, b, sc + 2 * S2, tmp, nullptr, nullptr, tmp, st));
            PG_TRY(precond(tmp, V));
        }
        PG_TRY(pg_dznrm2sq(n, V, sc + 2 * S1, red, st));
        PG_TRY(get(S1, 1, &h1));
        const double beta = sqrt(h1.real());
        res = beta;
        if (beta <= tol || it >= maxit) break;
        PG_TRY(put(S1, cplx(beta, 0.0)));
        PG_TRY(pg_zscal(n, sc + 2 * S1, 1, V, st));
        std::fill(g.begin(), g.end(), cplx(0.0));
        g[0] = beta;
        int k = 0;
        while (k < restart && it < maxit) {
            double *vk = V + 2 * n * (int64_t)k, *vk1 = V + 2 * n * (int64_t)(k + 1);
            PG_TRY(apply(vk, w));
            PG_TRY(pg_zmdotc(n, k + 1, V, n, w, sc + 2 * HC, red, st));
            PG_TRY(pg_zmaxpy_nrm2sq(n, k + 1, sc + 2 * HC, -1.0, V, n, w, sc + 2 * (HC + k + 1), red, st));
            PG_TRY(get(HC, k + 2, hcol.data()));
            const double wn = sqrt(hcol[k + 1].real());
            hcol[k + 1] = wn;
            PG_TRY(put(S3, cplx(wn, 0.0)));
            PG_TRY(pg_zcopy_scaled(n, sc + 2 * S3, 1, w, vk1, st));  // v_{k+1} = w / ||w||
            for (int i = 0; i < k; ++i) {  // previous Givens rotations
                const cplx t = cs[i] * hcol[i] + sn[i] * hcol[i + 1];
                hcol[i + 1] = -std::conj(sn[i]) * hcol[i] + cs[i] * hcol[i + 1];
                hcol[i] = t;
            }
            const cplx a_ = hcol[k], b_ = hcol[k + 1];
            const double den = sqrt(std::norm(a_) + std::norm(b_));
            if (std::abs(a_) == 0.0) {
                cs[k] = 0.0;
                sn[k] = 1.0;
            } else {
                cs[k] = std::abs(a_) / den;
                sn[k] = (a_ / std::abs(a_)) * std::conj(b_) / den;
            }
            hcol[k] = cs[k] * a_ + sn[k] * b_;
            hcol[k + 1] = 0.0;
            for (int i = 0; i <= k + 1 && i <= restart; ++i) H[(size_t)i * restart + k] = hcol[i];
            g[k + 1] = -std::conj(sn[k]) * g[k];
            g[k] = cs[k] * g[k];
            ++k;
            ++it;
            res = std::abs(g[k]);
            if (res <= tol || b_ == cplx(0.0)) break;
        }
        for (int i = k - 1; i >= 0; --i) {  // back substitution on the triangular factor
            cplx acc = g[i];
            for (int j = i + 1; j < k; ++j) acc -= H[(size_t)i * restart + j] * y[j];
            y[i] = acc / H[(size_t)i * restart + i];
        }
        if (k > 0) {
            PG_CUDA_OK(cudaMemcpyAsync(sc + 2 * YC, y.data(), 16 * (size_t)k, cudaMemcpyHostToDevice, st));
            PG_TRY(pg_zmaxpy(n, k, sc + 2 * YC, 1.0, V, n, x, st));  // x += V y
            PG_CUDA_OK(cudaStreamSynchronize(st));                    // y is reused by the next cycle
            x_is_zero = false;
        }
        if (res <= tol) break;
    }
    *iterations = it;
    *rel_residual = res / bnorm;
    return PG_OK;
#undef PG_TRY
}

}  // extern "C"

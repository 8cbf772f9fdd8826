// Kernels of the gradient-space (Hiptmair) preconditioner.
//
// Reference intent: the shipped option files ask PETSc for -pc_type sor / asm / gamg
// (examples/case1..5 petsc.opts, consumed by KSP.setFromOptions at solver.py:586-589).  None of
// those has a parallel, symmetric GPU form, and none addresses why the curl-curl system converges
// slowly at 2 Hz: the gradient fields grad(phi) lie in the null space of the curl-curl part, so
// A grad(phi) = -i w mu M_sigma grad(phi) is ~1e-3 of the diagonal.  The hybrid smoother of
// Hiptmair (1998) repairs exactly that and stays symmetric, so COCG/COCR remain valid:
//
//     M^-1 = D^-1 + G diag(G^T A G)^-1 G^T,
//
// G = discrete gradient of the hierarchical H1 space (vertex functions and, for p >= 2, the
// quadratic edge functions) expressed in the Nedelec dofs: 1-3 REAL entries per row.  Two small
// gather products per application (about 12 % of the bytes of one SpMV with A).
//
//   pg_rcsr_apply         Y = s .* (R X) + a .* Z, R a real CSR matrix (G or G^T), X complex, K rhs
//   pg_galerkin_diagonal  d_k = sum_{i owned} sum_j R[k,i] A[i,j] R[k,j]   (R = G^T rows)
#include <algorithm>

#include "pg_common.cuh"

namespace pg {
namespace {

__device__ __forceinline__ double2 acmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// one thread per (row, right-hand side): rows are short (<= 3 entries for G, ~14 for G^T)
template <int K>
__global__ void __launch_bounds__(256) rcsr_apply_kernel(int64_t rows, const int32_t *__restrict__ rowptr,
                                                         const int32_t *__restrict__ colidx,
                                                         const double *__restrict__ vals,
                                                         const double2 *__restrict__ X,
                                                         const double2 *__restrict__ s,
                                                         const double2 *__restrict__ a,
                                                         const double2 *__restrict__ Z, double2 *__restrict__ Y) {
    const int64_t total = rows * K;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / K;
        const int r = (int)(idx - i * K);
        const int b = __ldg(rowptr + i), e = __ldg(rowptr + i + 1);
        double2 acc = make_double2(0.0, 0.0);
        int j = b;
        // rows of G^T hold ~11-14 entries: four independent (index, value, gather) chains in flight per thread
        // (one chain per entry left the kernel at a third of the HBM roofline: 0.87 ms for 1.7 GB at C3)
        for (; j + 3 < e; j += 4) {
            const int c0 = __ldg(colidx + j), c1 = __ldg(colidx + j + 1), c2 = __ldg(colidx + j + 2),
                      c3 = __ldg(colidx + j + 3);
            const double v0 = __ldg(vals + j), v1 = __ldg(vals + j + 1), v2 = __ldg(vals + j + 2),
                         v3 = __ldg(vals + j + 3);
            const double2 x0 = __ldg(X + (int64_t)c0 * K + r), x1 = __ldg(X + (int64_t)c1 * K + r);
            const double2 x2 = __ldg(X + (int64_t)c2 * K + r), x3 = __ldg(X + (int64_t)c3 * K + r);
            acc.x = fma(v0, x0.x, acc.x);
            acc.y = fma(v0, x0.y, acc.y);
            acc.x = fma(v1, x1.x, acc.x);
            acc.y = fma(v1, x1.y, acc.y);
            acc.x = fma(v2, x2.x, acc.x);
            acc.y = fma(v2, x2.y, acc.y);
            acc.x = fma(v3, x3.x, acc.x);
            acc.y = fma(v3, x3.y, acc.y);
        }
        for (; j < e; ++j) {
            const double v = __ldg(vals + j);
            const double2 x = __ldg(X + (int64_t)__ldg(colidx + j) * K + r);
            acc.x = fma(v, x.x, acc.x);
            acc.y = fma(v, x.y, acc.y);
        }
        if (s) acc = acmul(__ldg(s + i), acc);
        if (a) {
            const double2 t = acmul(__ldg(a + i), __ldg(Z + idx));
            acc.x += t.x;
            acc.y += t.y;
        }
        Y[idx] = acc;
    }
}

// one warp per row k of R (an H1 function): its support (i_s, g_s), s < S, is sorted by column.
// For every OWNED row i_s of A the lanes sweep the row and pick the entries whose column is in the
// support (binary search); fixed summation order -> bit-reproducible.
__global__ void __launch_bounds__(256) galerkin_diag_kernel(int64_t rows, const int32_t *__restrict__ rptr,
                                                            const int32_t *__restrict__ rcol,
                                                            const double *__restrict__ rval, int64_t a_rows,
                                                            const int64_t *__restrict__ a_rowptr,
                                                            const int32_t *__restrict__ a_colidx,
                                                            const double2 *__restrict__ a_vals,
                                                            double2 *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < rows; k += nwarps) {
        const int b = __ldg(rptr + k), S = __ldg(rptr + k + 1) - b;
        const int32_t *sc = rcol + b;
        const double *sv = rval + b;
        double2 acc = make_double2(0.0, 0.0);
        for (int s = 0; s < S; ++s) {
            const int64_t i = __ldg(sc + s);
            if (i >= a_rows) break;  // sorted: the rest are halo rows (owned by other ranks)
            const double gi = __ldg(sv + s);
            const int64_t r0 = __ldg(a_rowptr + i), r1 = __ldg(a_rowptr + i + 1);
            for (int64_t j = r0 + lane; j < r1; j += 32) {
                const int32_t c = __ldg(a_colidx + j);
                int lo = 0, hi = S;  // binary search of c in sc[0..S)
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(sc + mid) < c) lo = mid + 1; else hi = mid;
                }
                if (lo < S && __ldg(sc + lo) == c) {
                    const double w = gi * __ldg(sv + lo);
                    const double2 av = __ldg(a_vals + j);
                    acc.x = fma(w, av.x, acc.x);
                    acc.y = fma(w, av.y, acc.y);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_down_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_down_sync(0xffffffffu, acc.y, o);
        }
        if (lane == 0) out[k] = acc;
    }
}

// d[i] = 1/d[i], 0 where d[i] == 0 or mask[i] == 0 (H1 functions on the Dirichlet boundary)
__global__ void __launch_bounds__(256) masked_reciprocal_kernel(int64_t n, const uint8_t *__restrict__ mask,
                                                                double2 *__restrict__ d) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = d[i];
        const double den = v.x * v.x + v.y * v.y;
        const bool live = den != 0.0 && (!mask || mask[i]);
        d[i] = live ? make_double2(v.x / den, -v.y / den) : make_double2(0.0, 0.0);
    }
}

inline unsigned grid_for(int64_t work, int per_block) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + per_block - 1) / per_block, (int64_t)kNumSMs * 16));
}

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int pg_rcsr_apply(int64_t rows, const int32_t *rowptr, const int32_t *colidx, const double *vals, int k,
                  const double *X, const double *s, const double *a, const double *Z, double *Y, void *stream) {
    PG_REQUIRE(rows >= 0 && (rows == 0 || (rowptr && X && Y)), PG_EINVAL, "pg_rcsr_apply: bad argument");
    PG_REQUIRE(!a || Z, PG_EINVAL, "pg_rcsr_apply: a without Z");
    PG_REQUIRE(k == 1 || k == 2 || k == 4 || k == 8, PG_EINVAL, "pg_rcsr_apply: k = %d (1, 2, 4 or 8)", k);
    if (rows == 0) return PG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = grid_for(rows * k, 256);
    const double2 *X2 = reinterpret_cast<const double2 *>(X), *s2 = reinterpret_cast<const double2 *>(s);
    const double2 *a2 = reinterpret_cast<const double2 *>(a), *Z2 = reinterpret_cast<const double2 *>(Z);
    double2 *Y2 = reinterpret_cast<double2 *>(Y);
    switch (k) {
        case 1: rcsr_apply_kernel<1><<<grid, 256, 0, st>>>(rows, rowptr, colidx, vals, X2, s2, a2, Z2, Y2); break;
        case 2: rcsr_apply_kernel<2><<<grid, 256, 0, st>>>(rows, rowptr, colidx, vals, X2, s2, a2, Z2, Y2); break;
        case 4: rcsr_apply_kernel<4><<<grid, 256, 0, st>>>(rows, rowptr, colidx, vals, X2, s2, a2, Z2, Y2); break;
        default: rcsr_apply_kernel<8><<<grid, 256, 0, st>>>(rows, rowptr, colidx, vals, X2, s2, a2, Z2, Y2);
    }
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_galerkin_diagonal(int64_t rows, const int32_t *r_rowptr, const int32_t *r_colidx, const double *r_vals,
                         int64_t a_rows, const int64_t *a_rowptr, const int32_t *a_colidx, const double *a_vals,
                         double *out, void *stream) {
    PG_REQUIRE(rows >= 0 && (rows == 0 || (r_rowptr && out)), PG_EINVAL, "pg_galerkin_diagonal: bad argument");
    PG_REQUIRE(a_rows == 0 || (a_rowptr && a_colidx && a_vals), PG_EINVAL, "pg_galerkin_diagonal: null matrix");
    if (rows == 0) return PG_OK;
    galerkin_diag_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        rows, r_rowptr, r_colidx, r_vals, a_rows, a_rowptr, a_colidx, reinterpret_cast<const double2 *>(a_vals),
        reinterpret_cast<double2 *>(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_masked_reciprocal(int64_t n, const uint8_t *mask, double *d, void *stream) {
    PG_REQUIRE(n >= 0 && (n == 0 || d), PG_EINVAL, "pg_masked_reciprocal: bad argument");
    if (n == 0) return PG_OK;
    masked_reciprocal_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, mask,
                                                                                 reinterpret_cast<double2 *>(d));
    PG_LAUNCH_OK();
    return PG_OK;
}

}  // extern "C"

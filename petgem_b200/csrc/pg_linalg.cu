// Complex128 CSR SpMV and the vector kernels of the Krylov solve.
// Reference: the arithmetic PETSc performs inside KSP.solve (solver.py:584-590):
// MatMult (CSR SpMV), VecMDot / VecDot / VecNorm, VecMAXPY / VecAXPY / VecAYPX,
// VecPointwiseMult (PCJACOBI), and MatZeroRowsColumns (solver.py:562).
// All reductions are fixed-shape two-stage trees (no atomics): results are
// bit-reproducible run to run.
#include <stdlib.h>

#include "pg_common.cuh"

namespace pg {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(double2 &acc, double2 a, double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfma_conj(double2 &acc, double2 a, double2 b) {  // acc += conj(a)*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
}

// ---------------------------------------------------------------------------
// SpMV: G lanes per row, coalesced (colidx, vals) streams, x gathered through L2
// ---------------------------------------------------------------------------
template <int G, int HINT>
__global__ void __launch_bounds__(256, 6) spmv_kernel(int64_t rows, const int64_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ colidx,
                                                      const double2 *__restrict__ vals,
                                                      const double2 *__restrict__ x,
                                                      const double2 *__restrict__ dscale, double2 *__restrict__ y) {
    const int lane = threadIdx.x % G;
    const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / G;
    const unsigned gm = group_mask<G>();
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    for (int64_t row = grp; row < rows; row += ngrp) {
        const int64_t a = __ldg(rowptr + row), b = __ldg(rowptr + row + 1);
        double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
        // (loading the indices of 8 steps up front, as spmm_kernel does, was measured SLOWER here: C2, p = 1,
        // COCR + Jacobi 0.418 -> 0.483 s; at 40 registers the unconditional 4-wide steps serialise)
        int64_t j = a + lane;
        // 4 independent (index, value, x) streams per lane: the kernel is latency bound otherwise
        for (; j + 3 * G < b; j += 4 * G) {
            const int32_t c0 = ld_stream<HINT>(colidx + j, stream), c1 = ld_stream<HINT>(colidx + j + G, stream);
            const int32_t c2 = ld_stream<HINT>(colidx + j + 2 * G, stream), c3 = ld_stream<HINT>(colidx + j + 3 * G, stream);
            const double2 v0 = ld_stream<HINT>(vals + j, stream), v1 = ld_stream<HINT>(vals + j + G, stream);
            const double2 v2 = ld_stream<HINT>(vals + j + 2 * G, stream), v3 = ld_stream<HINT>(vals + j + 3 * G, stream);
            const double2 x0 = ld_keep<HINT>(x + c0, keep), x1 = ld_keep<HINT>(x + c1, keep);
            const double2 x2 = ld_keep<HINT>(x + c2, keep), x3 = ld_keep<HINT>(x + c3, keep);
            cfma(acc0, v0, x0);
            cfma(acc1, v1, x1);
            cfma(acc0, v2, x2);
            cfma(acc1, v3, x3);
        }
        for (; j < b; j += G) {
            const int32_t c0 = ld_stream<HINT>(colidx + j, stream);
            cfma(acc0, ld_stream<HINT>(vals + j, stream), ld_keep<HINT>(x + c0, keep));
        }
        acc0.x += acc1.x;
        acc0.y += acc1.y;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            acc0.x += __shfl_down_sync(gm, acc0.x, o, G);
            acc0.y += __shfl_down_sync(gm, acc0.y, o, G);
        }
        if (lane == 0) y[row] = dscale ? cmul(__ldg(dscale + row), acc0) : acc0;
    }
}

__global__ void __launch_bounds__(256) diagonal_kernel(int64_t rows, int64_t row_begin,
                                                       const int64_t *__restrict__ rowptr,
                                                       const int32_t *__restrict__ colidx,
                                                       const double2 *__restrict__ vals, double2 *__restrict__ diag) {
    const int lane = threadIdx.x & 7;
    const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    if (grp >= rows) return;
    const int64_t a = rowptr[grp], b = rowptr[grp + 1];
    const int32_t want = (int32_t)(row_begin + grp);
    double2 d = make_double2(0.0, 0.0);
    for (int64_t j = a + lane; j < b; j += 8)
        if (colidx[j] == want) d = vals[j];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        d.x += __shfl_down_sync(group_mask<8>(), d.x, o, 8);
        d.y += __shfl_down_sync(group_mask<8>(), d.y, o, 8);
    }
    if (lane == 0) diag[grp] = d;
}

// MatZeroRowsColumns on an assembled CSR block; 8 lanes per row
__global__ void __launch_bounds__(256) zero_rows_cols_kernel(int64_t rows, int64_t row_begin,
                                                             const int64_t *__restrict__ rowptr,
                                                             const int32_t *__restrict__ colidx,
                                                             const uint8_t *__restrict__ bd, double diag,
                                                             double2 *__restrict__ vals) {
    const int lane = threadIdx.x & 7;
    const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    if (grp >= rows) return;
    const int64_t a = rowptr[grp], b = rowptr[grp + 1];
    const int32_t me = (int32_t)(row_begin + grp);
    const bool rowbd = bd[me];
    for (int64_t j = a + lane; j < b; j += 8) {
        const int32_t c = colidx[j];
        if (rowbd)
            vals[j] = make_double2(c == me ? diag : 0.0, 0.0);
        else if (bd[c])
            vals[j] = make_double2(0.0, 0.0);
    }
}

// ---------------------------------------------------------------------------
// BLAS-1 (complex128), grid-stride, 16-byte accesses
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zaxpy_kernel(int64_t n, const double2 *__restrict__ alpha,
                                                    const double2 *__restrict__ x, double2 *__restrict__ y) {
    const double2 al = *alpha;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 v = y[i];
        cfma(v, al, x[i]);
        y[i] = v;
    }
}

__global__ void __launch_bounds__(256) zaypx_kernel(int64_t n, const double2 *__restrict__ beta,
                                                    const double2 *__restrict__ x, double2 *__restrict__ y) {
    const double2 be = *beta;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 v = x[i];
        cfma(v, be, y[i]);
        y[i] = v;
    }
}

__global__ void __launch_bounds__(256) zaxpbypcz_kernel(int64_t n, const double2 *a, const double2 *x,
                                                        const double2 *b, const double2 *y, const double2 *c,
                                                        const double2 *z, double2 *w) {
    const double2 ca = a ? *a : make_double2(0, 0), cb = b ? *b : make_double2(0, 0),
                  cc = c ? *c : make_double2(0, 0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (a) cfma(v, ca, x[i]);
        if (b) cfma(v, cb, y[i]);
        if (c) cfma(v, cc, z[i]);
        w[i] = v;
    }
}

__global__ void __launch_bounds__(256) zscal_kernel(int64_t n, const double2 *__restrict__ alpha, int inv_real,
                                                    double2 *__restrict__ x) {
    double2 al = *alpha;
    if (inv_real) al = make_double2(1.0 / al.x, 0.0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = cmul(al, x[i]);
}

__global__ void __launch_bounds__(256) zpointwise_kernel(int64_t n, const double2 *__restrict__ x,
                                                         const double2 *__restrict__ y, double2 *__restrict__ z) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = cmul(x[i], y[i]);
}

// ---------------------------------------------------------------------------
// reductions: stage 1 = kRedBlocks blocks of 256 threads, stage 2 = one block
// ---------------------------------------------------------------------------
constexpr int kRedBlocks = kNumSMs * 4;
constexpr int kRedThreads = 256;
constexpr int kDotChunk = 8;  // vectors per pass of VecMDot

template <int K>
__device__ __forceinline__ void block_reduce_store(double2 (&acc)[K], int kcount, double2 *partial, int stride) {
    __shared__ double2 s_red[kRedThreads / 32][K];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc[i].x += __shfl_down_sync(0xffffffffu, acc[i].x, o);
            acc[i].y += __shfl_down_sync(0xffffffffu, acc[i].y, o);
        }
        if (lane == 0) s_red[w][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < K && (int)threadIdx.x < kcount) {
        double2 s = make_double2(0.0, 0.0);
#pragma unroll
        for (int ww = 0; ww < kRedThreads / 32; ++ww) {
            s.x += s_red[ww][threadIdx.x].x;
            s.y += s_red[ww][threadIdx.x].y;
        }
        partial[(int64_t)threadIdx.x * stride + blockIdx.x] = s;
    }
}

// partial[i*kRedBlocks + block] = sum over the block's slice of conj(V_i) * w
template <int K>
__global__ void __launch_bounds__(kRedThreads) mdot_stage1(int64_t n, int kcount, const double2 *__restrict__ V,
                                                           int64_t ldv, const double2 *__restrict__ w,
                                                           double2 *__restrict__ partial) {
    double2 acc[K];
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = make_double2(0.0, 0.0);
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 wj = w[j];
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (i < kcount) cfma_conj(acc[i], V[i * ldv + j], wj);
    }
    block_reduce_store<K>(acc, kcount, partial, kRedBlocks);
}

// unconjugated dot x^T y (COCG on the complex symmetric system)
__global__ void __launch_bounds__(kRedThreads) dotu_stage1(int64_t n, const double2 *__restrict__ x,
                                                           const double2 *__restrict__ y,
                                                           double2 *__restrict__ partial) {
    double2 acc[1] = {make_double2(0.0, 0.0)};
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        cfma(acc[0], x[j], y[j]);
    block_reduce_store<1>(acc, 1, partial, kRedBlocks);
}

__global__ void __launch_bounds__(kRedThreads) nrm2_stage1(int64_t n, const double2 *__restrict__ x,
                                                           double2 *__restrict__ partial) {
    double2 acc[1] = {make_double2(0.0, 0.0)};
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = x[j];
        acc[0].x = fma(v.x, v.x, acc[0].x);
        acc[0].x = fma(v.y, v.y, acc[0].x);
    }
    block_reduce_store<1>(acc, 1, partial, kRedBlocks);
}

// out[i] = sum_b partial[i*kRedBlocks + b], one block per i
__global__ void __launch_bounds__(kRedThreads) reduce_stage2(const double2 *__restrict__ partial,
                                                             double2 *__restrict__ out) {
    __shared__ double2 s[kRedThreads / 32];
    const double2 *p = partial + (int64_t)blockIdx.x * kRedBlocks;
    double2 a = make_double2(0.0, 0.0);
    for (int b = threadIdx.x; b < kRedBlocks; b += kRedThreads) {
        a.x += p[b].x;
        a.y += p[b].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_down_sync(0xffffffffu, a.x, o);
        a.y += __shfl_down_sync(0xffffffffu, a.y, o);
    }
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double2 r = make_double2(0.0, 0.0);
        for (int w = 0; w < kRedThreads / 32; ++w) {
            r.x += s[w].x;
            r.y += s[w].y;
        }
        out[blockIdx.x] = r;
    }
}

// w += scale * sum_i alpha[i] V_i, one pass over w per chunk of K vectors; NORM: also the partial
// sums of |w|^2 of the updated vector (the VecNorm that follows VecMAXPY in GMRES, fused)
template <int K, bool NORM>
__global__ void __launch_bounds__(kRedThreads) maxpy_kernel(int64_t n, int kcount, const double2 *__restrict__ alpha,
                                                            double scale, const double2 *__restrict__ V,
                                                            int64_t ldv, double2 *__restrict__ w,
                                                            double2 *__restrict__ partial) {
    double2 al[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        al[i] = (i < kcount) ? alpha[i] : make_double2(0.0, 0.0);
        al[i].x *= scale;
        al[i].y *= scale;
    }
    double2 nrm[1] = {make_double2(0.0, 0.0)};
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double2 v = w[j];
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (i < kcount) cfma(v, al[i], V[i * ldv + j]);
        w[j] = v;
        if (NORM) {
            nrm[0].x = fma(v.x, v.x, nrm[0].x);
            nrm[0].x = fma(v.y, v.y, nrm[0].x);
        }
    }
    if (NORM) block_reduce_store<1>(nrm, 1, partial, kRedBlocks);
}

// out[0] = +-a/b (and out[1] = -out[0]): Krylov coefficients formed on the device, no host round trip
__global__ void zdiv_kernel(const double2 *__restrict__ a, const double2 *__restrict__ b, int negate,
                            double2 *__restrict__ out) {
    const double2 x = *a, y = *b;
    const double den = y.x * y.x + y.y * y.y;
    double2 q = make_double2((x.x * y.x + x.y * y.y) / den, (x.y * y.x - x.x * y.y) / den);
    if (negate) q = make_double2(-q.x, -q.y);
    out[0] = q;
    out[1] = make_double2(-q.x, -q.y);
}

// y = alpha x, or y = x / Re(alpha) (VecCopy + VecScale of the GMRES normalisation, fused)
__global__ void __launch_bounds__(256) zcopy_scaled_kernel(int64_t n, const double2 *__restrict__ alpha, int inv_real,
                                                           const double2 *__restrict__ x, double2 *__restrict__ y) {
    double2 al = *alpha;
    if (inv_real) al = make_double2(1.0 / al.x, 0.0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = cmul(al, x[i]);
}

static int g_spmv_hint_override = -1;
int spmv_hint_mode() {
    static const int mode = [] {
        const char *e = getenv("PG_SPMV_HINTS");
        // measured (tools/spmv_probe.py, profiles/r2_spmv_l2_probe.jsonl): explicit evict-first on the streams (1)
        // beats ld.cs (0), evict-last x (2) and L1 bypass (3); 256-bit sector loads on top of it (4) win 7-10 %
        return e ? atoi(e) : 4;
    }();
    return g_spmv_hint_override >= 0 ? g_spmv_hint_override : mode;
}
void set_spmv_hint_mode(int mode) { g_spmv_hint_override = mode; }

static inline unsigned ew_grid(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)std::max<int64_t>(1, std::min(b, cap));
}

}  // namespace pg

using namespace pg;
#define D2(p) reinterpret_cast<double2 *>(p)
#define CD2(p) reinterpret_cast<const double2 *>(p)

extern "C" {

int pg_spmv(int64_t rows, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *x,
            double *y, void *stream) {
    return pg_spmv_scaled(rows, rowptr, colidx, vals, x, nullptr, y, stream);
}

int pg_spmv_scaled(int64_t rows, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *x,
                   const double *dscale, double *y, void *stream) {
    PG_REQUIRE(rows >= 0, PG_EINVAL, "pg_spmv: rows < 0");
    if (rows == 0) return PG_OK;
    PG_REQUIRE(rowptr && x && y, PG_EINVAL, "pg_spmv: null pointer");  // colidx/vals may be null when nnz == 0
    cudaStream_t st = (cudaStream_t)stream;
    // lanes per row from the mean row length (host knows it only through the caller: use 16,
    // the best trade-off for the 15..120 nnz/row of p=1..2; long rows are still coalesced)
    constexpr int G = 8;
    int64_t blocks = std::min<int64_t>((rows * G + 255) / 256, (int64_t)kNumSMs * 96);
    switch (spmv_hint_mode()) {
        case 1:
        case 4:  // the CSR kernel reads whole sectors per request already (consecutive lanes, 16 B each)
            spmv_kernel<G, 1><<<(unsigned)blocks, 256, 0, st>>>(rows, rowptr, colidx, CD2(vals), CD2(x), CD2(dscale), D2(y)); break;
        case 2: spmv_kernel<G, 2><<<(unsigned)blocks, 256, 0, st>>>(rows, rowptr, colidx, CD2(vals), CD2(x), CD2(dscale), D2(y)); break;
        case 3: spmv_kernel<G, 3><<<(unsigned)blocks, 256, 0, st>>>(rows, rowptr, colidx, CD2(vals), CD2(x), CD2(dscale), D2(y)); break;
        default: spmv_kernel<G, 0><<<(unsigned)blocks, 256, 0, st>>>(rows, rowptr, colidx, CD2(vals), CD2(x), CD2(dscale), D2(y));
    }
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_csr_diagonal(int64_t rows, int64_t row_begin, const int64_t *rowptr, const int32_t *colidx,
                    const double *vals, double *diag, void *stream) {
    if (rows == 0) return PG_OK;
    PG_REQUIRE(rows > 0 && rowptr && colidx && vals && diag, PG_EINVAL, "pg_csr_diagonal: bad argument");
    diagonal_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rows, row_begin, rowptr, colidx, CD2(vals), D2(diag));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zero_rows_columns(int64_t rows, int64_t row_begin, const int64_t *rowptr, const int32_t *colidx,
                         const uint8_t *bd_mask, double diag, double *vals, void *stream) {
    if (rows == 0) return PG_OK;
    PG_REQUIRE(rows > 0 && rowptr && colidx && bd_mask && vals, PG_EINVAL, "pg_zero_rows_columns: bad argument");
    zero_rows_cols_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rows, row_begin, rowptr, colidx, bd_mask, diag, D2(vals));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zaxpy(int64_t n, const double *alpha, const double *x, double *y, void *stream) {
    if (n == 0) return PG_OK;
    PG_REQUIRE(n > 0 && alpha && x && y, PG_EINVAL, "pg_zaxpy: bad argument");
    zaxpy_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, CD2(alpha), CD2(x), D2(y));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zaypx(int64_t n, const double *beta, const double *x, double *y, void *stream) {
    if (n == 0) return PG_OK;
    PG_REQUIRE(n > 0 && beta && x && y, PG_EINVAL, "pg_zaypx: bad argument");
    zaypx_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, CD2(beta), CD2(x), D2(y));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zaxpbypcz(int64_t n, const double *a, const double *x, const double *b, const double *y, const double *c,
                 const double *z, double *w, void *stream) {
    if (n == 0) return PG_OK;
    PG_REQUIRE(n > 0 && w && (!a || x) && (!b || y) && (!c || z), PG_EINVAL, "pg_zaxpbypcz: bad argument");
    zaxpbypcz_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, CD2(a), CD2(x), CD2(b), CD2(y), CD2(c), CD2(z),
                                                                   D2(w));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zscal(int64_t n, const double *alpha, int inv_real, double *x, void *stream) {
    if (n == 0) return PG_OK;
    PG_REQUIRE(n > 0 && alpha && x, PG_EINVAL, "pg_zscal: bad argument");
    zscal_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, CD2(alpha), inv_real, D2(x));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zpointwise_mult(int64_t n, const double *x, const double *y, double *z, void *stream) {
    if (n == 0) return PG_OK;
    PG_REQUIRE(n > 0 && x && y && z, PG_EINVAL, "pg_zpointwise_mult: bad argument");
    zpointwise_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, CD2(x), CD2(y), D2(z));
    PG_LAUNCH_OK();
    return PG_OK;
}

int64_t pg_reduce_workspace_bytes(int kmax) {
    if (kmax < 1) kmax = 1;
    return (int64_t)kmax * kRedBlocks * 16;
}

int pg_zmdotc(int64_t n, int k, const double *V, int64_t ldv, const double *w, double *out, void *work,
              void *stream) {
    PG_REQUIRE(n >= 0 && k >= 0, PG_EINVAL, "pg_zmdotc: bad size");
    if (k == 0) return PG_OK;
    PG_REQUIRE(V && w && out && work, PG_EINVAL, "pg_zmdotc: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double2 *partial = D2(work);
    for (int c0 = 0; c0 < k; c0 += kDotChunk) {
        const int kc = std::min(kDotChunk, k - c0);
        mdot_stage1<kDotChunk><<<kRedBlocks, kRedThreads, 0, st>>>(n, kc, CD2(V) + (int64_t)c0 * ldv, ldv, CD2(w),
                                                                   partial + (int64_t)c0 * kRedBlocks);
        PG_LAUNCH_OK();
    }
    reduce_stage2<<<k, kRedThreads, 0, st>>>(partial, D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zdotc(int64_t n, const double *x, const double *y, double *out, void *work, void *stream) {
    return pg_zmdotc(n, 1, x, 0, y, out, work, stream);
}

int pg_zdotu(int64_t n, const double *x, const double *y, double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && x && y && out && work, PG_EINVAL, "pg_zdotu: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    dotu_stage1<<<kRedBlocks, kRedThreads, 0, st>>>(n, CD2(x), CD2(y), D2(work));
    PG_LAUNCH_OK();
    reduce_stage2<<<1, kRedThreads, 0, st>>>(D2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_dznrm2sq(int64_t n, const double *x, double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && x && out && work, PG_EINVAL, "pg_dznrm2sq: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    nrm2_stage1<<<kRedBlocks, kRedThreads, 0, st>>>(n, CD2(x), D2(work));
    PG_LAUNCH_OK();
    reduce_stage2<<<1, kRedThreads, 0, st>>>(D2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zmaxpy(int64_t n, int k, const double *alpha, double scale, const double *V, int64_t ldv, double *w,
              void *stream) {
    PG_REQUIRE(n >= 0 && k >= 0, PG_EINVAL, "pg_zmaxpy: bad size");
    if (k == 0 || n == 0) return PG_OK;
    PG_REQUIRE(alpha && V && w, PG_EINVAL, "pg_zmaxpy: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    for (int c0 = 0; c0 < k; c0 += kDotChunk) {
        const int kc = std::min(kDotChunk, k - c0);
        maxpy_kernel<kDotChunk, false><<<ew_grid(n), kRedThreads, 0, st>>>(
            n, kc, CD2(alpha) + c0, scale, CD2(V) + (int64_t)c0 * ldv, ldv, D2(w), nullptr);
        PG_LAUNCH_OK();
    }
    return PG_OK;
}

int pg_zmaxpy_nrm2sq(int64_t n, int k, const double *alpha, double scale, const double *V, int64_t ldv, double *w,
                     double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && k >= 1, PG_EINVAL, "pg_zmaxpy_nrm2sq: bad size");
    PG_REQUIRE(alpha && V && w && out && work, PG_EINVAL, "pg_zmaxpy_nrm2sq: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    for (int c0 = 0; c0 < k; c0 += kDotChunk) {
        const int kc = std::min(kDotChunk, k - c0);
        if (c0 + kDotChunk < k)
            maxpy_kernel<kDotChunk, false><<<ew_grid(n), kRedThreads, 0, st>>>(
                n, kc, CD2(alpha) + c0, scale, CD2(V) + (int64_t)c0 * ldv, ldv, D2(w), nullptr);
        else  // last pass: fixed reduction grid so that the result is reproducible
            maxpy_kernel<kDotChunk, true><<<kRedBlocks, kRedThreads, 0, st>>>(
                n, kc, CD2(alpha) + c0, scale, CD2(V) + (int64_t)c0 * ldv, ldv, D2(w), D2(work));
        PG_LAUNCH_OK();
    }
    reduce_stage2<<<1, kRedThreads, 0, st>>>(D2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zdiv(const double *a, const double *b, int negate, double *out, void *stream) {
    PG_REQUIRE(a && b && out, PG_EINVAL, "pg_zdiv: null pointer");
    zdiv_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(CD2(a), CD2(b), negate, D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zcopy_scaled(int64_t n, const double *alpha, int inv_real, const double *x, double *y, void *stream) {
    if (n == 0) return PG_OK;
    PG_REQUIRE(n > 0 && alpha && x && y, PG_EINVAL, "pg_zcopy_scaled: bad argument");
    zcopy_scaled_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, CD2(alpha), inv_real, CD2(x), D2(y));
    PG_LAUNCH_OK();
    return PG_OK;
}

}  // extern "C"

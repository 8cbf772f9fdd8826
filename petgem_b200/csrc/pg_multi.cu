// Multi-right-hand-side kernels: several sources (or the two MT polarizations) share
// the matrix, so one pass over (colidx, vals) serves K right-hand sides.
// Reference: KSP.solve is called once per right-hand side with the same A (solver.py:584-590);
// here the K solves advance in lockstep and MatMult becomes a sparse matrix times K vectors.
//
// Layout: vectors are INTERLEAVED, X[i*K + r] = entry i of right-hand side r, K in {1, 2, 4, 8}:
// the gather of column c fetches 16*K contiguous bytes, and every element-wise kernel finds the
// right-hand side of element e as e & (K-1).  Reductions are fixed two-stage trees per right-hand
// side (bit-reproducible), scalars stay on the device.
#include <string.h>

#include <algorithm>

#include "pg_common.cuh"

namespace pg {
namespace {

__device__ __forceinline__ double2 mcmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void mcfma(double2 &acc, double2 a, double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

constexpr int kG = 8;  // lanes per row

// Y[row, :] = (dscale[row]) * sum_j vals[j] * X[colidx[j], :]
// Eight lanes per row, one nonzero per lane and step, K accumulators per lane: the K gathers of a
// nonzero are 16*K contiguous bytes.  (K lanes per nonzero with one accumulator each was measured
// slower, 18.0 vs 14.4 ms at C3 with K = 4: the kernel is latency bound and that form keeps fewer
// unique bytes in flight.  For p = 2 the entity-blocked form pg_spmm_blocked halves the gathers.)
template <int K>
__global__ void __launch_bounds__(256) spmm_kernel(int64_t rows, const int64_t *__restrict__ rowptr,
                                                   const int32_t *__restrict__ colidx,
                                                   const double2 *__restrict__ vals, const double2 *__restrict__ X,
                                                   const double2 *__restrict__ dscale, double2 *__restrict__ Y) {
    const int lane = threadIdx.x % kG;
    const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kG;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / kG;
    const unsigned gm = ((1u << kG) - 1u) << ((threadIdx.x & 31) / kG * kG);
    const uint64_t stream = l2_policy_evict_first();
    for (int64_t row = grp; row < rows; row += ngrp) {
        const int64_t a = __ldg(rowptr + row), b = __ldg(rowptr + row + 1);
        double2 acc[K];
#pragma unroll
        for (int r = 0; r < K; ++r) acc[r] = make_double2(0.0, 0.0);
        // The column indices of NPF steps are loaded up front (one latency for all of them), then the steps run two
        // at a time (index, value, gather streams): the gathers no longer wait for an index load each
        // (the plain two-stream loop was latency bound: 0.52 of the copy peak at C4, p = 3, k = 2).
        constexpr int NPF = K <= 2 ? 8 : 4;  // steps whose indices are in flight together
        for (int64_t j = a + lane; j < b; j += NPF * kG) {
            int32_t c[NPF];
#pragma unroll
            for (int q = 0; q < NPF; ++q) c[q] = j + q * kG < b ? ld_stream<1>(colidx + j + q * kG, stream) : -1;
#pragma unroll
            for (int q = 0; q < NPF; q += 2) {
                if (c[q] < 0) break;
                const bool two = c[q + 1] >= 0;
                const int64_t j1 = two ? j + (q + 1) * kG : j + q * kG;  // always a valid position: loads unconditional
                const double2 v0 = ld_stream<1>(vals + j + q * kG, stream);
                double2 v1 = ld_stream<1>(vals + j1, stream);
                if (!two) v1 = make_double2(0.0, 0.0);
                const double2 *x0 = X + (int64_t)c[q] * K, *x1 = X + (int64_t)(two ? c[q + 1] : c[q]) * K;
                double2 xa[K], xb[K];
                if (K >= 2) {  // 32-byte sectors in one request each (see ld256_stream)
#pragma unroll
                    for (int r = 0; r < K; r += 2) {
                        const double2x2 ta = ld256(x0 + r), tb = ld256(x1 + r);
                        xa[r] = ta.a, xa[r + 1 < K ? r + 1 : r] = ta.b, xb[r] = tb.a, xb[r + 1 < K ? r + 1 : r] = tb.b;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < K; ++r) xa[r] = __ldg(x0 + r), xb[r] = __ldg(x1 + r);
                }
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    mcfma(acc[r], v0, xa[r]);
                    mcfma(acc[r], v1, xb[r]);
                }
            }
        }
        // butterfly: every lane ends with the row sums, lane r stores right-hand side r (coalesced)
#pragma unroll
        for (int r = 0; r < K; ++r) {
#pragma unroll
            for (int o = kG / 2; o > 0; o >>= 1) {
                acc[r].x += __shfl_xor_sync(gm, acc[r].x, o, kG);
                acc[r].y += __shfl_xor_sync(gm, acc[r].y, o, kG);
            }
        }
        double2 mine = acc[0];
#pragma unroll
        for (int r = 1; r < K; ++r)
            if (lane == r) mine = acc[r];
        if (lane < K) Y[row * K + lane] = dscale ? mcmul(__ldg(dscale + row), mine) : mine;
    }
}

template <int K>
__global__ void __launch_bounds__(256) zbaxpy_kernel(int64_t ntot, const double2 *__restrict__ alpha,
                                                     const double2 *__restrict__ X, double2 *__restrict__ Y) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double2 al = alpha[i0 & (K - 1)];  // the stride is a multiple of K: fixed right-hand side
    for (int64_t i = i0; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        double2 y = Y[i];
        mcfma(y, al, X[i]);
        Y[i] = y;
    }
}

template <int K>
__global__ void __launch_bounds__(256) zbaypx_kernel(int64_t ntot, const double2 *__restrict__ beta,
                                                     const double2 *__restrict__ X, double2 *__restrict__ Y) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double2 be = beta[i0 & (K - 1)];
    for (int64_t i = i0; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        double2 r = X[i];
        mcfma(r, be, Y[i]);
        Y[i] = r;
    }
}

template <int K>
__global__ void __launch_bounds__(256) zbscale_rows_kernel(int64_t ntot, const double2 *__restrict__ d,
                                                           const double2 *__restrict__ X, double2 *__restrict__ Y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ntot; i += (int64_t)gridDim.x * blockDim.x)
        Y[i] = mcmul(__ldg(d + i / K), X[i]);
}

constexpr int kRedBlocksM = kNumSMs * 4;
constexpr int kRedThreadsM = 256;

// per-right-hand-side block sums of NQ quantities; partial[(q*K + r)*kRedBlocksM + block]
template <int K, int NQ>
__device__ __forceinline__ void block_reduce_store_rhs(double2 (&acc)[NQ], double2 *partial) {
    __shared__ double2 s_red[kRedThreadsM / 32][NQ][K];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
        for (int o = 16; o >= K; o >>= 1) {  // lanes l, l^o share the right-hand side l & (K-1)
            acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
            acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
        }
        if (lane < K) s_red[w][q][lane] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < NQ * K) {
        const int q = threadIdx.x / K, r = threadIdx.x % K;
        double2 s = make_double2(0.0, 0.0);
#pragma unroll
        for (int ww = 0; ww < kRedThreadsM / 32; ++ww) {
            s.x += s_red[ww][q][r].x;
            s.y += s_red[ww][q][r].y;
        }
        partial[(int64_t)(q * K + r) * kRedBlocksM + blockIdx.x] = s;
    }
}

// out[i] = sum_b partial[i*kRedBlocksM + b], one block per i, fixed tree
__global__ void __launch_bounds__(kRedThreadsM) reduce_stage2_m(const double2 *__restrict__ partial,
                                                                double2 *__restrict__ out) {
    __shared__ double2 s[kRedThreadsM / 32];
    const double2 *p = partial + (int64_t)blockIdx.x * kRedBlocksM;
    double2 a = make_double2(0.0, 0.0);
    for (int b = threadIdx.x; b < kRedBlocksM; b += kRedThreadsM) {
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
        for (int w = 0; w < kRedThreadsM / 32; ++w) {
            r.x += s[w].x;
            r.y += s[w].y;
        }
        out[blockIdx.x] = r;
    }
}

// out[r] = sum_i X[i,r] w[i] Y[i,r] (unconjugated; w == nullptr: no weight), or sum |X[i,r]|^2 when
// Y == nullptr
template <int K>
__global__ void __launch_bounds__(kRedThreadsM) zbdot_stage1(int64_t ntot, const double2 *__restrict__ X,
                                                             const double2 *__restrict__ Y,
                                                             const double2 *__restrict__ w,
                                                             double2 *__restrict__ partial) {
    double2 acc[1] = {make_double2(0.0, 0.0)};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 x = X[i];
        if (Y) {
            mcfma(acc[0], x, w ? mcmul(__ldg(w + i / K), Y[i]) : Y[i]);
        } else {
            acc[0].x = fma(x.x, x.x, acc[0].x);
            acc[0].x = fma(x.y, x.y, acc[0].x);
        }
    }
    block_reduce_store_rhs<K, 1>(acc, partial);
}

// One fused pass of the COCG iteration for K right-hand sides:
//   X += alpha P,  R -= alpha Q,  Z = dinv .* R,  partial sums of R^T Z and |Z|^2
template <int K>
__global__ void __launch_bounds__(kRedThreadsM) cocg_step_kernel(int64_t ntot, const double2 *__restrict__ alpha2,
                                                                 const double2 *__restrict__ P,
                                                                 const double2 *__restrict__ Q,
                                                                 const double2 *__restrict__ dinv,
                                                                 double2 *__restrict__ X, double2 *__restrict__ R,
                                                                 double2 *__restrict__ Z,
                                                                 double2 *__restrict__ partial) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double2 al = alpha2[i0 & (K - 1)], nal = alpha2[K + (i0 & (K - 1))];
    double2 acc[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    for (int64_t i = i0; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        double2 x = X[i], r = R[i];
        mcfma(x, al, P[i]);
        mcfma(r, nal, Q[i]);
        X[i] = x;
        R[i] = r;
        const double2 z = dinv ? mcmul(__ldg(dinv + i / K), r) : r;
        Z[i] = z;
        mcfma(acc[0], r, z);
        acc[1].x = fma(z.x, z.x, acc[1].x);
        acc[1].x = fma(z.y, z.y, acc[1].x);
    }
    block_reduce_store_rhs<K, 2>(acc, partial);
}

// COCR (conjugate-orthogonal conjugate residuals), the two element-wise passes of an iteration:
//   update:    X += alpha P,  RT -= alpha dinv .* AP        (RT = M^-1 r, the preconditioned residual)
//   direction: P = RT + beta P,  AP = ART + beta AP         (ART = A RT from the SpMV in between)
template <int K>
__global__ void __launch_bounds__(256) cocr_update_kernel(int64_t ntot, const double2 *__restrict__ alpha2,
                                                          const double2 *__restrict__ P,
                                                          const double2 *__restrict__ AP,
                                                          const double2 *__restrict__ dinv, double2 *__restrict__ X,
                                                          double2 *__restrict__ RT) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double2 al = alpha2[i0 & (K - 1)], nal = alpha2[K + (i0 & (K - 1))];
    for (int64_t i = i0; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        double2 x = X[i], r = RT[i];
        const double2 ap = AP[i];
        mcfma(x, al, P[i]);
        mcfma(r, nal, dinv ? mcmul(__ldg(dinv + i / K), ap) : ap);
        X[i] = x;
        RT[i] = r;
    }
}

template <int K>
__global__ void __launch_bounds__(256) cocr_direction_kernel(int64_t ntot, const double2 *__restrict__ beta,
                                                             const double2 *__restrict__ RT,
                                                             const double2 *__restrict__ ART,
                                                             double2 *__restrict__ P, double2 *__restrict__ AP) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double2 be = beta[i0 & (K - 1)];
    for (int64_t i = i0; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = RT[i], ap = ART[i];
        mcfma(p, be, P[i]);
        mcfma(ap, be, AP[i]);
        P[i] = p;
        AP[i] = ap;
    }
}

// the same pass, and the weighted dot product of the NEW A p with itself that opens the next iteration:
// partial sums of AP[i,r] w[i] AP[i,r] (w = D^-1; fixed reduction grid)
template <int K>
__global__ void __launch_bounds__(kRedThreadsM) cocr_direction_dot_kernel(int64_t ntot, const double2 *__restrict__ beta,
                                                                          const double2 *__restrict__ RT,
                                                                          const double2 *__restrict__ ART,
                                                                          const double2 *__restrict__ w,
                                                                          double2 *__restrict__ P,
                                                                          double2 *__restrict__ AP,
                                                                          double2 *__restrict__ partial) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double2 be = beta[i0 & (K - 1)];
    double2 acc[1] = {make_double2(0.0, 0.0)};
    for (int64_t i = i0; i < ntot; i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = RT[i], ap = ART[i];
        mcfma(p, be, P[i]);
        mcfma(ap, be, AP[i]);
        P[i] = p;
        AP[i] = ap;
        mcfma(acc[0], ap, w ? mcmul(__ldg(w + i / K), ap) : ap);
    }
    block_reduce_store_rhs<K, 1>(acc, partial);
}

// out[r] = a[r] / b[r] (0 when b[r] == 0: an all-zero right-hand side stays zero), out[K + r] = -out[r]
__global__ void zbdiv_kernel(int k, const double2 *__restrict__ a, const double2 *__restrict__ b,
                             double2 *__restrict__ out) {
    const int r = threadIdx.x;
    if (r >= k) return;
    const double2 x = a[r], y = b[r];
    const double den = y.x * y.x + y.y * y.y;
    double2 q = make_double2(0.0, 0.0);
    if (den != 0.0) q = make_double2((x.x * y.x + x.y * y.y) / den, (x.y * y.x - x.x * y.y) / den);
    out[r] = q;
    out[k + r] = make_double2(-q.x, -q.y);
}

inline bool valid_k(int k) { return k == 1 || k == 2 || k == 4 || k == 8; }

}  // namespace
}  // namespace pg

using namespace pg;

#define PG_K_SWITCH(k, CALL)  \
    switch (k) {              \
        case 1: CALL(1); break; \
        case 2: CALL(2); break; \
        case 4: CALL(4); break; \
        case 8: CALL(8); break; \
    }

#define CD2(p) reinterpret_cast<const double2 *>(p)
#define D2(p) reinterpret_cast<double2 *>(p)

extern "C" {

int pg_spmm(int64_t local_rows, const int64_t *rowptr, const int32_t *colidx, const double *vals, int k,
            const double *X, const double *dscale, double *Y, void *stream) {
    PG_REQUIRE(local_rows >= 0 && valid_k(k), PG_EINVAL, "pg_spmm: bad size (rows %lld, k %d: 1, 2, 4 or 8)",
               (long long)local_rows, k);
    if (local_rows == 0) return PG_OK;
    PG_REQUIRE(rowptr && X && Y, PG_EINVAL, "pg_spmm: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t blocks = std::min<int64_t>((local_rows * kG + 255) / 256, (int64_t)kNumSMs * 8);
#define CALL(KK) spmm_kernel<KK><<<(unsigned)blocks, 256, 0, st>>>(local_rows, rowptr, colidx, CD2(vals), CD2(X), CD2(dscale), D2(Y))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    return PG_OK;
}

static inline unsigned ew_blocks(int64_t ntot) {
    // grid stride must stay a multiple of K: 256 threads x any block count is
    return (unsigned)std::min<int64_t>((ntot + 255) / 256, (int64_t)kNumSMs * 8);
}

int pg_zbaxpy(int64_t n, int k, const double *alpha, const double *X, double *Y, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_zbaxpy: bad size");
    if (n == 0) return PG_OK;
    PG_REQUIRE(alpha && X && Y, PG_EINVAL, "pg_zbaxpy: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) zbaxpy_kernel<KK><<<ew_blocks(n * k), 256, 0, st>>>(n * k, CD2(alpha), CD2(X), D2(Y))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zbaypx(int64_t n, int k, const double *beta, const double *X, double *Y, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_zbaypx: bad size");
    if (n == 0) return PG_OK;
    PG_REQUIRE(beta && X && Y, PG_EINVAL, "pg_zbaypx: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) zbaypx_kernel<KK><<<ew_blocks(n * k), 256, 0, st>>>(n * k, CD2(beta), CD2(X), D2(Y))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zbscale_rows(int64_t n, int k, const double *d, const double *X, double *Y, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_zbscale_rows: bad size");
    if (n == 0) return PG_OK;
    PG_REQUIRE(d && X && Y, PG_EINVAL, "pg_zbscale_rows: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) zbscale_rows_kernel<KK><<<ew_blocks(n * k), 256, 0, st>>>(n * k, CD2(d), CD2(X), D2(Y))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zbdotu(int64_t n, int k, const double *X, const double *Y, double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_zbdotu: bad size");
    PG_REQUIRE(X && Y && out && work, PG_EINVAL, "pg_zbdotu: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) zbdot_stage1<KK><<<kRedBlocksM, kRedThreadsM, 0, st>>>(n * k, CD2(X), CD2(Y), nullptr, D2(work))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    reduce_stage2_m<<<k, kRedThreadsM, 0, st>>>(CD2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zbdotu_w(int64_t n, int k, const double *X, const double *Y, const double *w, double *out, void *work,
                void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_zbdotu_w: bad size");
    PG_REQUIRE(X && Y && out && work, PG_EINVAL, "pg_zbdotu_w: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) zbdot_stage1<KK><<<kRedBlocksM, kRedThreadsM, 0, st>>>(n * k, CD2(X), CD2(Y), CD2(w), D2(work))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    reduce_stage2_m<<<k, kRedThreadsM, 0, st>>>(CD2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_cocr_update(int64_t n, int k, const double *alpha2, const double *P, const double *AP, const double *dinv,
                   double *X, double *RT, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_cocr_update: bad size");
    if (n == 0) return PG_OK;
    PG_REQUIRE(alpha2 && P && AP && X && RT, PG_EINVAL, "pg_cocr_update: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) cocr_update_kernel<KK><<<ew_blocks(n * k), 256, 0, st>>>(n * k, CD2(alpha2), CD2(P), CD2(AP), CD2(dinv), D2(X), D2(RT))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_cocr_direction(int64_t n, int k, const double *beta, const double *RT, const double *ART, double *P,
                      double *AP, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_cocr_direction: bad size");
    if (n == 0) return PG_OK;
    PG_REQUIRE(beta && RT && ART && P && AP, PG_EINVAL, "pg_cocr_direction: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) cocr_direction_kernel<KK><<<ew_blocks(n * k), 256, 0, st>>>(n * k, CD2(beta), CD2(RT), CD2(ART), D2(P), D2(AP))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_cocr_direction_dot(int64_t n, int k, const double *beta, const double *RT, const double *ART, const double *w,
                          double *P, double *AP, double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_cocr_direction_dot: bad size");
    PG_REQUIRE(beta && RT && ART && P && AP && out && work, PG_EINVAL, "pg_cocr_direction_dot: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) cocr_direction_dot_kernel<KK><<<kRedBlocksM, kRedThreadsM, 0, st>>>(n * k, CD2(beta), CD2(RT), CD2(ART), CD2(w), D2(P), D2(AP), D2(work))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    reduce_stage2_m<<<k, kRedThreadsM, 0, st>>>(CD2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zbnrm2sq(int64_t n, int k, const double *X, double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_zbnrm2sq: bad size");
    PG_REQUIRE(X && out && work, PG_EINVAL, "pg_zbnrm2sq: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) zbdot_stage1<KK><<<kRedBlocksM, kRedThreadsM, 0, st>>>(n * k, CD2(X), nullptr, nullptr, D2(work))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    reduce_stage2_m<<<k, kRedThreadsM, 0, st>>>(CD2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_cocg_step(int64_t n, int k, const double *alpha2, const double *P, const double *Q, const double *dinv,
                 double *X, double *R, double *Z, double *out, void *work, void *stream) {
    PG_REQUIRE(n >= 0 && valid_k(k), PG_EINVAL, "pg_cocg_step: bad size");
    PG_REQUIRE(alpha2 && P && Q && X && R && Z && out && work, PG_EINVAL, "pg_cocg_step: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(KK) cocg_step_kernel<KK><<<kRedBlocksM, kRedThreadsM, 0, st>>>(n * k, CD2(alpha2), CD2(P), CD2(Q), CD2(dinv), D2(X), D2(R), D2(Z), D2(work))
    PG_K_SWITCH(k, CALL)
#undef CALL
    PG_LAUNCH_OK();
    reduce_stage2_m<<<2 * k, kRedThreadsM, 0, st>>>(CD2(work), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_zbdiv(int k, const double *a, const double *b, double *out, void *stream) {
    PG_REQUIRE(valid_k(k), PG_EINVAL, "pg_zbdiv: bad k %d", k);
    PG_REQUIRE(a && b && out, PG_EINVAL, "pg_zbdiv: null pointer");
    zbdiv_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(k, CD2(a), CD2(b), D2(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

// ---- CUDA graph of a batch of launches -----------------------------------------------------------
// The Krylov drivers issue ~10 small launches per iteration; between two host checks of the residual
// nothing depends on the host, so the batch is captured once and replayed (stream must not be the
// legacy default stream).
int pg_graph_begin(void *stream) {
    PG_REQUIRE(stream, PG_EINVAL, "pg_graph_begin: capture needs a non-default stream");
    PG_CUDA_OK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return PG_OK;
}

int pg_graph_end(void *stream, void **graph_exec) {
    PG_REQUIRE(stream && graph_exec, PG_EINVAL, "pg_graph_end: null pointer");
    cudaGraph_t g = nullptr;
    *graph_exec = nullptr;
    PG_CUDA_OK(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    PG_CUDA_OK(e);
    *graph_exec = ge;
    return PG_OK;
}

int pg_graph_launch(void *graph_exec, void *stream) {
    PG_REQUIRE(graph_exec, PG_EINVAL, "pg_graph_launch: null graph");
    PG_CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return PG_OK;
}

void pg_graph_destroy(void *graph_exec) {
    if (graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)graph_exec);
}

// ---- L2 residency of the vector MatMult gathers from --------------------------------------------------------
int pg_tune_spmv_hints(int mode) {
    PG_REQUIRE(mode >= -1 && mode <= 4, PG_EINVAL, "pg_tune_spmv_hints: mode %d (-1 = environment default, 0..4)", mode);
    set_spmv_hint_mode(mode);
    return PG_OK;
}

int pg_l2_fetch_granularity(int bytes) {
    PG_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, PG_EINVAL, "pg_l2_fetch_granularity: %d (32, 64, 128)", bytes);
    PG_CUDA_OK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return PG_OK;
}

int pg_l2_persist(const void *ptr, int64_t bytes, double hit_ratio, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (!ptr || bytes <= 0) {  // clear the window and release the persisting lines
        attr.accessPolicyWindow.base_ptr = nullptr;
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitRatio = 0.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        PG_CUDA_OK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
        PG_CUDA_OK(cudaCtxResetPersistingL2Cache());
        PG_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));  // give the set-aside back to the normal L2
        return PG_OK;
    }
    int dev = 0, max_persist = 0, max_window = 0;
    PG_CUDA_OK(cudaGetDevice(&dev));
    PG_CUDA_OK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    PG_CUDA_OK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    PG_REQUIRE(max_persist > 0 && max_window > 0, PG_ECUDA, "pg_l2_persist: the device has no persisting L2");
    const size_t window = (size_t)std::min<int64_t>(bytes, max_window);
    PG_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
    attr.accessPolicyWindow.num_bytes = window;
    // lines of the window that are kept: all of it when it fits the set-aside, else that fraction
    float ratio = hit_ratio > 0.0 ? (float)hit_ratio : std::min(1.0f, (float)max_persist / (float)window);
    attr.accessPolicyWindow.hitRatio = ratio;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    PG_CUDA_OK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
    return PG_OK;
}

int64_t pg_l2_persist_capacity(void) {
    int dev = 0, max_persist = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) return 0;
    return max_persist;
}

}  // extern "C"

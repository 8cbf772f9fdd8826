// Entity-blocked SpMV for p = 2.
//
// The assembled matrix has more structure than CSR shows: the R = 2 rows of an entity share one
// column list, and the columns come in pairs (the 2 dofs of a column entity).  Reading the plan's
// per-entity column-entity list instead of colidx costs 4 bytes per 2x2 block instead of 16:
// 17 B per nonzero instead of 20, half the x gathers, a third of the load instructions.
// Same arithmetic order per row as the CSR kernel up to the lane assignment; MatMult semantics
// (solver.py:589, inside KSP).
#include <stdlib.h>

#include <algorithm>

#include "pg_plan.cuh"

namespace pg {

__device__ __forceinline__ void cfma2(double2 &acc, double2 a, double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

constexpr int kBG = 8;  // lanes per entity

template <int HINT>
__global__ void __launch_bounds__(256, 6) spmv_blocked2_kernel(int64_t nb, const EntHdr *__restrict__ hdr,
                                                               const int32_t *__restrict__ colstart,
                                                               const double2 *__restrict__ vals,
                                                               const double2 *__restrict__ x,
                                                               const double2 *__restrict__ dscale,
                                                               double2 *__restrict__ y, int chunk) {
    const int lane = threadIdx.x % kBG;
    // In-order grid: block b takes the `chunk` consecutive tiles of 32 entities starting at b*chunk, so a
    // single front moves through the matrix and the x entries shared by neighbouring rows are still in
    // L2 when they are needed again (a capped grid-stride launch runs ~12 interleaved sweeps instead:
    // measured at C3, SpMV 6.17 -> 5.19 ms, four right-hand sides 10.8 -> 9.2 ms).
    const int gpb = blockDim.x / kBG;
    const unsigned gm = ((1u << kBG) - 1u) << ((threadIdx.x & 31) / kBG * kBG);
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    for (int c = 0; c < chunk; ++c) {
        const int64_t i = (blockIdx.x * (int64_t)chunk + c) * gpb + threadIdx.x / kBG;
        if (i >= nb) break;
        const int4 *hp = reinterpret_cast<const int4 *>(hdr + i);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const int64_t valoff = ((int64_t)(unsigned)h0.x) | ((int64_t)h0.y << 32);
        const int L = (h1.x >> 16) & 0xffff;
        const int row = h1.z, cbase = h1.w;
        const int nc = L >> 1;
        const double2 *v0 = vals + valoff, *v1 = v0 + L;
        const int32_t *cs = colstart + cbase;
        double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
        int j = lane;
        for (; j + kBG < nc; j += 2 * kBG) {  // two column entities per lane in flight
            const int32_t c0 = ld_stream<HINT>(cs + j, stream), c1 = ld_stream<HINT>(cs + j + kBG, stream);
            const double2 p00 = ld_stream<HINT>(v0 + 2 * j, stream), p01 = ld_stream<HINT>(v0 + 2 * j + 1, stream);
            const double2 p10 = ld_stream<HINT>(v1 + 2 * j, stream), p11 = ld_stream<HINT>(v1 + 2 * j + 1, stream);
            const double2 q00 = ld_stream<HINT>(v0 + 2 * (j + kBG), stream), q01 = ld_stream<HINT>(v0 + 2 * (j + kBG) + 1, stream);
            const double2 q10 = ld_stream<HINT>(v1 + 2 * (j + kBG), stream), q11 = ld_stream<HINT>(v1 + 2 * (j + kBG) + 1, stream);
            const double2 x0 = ld_keep<HINT>(x + c0, keep), x1 = ld_keep<HINT>(x + c0 + 1, keep);
            const double2 z0 = ld_keep<HINT>(x + c1, keep), z1 = ld_keep<HINT>(x + c1 + 1, keep);
            cfma2(a0, p00, x0);
            cfma2(a1, p10, x0);
            cfma2(a0, p01, x1);
            cfma2(a1, p11, x1);
            cfma2(a0, q00, z0);
            cfma2(a1, q10, z0);
            cfma2(a0, q01, z1);
            cfma2(a1, q11, z1);
        }
        if (j < nc) {
            const int32_t c0 = ld_stream<HINT>(cs + j, stream);
            const double2 p00 = ld_stream<HINT>(v0 + 2 * j, stream), p01 = ld_stream<HINT>(v0 + 2 * j + 1, stream);
            const double2 p10 = ld_stream<HINT>(v1 + 2 * j, stream), p11 = ld_stream<HINT>(v1 + 2 * j + 1, stream);
            const double2 x0 = ld_keep<HINT>(x + c0, keep), x1 = ld_keep<HINT>(x + c0 + 1, keep);
            cfma2(a0, p00, x0);
            cfma2(a1, p10, x0);
            cfma2(a0, p01, x1);
            cfma2(a1, p11, x1);
        }
#pragma unroll
        for (int o = kBG / 2; o > 0; o >>= 1) {
            a0.x += __shfl_down_sync(gm, a0.x, o, kBG);
            a0.y += __shfl_down_sync(gm, a0.y, o, kBG);
            a1.x += __shfl_down_sync(gm, a1.x, o, kBG);
            a1.y += __shfl_down_sync(gm, a1.y, o, kBG);
        }
        if (lane == 0) {
            if (dscale) {
                const double2 d0 = __ldg(dscale + row), d1 = __ldg(dscale + row + 1);
                a0 = make_double2(d0.x * a0.x - d0.y * a0.y, d0.x * a0.y + d0.y * a0.x);
                a1 = make_double2(d1.x * a1.x - d1.y * a1.y, d1.x * a1.y + d1.y * a1.x);
            }
            y[row] = a0;
            y[row + 1] = a1;
        }
    }
}

// The same kernel with 256-bit loads: every 2x1 block of values and every (x[c], x[c+1]) pair is ONE request for
// ONE 32-byte sector (all of them are 32-byte aligned: rows hold an even number of entries, entities start on
// even dofs, the halo keeps the two dofs of a column entity adjacent).
// DOT: the block also leaves sum_rows x[row] * y[row] (unconjugated; x[row] = the multiplied vector's own entry)
// of its rows in dot_partial[blockIdx.x]: the x^T A x of COCG / COCR without a second pass over x and y.
template <bool DOT>
__global__ void __launch_bounds__(256, 6) spmv_blocked2_v256_kernel(int64_t nb, const EntHdr *__restrict__ hdr,
                                                                    const int32_t *__restrict__ colstart,
                                                                    const double2 *__restrict__ vals,
                                                                    const double2 *__restrict__ x,
                                                                    const double2 *__restrict__ dscale,
                                                                    double2 *__restrict__ y, int chunk,
                                                                    double2 *__restrict__ dot_partial) {
    const int lane = threadIdx.x % kBG;
    const int gpb = blockDim.x / kBG;  // in-order grid, see spmv_blocked2_kernel
    const unsigned gm = ((1u << kBG) - 1u) << ((threadIdx.x & 31) / kBG * kBG);
    const uint64_t stream = l2_policy_evict_first();
    double2 pd = make_double2(0.0, 0.0);
    for (int c = 0; c < chunk; ++c) {
        const int64_t i = (blockIdx.x * (int64_t)chunk + c) * gpb + threadIdx.x / kBG;
        if (i >= nb) break;
        const int4 *hp = reinterpret_cast<const int4 *>(hdr + i);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const int64_t valoff = ((int64_t)(unsigned)h0.x) | ((int64_t)h0.y << 32);
        const int L = (h1.x >> 16) & 0xffff;
        const int row = h1.z, cbase = h1.w;
        const int nc = L >> 1;
        const double2 *v0 = vals + valoff, *v1 = v0 + L;
        const int32_t *cs = colstart + cbase;
        double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
        int j = lane;
        for (; j + kBG < nc; j += 2 * kBG) {  // two column entities per lane in flight
            const int32_t c0 = ld_stream<1>(cs + j, stream), c1 = ld_stream<1>(cs + j + kBG, stream);
            const double2x2 p0 = ld256_stream(v0 + 2 * j, stream), p1 = ld256_stream(v1 + 2 * j, stream);
            const double2x2 q0 = ld256_stream(v0 + 2 * (j + kBG), stream), q1 = ld256_stream(v1 + 2 * (j + kBG), stream);
            const double2x2 xx = ld256(x + c0), zz = ld256(x + c1);
            cfma2(a0, p0.a, xx.a);
            cfma2(a1, p1.a, xx.a);
            cfma2(a0, p0.b, xx.b);
            cfma2(a1, p1.b, xx.b);
            cfma2(a0, q0.a, zz.a);
            cfma2(a1, q1.a, zz.a);
            cfma2(a0, q0.b, zz.b);
            cfma2(a1, q1.b, zz.b);
        }
        if (j < nc) {
            const int32_t c0 = ld_stream<1>(cs + j, stream);
            const double2x2 p0 = ld256_stream(v0 + 2 * j, stream), p1 = ld256_stream(v1 + 2 * j, stream);
            const double2x2 xx = ld256(x + c0);
            cfma2(a0, p0.a, xx.a);
            cfma2(a1, p1.a, xx.a);
            cfma2(a0, p0.b, xx.b);
            cfma2(a1, p1.b, xx.b);
        }
#pragma unroll
        for (int o = kBG / 2; o > 0; o >>= 1) {
            a0.x += __shfl_down_sync(gm, a0.x, o, kBG);
            a0.y += __shfl_down_sync(gm, a0.y, o, kBG);
            a1.x += __shfl_down_sync(gm, a1.x, o, kBG);
            a1.y += __shfl_down_sync(gm, a1.y, o, kBG);
        }
        if (lane == 0) {
            if (dscale) {
                const double2 d0 = __ldg(dscale + row), d1 = __ldg(dscale + row + 1);
                a0 = make_double2(d0.x * a0.x - d0.y * a0.y, d0.x * a0.y + d0.y * a0.x);
                a1 = make_double2(d1.x * a1.x - d1.y * a1.y, d1.x * a1.y + d1.y * a1.x);
            }
            y[row] = a0;
            y[row + 1] = a1;
            if (DOT) {
                const double2x2 xr = ld256(x + row);
                cfma2(pd, xr.a, a0);
                cfma2(pd, xr.b, a1);
            }
        }
    }
    if (DOT) {  // fixed order: lanes of a warp, warps of the block
        __shared__ double2 s_dot[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pd.x += __shfl_xor_sync(0xffffffffu, pd.x, o);
            pd.y += __shfl_xor_sync(0xffffffffu, pd.y, o);
        }
        if ((threadIdx.x & 31) == 0) s_dot[threadIdx.x >> 5] = pd;
        __syncthreads();
        if (threadIdx.x == 0) {
            double2 t = s_dot[0];
            for (int w = 1; w < 8; ++w) {
                t.x += s_dot[w].x;
                t.y += s_dot[w].y;
            }
            dot_partial[blockIdx.x] = t;
        }
    }
}

// Sum of the per-block partials, two fixed-order levels (the C3 matrix has ~500 000 blocks: one block alone would
// crawl through 8 MB at the latency of a single CTA).  Level a: kDotRed blocks, block g sums the partials
// b = g*per .. (g+1)*per of right-hand side r = blockIdx.y into mid[r*kDotRed + g]; level b: one block per
// right-hand side sums the kDotRed values.
constexpr int kDotRed = kNumSMs * 4;

__device__ __forceinline__ double2 block_sum_256(double2 a) {
    __shared__ double2 s[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_down_sync(0xffffffffu, a.x, o);
        a.y += __shfl_down_sync(0xffffffffu, a.y, o);
    }
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    double2 t = s[0];
    for (int w = 1; w < 8; ++w) {
        t.x += s[w].x;
        t.y += s[w].y;
    }
    return t;  // valid in every thread
}

__global__ void __launch_bounds__(256) dot_stage2a_var(const double2 *__restrict__ partial, int64_t nblocks, int K,
                                                       double2 *__restrict__ mid) {
    const int r = blockIdx.y;
    const int64_t per = (nblocks + kDotRed - 1) / kDotRed;
    const int64_t b0 = blockIdx.x * per, b1 = min(b0 + per, nblocks);
    double2 a = make_double2(0.0, 0.0);
    for (int64_t b = b0 + threadIdx.x; b < b1; b += 256) {
        const double2 v = partial[b * K + r];
        a.x += v.x;
        a.y += v.y;
    }
    const double2 t = block_sum_256(a);
    if (threadIdx.x == 0) mid[(int64_t)r * kDotRed + blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) dot_stage2b_var(const double2 *__restrict__ mid, double2 *__restrict__ out) {
    const int r = blockIdx.x;
    double2 a = make_double2(0.0, 0.0);
    for (int g = threadIdx.x; g < kDotRed; g += 256) {
        const double2 v = mid[(int64_t)r * kDotRed + g];
        a.x += v.x;
        a.y += v.y;
    }
    const double2 t = block_sum_256(a);
    if (threadIdx.x == 0) out[r] = t;
}

// The same for K interleaved right-hand sides (X[i*K + r], see pg_multi.cu): K lanes share a 2x2 block,
// lane r of them gathers X[c0, r] and X[c0+1, r] -- together the 32*K contiguous bytes of the column
// entity, which serve 4 nonzeros x K right-hand sides (8 B gathered per nonzero and right-hand side instead
// of 16) -- and the 8/K lane sets of a group take alternate column entities.
template <int K>
__global__ void __launch_bounds__(256) spmm_blocked2_kernel(int64_t nb, const EntHdr *__restrict__ hdr,
                                                            const int32_t *__restrict__ colstart,
                                                            const double2 *__restrict__ vals,
                                                            const double2 *__restrict__ X,
                                                            const double2 *__restrict__ dscale,
                                                            double2 *__restrict__ Y, int chunk) {
    constexpr int NS = kBG / K;  // column entities per step of a group
    const int lane = threadIdx.x % kBG;
    const int r = lane % K, sub = lane / K;
    const int gpb = blockDim.x / kBG;  // in-order grid, see spmv_blocked2_kernel
    const unsigned gm = ((1u << kBG) - 1u) << ((threadIdx.x & 31) / kBG * kBG);
    const uint64_t stream = l2_policy_evict_first();
    const double2 *Xr = X + r;
    for (int c = 0; c < chunk; ++c) {
        const int64_t i = (blockIdx.x * (int64_t)chunk + c) * gpb + threadIdx.x / kBG;
        if (i >= nb) break;
        const int4 *hp = reinterpret_cast<const int4 *>(hdr + i);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const int64_t valoff = ((int64_t)(unsigned)h0.x) | ((int64_t)h0.y << 32);
        const int L = (h1.x >> 16) & 0xffff;
        const int row = h1.z, cbase = h1.w;
        const int nc = L >> 1;
        const double2 *v0 = vals + valoff, *v1 = v0 + L;
        const int32_t *cs = colstart + cbase;
        double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
        int j = sub;
        for (; j + NS < nc; j += 2 * NS) {  // two column entities per lane in flight
            const int32_t c0 = ld_stream<1>(cs + j, stream), c1 = ld_stream<1>(cs + j + NS, stream);
            const double2 p00 = ld_stream<1>(v0 + 2 * j, stream), p01 = ld_stream<1>(v0 + 2 * j + 1, stream);
            const double2 p10 = ld_stream<1>(v1 + 2 * j, stream), p11 = ld_stream<1>(v1 + 2 * j + 1, stream);
            const double2 q00 = ld_stream<1>(v0 + 2 * (j + NS), stream), q01 = ld_stream<1>(v0 + 2 * (j + NS) + 1, stream);
            const double2 q10 = ld_stream<1>(v1 + 2 * (j + NS), stream), q11 = ld_stream<1>(v1 + 2 * (j + NS) + 1, stream);
            const double2 x0 = __ldg(Xr + (int64_t)c0 * K), x1 = __ldg(Xr + ((int64_t)c0 + 1) * K);
            const double2 z0 = __ldg(Xr + (int64_t)c1 * K), z1 = __ldg(Xr + ((int64_t)c1 + 1) * K);
            cfma2(a0, p00, x0);
            cfma2(a1, p10, x0);
            cfma2(a0, p01, x1);
            cfma2(a1, p11, x1);
            cfma2(a0, q00, z0);
            cfma2(a1, q10, z0);
            cfma2(a0, q01, z1);
            cfma2(a1, q11, z1);
        }
        if (j < nc) {
            const int32_t c0 = ld_stream<1>(cs + j, stream);
            const double2 p00 = ld_stream<1>(v0 + 2 * j, stream), p01 = ld_stream<1>(v0 + 2 * j + 1, stream);
            const double2 p10 = ld_stream<1>(v1 + 2 * j, stream), p11 = ld_stream<1>(v1 + 2 * j + 1, stream);
            const double2 x0 = __ldg(Xr + (int64_t)c0 * K), x1 = __ldg(Xr + ((int64_t)c0 + 1) * K);
            cfma2(a0, p00, x0);
            cfma2(a1, p10, x0);
            cfma2(a0, p01, x1);
            cfma2(a1, p11, x1);
        }
#pragma unroll
        for (int o = K; o < kBG; o <<= 1) {  // sum over the lane sets that share right-hand side r
            a0.x += __shfl_xor_sync(gm, a0.x, o, kBG);
            a0.y += __shfl_xor_sync(gm, a0.y, o, kBG);
            a1.x += __shfl_xor_sync(gm, a1.x, o, kBG);
            a1.y += __shfl_xor_sync(gm, a1.y, o, kBG);
        }
        if (sub == 0) {
            if (dscale) {
                const double2 d0 = __ldg(dscale + row), d1 = __ldg(dscale + row + 1);
                a0 = make_double2(d0.x * a0.x - d0.y * a0.y, d0.x * a0.y + d0.y * a0.x);
                a1 = make_double2(d1.x * a1.x - d1.y * a1.y, d1.x * a1.y + d1.y * a1.x);
            }
            Y[(int64_t)row * K + r] = a0;
            Y[((int64_t)row + 1) * K + r] = a1;
        }
    }
}

// Same work assignment, different schedule: the gathers of X depend on the column-entity indices, and with the
// plain loop every step pays two dependent memory latencies (index, then X).  Here the 8 lanes of a group load
// up to 32 indices of the entity at once (coalesced) and every step takes its index from a register by
// shuffle, so all X gathers and value loads of the entity are independent of any in-flight load.
template <int K, int HINT, bool DOT = false>
__global__ void __launch_bounds__(256) spmm_blocked2_pf_kernel(int64_t nb, const EntHdr *__restrict__ hdr,
                                                               const int32_t *__restrict__ colstart,
                                                               const double2 *__restrict__ vals,
                                                               const double2 *__restrict__ X,
                                                               const double2 *__restrict__ dscale,
                                                               double2 *__restrict__ Y, int chunk,
                                                               double2 *__restrict__ dot_partial = nullptr) {
    double2 pd = make_double2(0.0, 0.0);  // DOT: sum_rows X[row, r] * Y[row, r] of this thread's right-hand side
    constexpr int NS = kBG / K;  // column entities per step of a group
    const int lane = threadIdx.x % kBG;
    const int r = lane % K, sub = lane / K;
    const int gpb = blockDim.x / kBG;  // in-order grid, see spmv_blocked2_kernel
    const unsigned gm = ((1u << kBG) - 1u) << ((threadIdx.x & 31) / kBG * kBG);
    const uint64_t stream = l2_policy_evict_first();
    const double2 *Xr = X + r;
    for (int c = 0; c < chunk; ++c) {
        const int64_t i = (blockIdx.x * (int64_t)chunk + c) * gpb + threadIdx.x / kBG;
        if (i >= nb) break;
        const int4 *hp = reinterpret_cast<const int4 *>(hdr + i);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const int64_t valoff = ((int64_t)(unsigned)h0.x) | ((int64_t)h0.y << 32);
        const int L = (h1.x >> 16) & 0xffff;
        const int row = h1.z, cbase = h1.w;
        const int nc = L >> 1;
        const double2 *v0 = vals + valoff, *v1 = v0 + L;
        const int32_t *cs = colstart + cbase;
        double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
        for (int base = 0; base < nc; base += 4 * kBG) {
            int32_t idx[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = base + q * kBG + lane;
                idx[q] = j < nc ? ld_stream<(HINT == 4 ? 1 : HINT)>(cs + j, stream) : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (base + q * kBG >= nc) break;  // uniform over the group
#pragma unroll
                for (int s0 = 0; s0 < kBG; s0 += NS) {
                    const int jl = s0 + sub;                       // position inside this 8-block of column entities
                    const int32_t c0 = __shfl_sync(gm, idx[q], jl, kBG);
                    const int j = base + q * kBG + jl;
                    if (j < nc) {
                        double2 p00, p01, p10, p11;
                        if (HINT == 4) {  // one request per 32-byte sector of values
                            const double2x2 p0 = ld256_stream(v0 + 2 * j, stream), p1 = ld256_stream(v1 + 2 * j, stream);
                            p00 = p0.a, p01 = p0.b, p10 = p1.a, p11 = p1.b;
                        } else {
                            p00 = ld_stream<HINT>(v0 + 2 * j, stream), p01 = ld_stream<HINT>(v0 + 2 * j + 1, stream);
                            p10 = ld_stream<HINT>(v1 + 2 * j, stream), p11 = ld_stream<HINT>(v1 + 2 * j + 1, stream);
                        }
                        const double2 x0 = __ldg(Xr + (int64_t)c0 * K), x1 = __ldg(Xr + ((int64_t)c0 + 1) * K);
                        cfma2(a0, p00, x0);
                        cfma2(a1, p10, x0);
                        cfma2(a0, p01, x1);
                        cfma2(a1, p11, x1);
                    }
                }
            }
        }
#pragma unroll
        for (int o = K; o < kBG; o <<= 1) {  // sum over the lane sets that share right-hand side r
            a0.x += __shfl_xor_sync(gm, a0.x, o, kBG);
            a0.y += __shfl_xor_sync(gm, a0.y, o, kBG);
            a1.x += __shfl_xor_sync(gm, a1.x, o, kBG);
            a1.y += __shfl_xor_sync(gm, a1.y, o, kBG);
        }
        if (sub == 0) {
            if (dscale) {
                const double2 d0 = __ldg(dscale + row), d1 = __ldg(dscale + row + 1);
                a0 = make_double2(d0.x * a0.x - d0.y * a0.y, d0.x * a0.y + d0.y * a0.x);
                a1 = make_double2(d1.x * a1.x - d1.y * a1.y, d1.x * a1.y + d1.y * a1.x);
            }
            Y[(int64_t)row * K + r] = a0;
            Y[((int64_t)row + 1) * K + r] = a1;
            if (DOT) {
                cfma2(pd, __ldg(X + (int64_t)row * K + r), a0);
                cfma2(pd, __ldg(X + ((int64_t)row + 1) * K + r), a1);
            }
        }
    }
    if (DOT) {
        __shared__ double2 s_dot[8][K];
#pragma unroll
        for (int o = 16; o >= K; o >>= 1) {  // lanes l, l ^ o serve the same right-hand side l % K
            pd.x += __shfl_xor_sync(0xffffffffu, pd.x, o);
            pd.y += __shfl_xor_sync(0xffffffffu, pd.y, o);
        }
        if ((threadIdx.x & 31) < K) s_dot[threadIdx.x >> 5][threadIdx.x & 31] = pd;
        __syncthreads();
        if (threadIdx.x < K) {
            double2 t = s_dot[0][threadIdx.x];
            for (int w = 1; w < 8; ++w) {
                t.x += s_dot[w][threadIdx.x].x;
                t.y += s_dot[w][threadIdx.x].y;
            }
            dot_partial[(int64_t)blockIdx.x * K + threadIdx.x] = t;
        }
    }
}

// Variant with one lane per column entity and all K right-hand sides in that lane (2K accumulators):
// no value is loaded twice.  Measured at C3: better at K = 2 (7.3 vs 7.6 ms), worse at K = 4 (12.8 vs
// 10.6 ms, registers), so it serves K = 2 only.
template <int K>
__global__ void __launch_bounds__(256) spmm_blocked2_lane_kernel(int64_t nb, const EntHdr *__restrict__ hdr,
                                                                 const int32_t *__restrict__ colstart,
                                                                 const double2 *__restrict__ vals,
                                                                 const double2 *__restrict__ X,
                                                                 const double2 *__restrict__ dscale,
                                                                 double2 *__restrict__ Y, int chunk) {
    const int lane = threadIdx.x % kBG;
    const int gpb = blockDim.x / kBG;  // in-order grid, see spmv_blocked2_kernel
    const unsigned gm = ((1u << kBG) - 1u) << ((threadIdx.x & 31) / kBG * kBG);
    const uint64_t stream = l2_policy_evict_first();
    for (int c = 0; c < chunk; ++c) {
        const int64_t i = (blockIdx.x * (int64_t)chunk + c) * gpb + threadIdx.x / kBG;
        if (i >= nb) break;
        const int4 *hp = reinterpret_cast<const int4 *>(hdr + i);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const int64_t valoff = ((int64_t)(unsigned)h0.x) | ((int64_t)h0.y << 32);
        const int L = (h1.x >> 16) & 0xffff;
        const int row = h1.z, cbase = h1.w;
        const int nc = L >> 1;
        const double2 *v0 = vals + valoff, *v1 = v0 + L;
        const int32_t *cs = colstart + cbase;
        double2 a0[K], a1[K];
#pragma unroll
        for (int r = 0; r < K; ++r) a0[r] = a1[r] = make_double2(0.0, 0.0);
        for (int j = lane; j < nc; j += kBG) {
            const int32_t c0 = ld_stream<1>(cs + j, stream);
            // one request per 32-byte sector (see ld256_stream); callers pass 32-byte aligned arrays
            const double2x2 p0 = ld256_stream(v0 + 2 * j, stream), p1 = ld256_stream(v1 + 2 * j, stream);
            const double2 p00 = p0.a, p01 = p0.b, p10 = p1.a, p11 = p1.b;
            const double2 *xp = X + (int64_t)c0 * K;
            double2 x0[K], x1[K];
            if (K == 2) {
                const double2x2 xa = ld256(xp), xb = ld256(xp + 2);
                x0[0] = xa.a, x0[K - 1] = xa.b, x1[0] = xb.a, x1[K - 1] = xb.b;
            } else {
#pragma unroll
                for (int r = 0; r < K; ++r) x0[r] = __ldg(xp + r), x1[r] = __ldg(xp + K + r);
            }
#pragma unroll
            for (int r = 0; r < K; ++r) {
                cfma2(a0[r], p00, x0[r]);
                cfma2(a1[r], p10, x0[r]);
                cfma2(a0[r], p01, x1[r]);
                cfma2(a1[r], p11, x1[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < K; ++r) {
#pragma unroll
            for (int o = kBG / 2; o > 0; o >>= 1) {
                a0[r].x += __shfl_xor_sync(gm, a0[r].x, o, kBG);
                a0[r].y += __shfl_xor_sync(gm, a0[r].y, o, kBG);
                a1[r].x += __shfl_xor_sync(gm, a1[r].x, o, kBG);
                a1[r].y += __shfl_xor_sync(gm, a1[r].y, o, kBG);
            }
        }
        // lanes 0..K-1 store row 0, lanes K..2K-1 row 1 (2K <= 8)
        double2 mine = a0[0];
#pragma unroll
        for (int r = 0; r < K; ++r) {
            if (lane == r) mine = a0[r];
            if (lane == K + r) mine = a1[r];
        }
        if (lane < 2 * K) {
            const int rr = lane / K;
            if (dscale) {
                const double2 d = __ldg(dscale + row + rr);
                mine = make_double2(d.x * mine.x - d.y * mine.y, d.x * mine.y + d.y * mine.x);
            }
            Y[(int64_t)row * K + lane] = mine;  // rows row, row+1 are contiguous: [row][K] then [row+1][K]
        }
    }
}

}  // namespace pg

using namespace pg;

// entities (processing order) whose column list reaches beyond the owned columns [0, n_own): flagged; the interior
// range is what lies between the last flagged entity of the first half and the first flagged one of the second half
__global__ void __launch_bounds__(256) halo_split_kernel(int64_t nb, const pg::EntHdr *__restrict__ hdr,
                                                         const int32_t *__restrict__ colstart, int32_t n_own,
                                                         int *__restrict__ lo_hi) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const pg::EntHdr h = hdr[i];
    const int nc = h.L >> 1;
    bool out = false;
    for (int j = 0; j < nc; ++j) out |= colstart[h.cbase + j] >= n_own;
    if (out) {
        if (i < nb / 2) atomicMax(lo_hi, (int)i + 1);
        else atomicMin(lo_hi + 1, (int)i);
    }
}

extern "C" int pg_plan_halo_split(const pg_plan *pl, const int32_t *colstart, int64_t n_own, int64_t *ent_begin_host,
                                  int64_t *ent_end_host, void *stream) {
    PG_REQUIRE(pl && colstart && ent_begin_host && ent_end_host, PG_EINVAL, "pg_plan_halo_split: null pointer");
    PG_REQUIRE(pl->p == 2, PG_EINVAL, "pg_plan_halo_split: p = 2 only (entity-blocked MatMult)");
    const int64_t nb = pl->b1 - pl->b0;
    *ent_begin_host = 0;
    *ent_end_host = nb;
    if (nb == 0) return PG_OK;
    PG_REQUIRE(nb < ((int64_t)1 << 31), PG_ERANGE, "pg_plan_halo_split: too many entities");
    int *d = nullptr;
    int h[2] = {0, (int)nb};
    PG_CUDA_OK(cudaMalloc((void **)&d, 8));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(d, h, 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        halo_split_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(nb, pl->hdr, colstart, (int32_t)n_own, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    PG_CUDA_OK(e);
    *ent_begin_host = h[0];
    *ent_end_host = std::max(h[0], h[1]);
    return PG_OK;
}

static int g_spmm_pf = -1;
static int g_spmv_chunk = 0;
extern "C" int pg_tune_spmv_chunk(int chunk) {  // consecutive 32-entity tiles per block of pg_spmv_blocked
    g_spmv_chunk = chunk;
    return PG_OK;
}
extern "C" int pg_tune_spmm_prefetch(int mode) {
    g_spmm_pf = mode;
    return PG_OK;
}

extern "C" int pg_spmv_blocked_range(const pg_plan *pl, int64_t ent_begin, int64_t ent_end, const int32_t *colstart,
                                     const double *vals, const double *x, const double *dscale, double *y, void *stream);

extern "C" int pg_spmv_blocked(const pg_plan *pl, const int32_t *colstart, const double *vals, const double *x,
                               const double *dscale, double *y, void *stream) {
    PG_REQUIRE(pl, PG_EINVAL, "pg_spmv_blocked: null pointer");
    return pg_spmv_blocked_range(pl, 0, pl->b1 - pl->b0, colstart, vals, x, dscale, y, stream);
}

// the rows of the entities [ent_begin, ent_end) of the plan's processing order only (interior / boundary split of
// the multi-GPU MatMult: the interior rows need no halo and run while it is in flight)
extern "C" int pg_spmv_blocked_range(const pg_plan *pl, int64_t ent_begin, int64_t ent_end, const int32_t *colstart,
                                     const double *vals, const double *x, const double *dscale, double *y, void *stream) {
    PG_REQUIRE(pl && vals && x && y, PG_EINVAL, "pg_spmv_blocked: null pointer");
    PG_REQUIRE(pl->p == 2, PG_EINVAL, "pg_spmv_blocked: only p = 2 has uniform 2x2 entity blocks (p = %d)", pl->p);
    PG_REQUIRE(ent_begin >= 0 && ent_begin <= ent_end && ent_end <= pl->b1 - pl->b0, PG_EINVAL,
               "pg_spmv_blocked_range: entity range [%lld, %lld)", (long long)ent_begin, (long long)ent_end);
    const int64_t nb = ent_end - ent_begin;
    if (nb == 0) return PG_OK;
    const EntHdr *hdr0 = pl->hdr + ent_begin;
    const int chunk = g_spmv_chunk > 0 ? g_spmv_chunk : 1;
    const int64_t tiles = (nb * kBG + 255) / 256, blocks = (tiles + chunk - 1) / chunk;
    const int32_t *cs = colstart ? colstart : pl->colstart;
    const double2 *v2 = reinterpret_cast<const double2 *>(vals), *x2 = reinterpret_cast<const double2 *>(x);
    const double2 *d2 = reinterpret_cast<const double2 *>(dscale);
    double2 *y2 = reinterpret_cast<double2 *>(y);
    cudaStream_t st = (cudaStream_t)stream;
    int mode = spmv_hint_mode();
    if (mode == 4 && (((uintptr_t)vals | (uintptr_t)x) & 31)) mode = 1;  // 256-bit loads need 32-byte aligned arrays
    switch (mode) {
        case 1: spmv_blocked2_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk); break;
        case 2: spmv_blocked2_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk); break;
        case 3: spmv_blocked2_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk); break;
        case 4: spmv_blocked2_v256_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk, nullptr); break;
        default: spmv_blocked2_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
    }
    PG_LAUNCH_OK();
    return PG_OK;
}

// MatMult fused with the unconjugated dot product x^T (A x) that COCG / COCR take right after it
// (rho = r~^T A r~, p^T A p): out[r] = sum_rows X[row, r] * Y[row, r], deterministic two-stage sum.
// Needs the 256-bit kernels (32-byte aligned arrays); returns PG_EINVAL otherwise so that the caller
// falls back to pg_spmv_blocked + pg_zbdotu.
extern "C" int64_t pg_spmv_dot_workspace_bytes(const pg_plan *pl, int k) {
    if (!pl) return 0;
    const int64_t nb = pl->b1 - pl->b0;
    const int64_t tiles = (nb * kBG + 255) / 256;
    return (tiles + 1 + kDotRed) * (int64_t)std::max(k, 1) * 16;
}

extern "C" int pg_spmm_blocked_dot(const pg_plan *pl, const int32_t *colstart, const double *vals, int k,
                                   const double *X, const double *dscale, double *Y, double *out, void *work,
                                   void *stream) {
    PG_REQUIRE(pl && vals && X && Y && out && work, PG_EINVAL, "pg_spmm_blocked_dot: null pointer");
    PG_REQUIRE(pl->p == 2, PG_EINVAL, "pg_spmm_blocked_dot: only p = 2 has uniform 2x2 entity blocks (p = %d)", pl->p);
    PG_REQUIRE(k == 1 || k == 4 || k == 8, PG_EINVAL, "pg_spmm_blocked_dot: k = %d (1, 4 or 8)", k);
    PG_REQUIRE(!(((uintptr_t)vals | (uintptr_t)X) & 31), PG_EINVAL, "pg_spmm_blocked_dot: arrays must be 32-byte aligned");
    const int64_t nb = pl->b1 - pl->b0;
    cudaStream_t st = (cudaStream_t)stream;
    if (nb == 0) {
        PG_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)k * 16, st));
        return PG_OK;
    }
    const int chunk = k == 1 ? (g_spmv_chunk > 0 ? g_spmv_chunk : 1) : 4;
    const int64_t tiles = (nb * kBG + 255) / 256, blocks = (tiles + chunk - 1) / chunk;
    const int32_t *cs = colstart ? colstart : pl->colstart;
    const double2 *v2 = reinterpret_cast<const double2 *>(vals), *x2 = reinterpret_cast<const double2 *>(X);
    const double2 *d2 = reinterpret_cast<const double2 *>(dscale);
    double2 *y2 = reinterpret_cast<double2 *>(Y), *part = static_cast<double2 *>(work);
    if (k == 1)
        spmv_blocked2_v256_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(nb, pl->hdr, cs, v2, x2, d2, y2, chunk, part);
    else if (k == 4)
        spmm_blocked2_pf_kernel<4, 4, true><<<(unsigned)blocks, 256, 0, st>>>(nb, pl->hdr, cs, v2, x2, d2, y2, chunk, part);
    else
        spmm_blocked2_pf_kernel<8, 4, true><<<(unsigned)blocks, 256, 0, st>>>(nb, pl->hdr, cs, v2, x2, d2, y2, chunk, part);
    PG_LAUNCH_OK();
    double2 *mid = part + (tiles + 1) * (int64_t)k;
    dot_stage2a_var<<<dim3(kDotRed, k), 256, 0, st>>>(part, blocks, k, mid);
    PG_LAUNCH_OK();
    dot_stage2b_var<<<k, 256, 0, st>>>(mid, reinterpret_cast<double2 *>(out));
    PG_LAUNCH_OK();
    return PG_OK;
}

extern "C" int pg_spmm_blocked_range(const pg_plan *pl, int64_t ent_begin, int64_t ent_end, const int32_t *colstart,
                                     const double *vals, int k, const double *X, const double *dscale, double *Y,
                                     void *stream);

extern "C" int pg_spmm_blocked(const pg_plan *pl, const int32_t *colstart, const double *vals, int k, const double *X,
                               const double *dscale, double *Y, void *stream) {
    PG_REQUIRE(pl, PG_EINVAL, "pg_spmm_blocked: null pointer");
    return pg_spmm_blocked_range(pl, 0, pl->b1 - pl->b0, colstart, vals, k, X, dscale, Y, stream);
}

extern "C" int pg_spmm_blocked_range(const pg_plan *pl, int64_t ent_begin, int64_t ent_end, const int32_t *colstart,
                                     const double *vals, int k, const double *X, const double *dscale, double *Y,
                                     void *stream) {
    PG_REQUIRE(pl && vals && X && Y, PG_EINVAL, "pg_spmm_blocked: null pointer");
    PG_REQUIRE(pl->p == 2, PG_EINVAL, "pg_spmm_blocked: only p = 2 has uniform 2x2 entity blocks (p = %d)", pl->p);
    PG_REQUIRE(k == 2 || k == 4 || k == 8, PG_EINVAL, "pg_spmm_blocked: k = %d (2, 4 or 8)", k);
    PG_REQUIRE(ent_begin >= 0 && ent_begin <= ent_end && ent_end <= pl->b1 - pl->b0, PG_EINVAL,
               "pg_spmm_blocked_range: entity range [%lld, %lld)", (long long)ent_begin, (long long)ent_end);
    const int64_t nb = ent_end - ent_begin;
    if (nb == 0) return PG_OK;
    const EntHdr *hdr0 = pl->hdr + ent_begin;
    static const int chunk_env = [] {
        const char *e = getenv("PG_SPMM_CHUNK");
        return e ? atoi(e) : 0;
    }();
    const int chunk = chunk_env > 0 ? chunk_env : 4;
    const int64_t tiles = (nb * kBG + 255) / 256, blocks = (tiles + chunk - 1) / chunk;
    const int32_t *cs = colstart ? colstart : pl->colstart;
    const double2 *v2 = reinterpret_cast<const double2 *>(vals), *x2 = reinterpret_cast<const double2 *>(X);
    const double2 *d2 = reinterpret_cast<const double2 *>(dscale);
    double2 *y2 = reinterpret_cast<double2 *>(Y);
    cudaStream_t st = (cudaStream_t)stream;
    // index-prefetch schedule (spmm_blocked2_pf_kernel) by default: measured at C3, k = 4: 10.96 -> 9.34 ms on the
    // whole matrix, 1.50 -> 1.23 ms on a 1/8 row block, bit-identical results (profiles/r2_spmv_l2_probe.jsonl);
    // PG_SPMM_PF=0 selects the plain loop
    static const int pf_env = [] {
        const char *e = getenv("PG_SPMM_PF");
        return e ? atoi(e) : 1;
    }();
    const int pf = g_spmm_pf >= 0 ? g_spmm_pf : pf_env;
    switch (k) {
        case 2: spmm_blocked2_lane_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk); break;
        case 4:
            if (pf && spmv_hint_mode() == 4 && !((uintptr_t)vals & 31)) spmm_blocked2_pf_kernel<4, 4><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            else if (pf && spmv_hint_mode() == 3) spmm_blocked2_pf_kernel<4, 3><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            else if (pf) spmm_blocked2_pf_kernel<4, 1><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            else spmm_blocked2_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            break;
        default:
            if (pf && spmv_hint_mode() == 4 && !((uintptr_t)vals & 31)) spmm_blocked2_pf_kernel<8, 4><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            else if (pf && spmv_hint_mode() == 3) spmm_blocked2_pf_kernel<8, 3><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            else if (pf) spmm_blocked2_pf_kernel<8, 1><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
            else spmm_blocked2_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(nb, hdr0, cs, v2, x2, d2, y2, chunk);
    }
    PG_LAUNCH_OK();
    return PG_OK;
}

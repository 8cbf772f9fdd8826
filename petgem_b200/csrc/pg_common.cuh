// Shared helpers for the petgem_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/petgem_b200.h"

namespace pg {

void set_error(const char *fmt, ...);

#define PG_CUDA_OK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            pg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? PG_ENOMEM : PG_ECUDA;                 \
        }                                                                                    \
    } while (0)

#define PG_LAUNCH_OK()                                                                       \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            pg::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return PG_ECUDA;                                                                 \
        }                                                                                    \
    } while (0)

#define PG_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            pg::set_error(__VA_ARGS__);        \
            return (code);                     \
        }                                      \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// compile-time facts about order P (hvfem.py:44-48)
template <int P>
struct Ord {
    static constexpr int p = P;
    static constexpr int ne = P;
    static constexpr int nf = P * (P - 1);
    static constexpr int nv = P * (P - 1) * (P - 2) / 2;
    static constexpr int n = P * (P + 2) * (P + 3) / 2;
    static constexpr int face_off = 6 * ne;           // local dof offset of the first face
    static constexpr int vol_off = 6 * ne + 4 * nf;   // local dof offset of the interior
    static constexpr int xface_off = 6 * ne;          // expanded offsets
    static constexpr int xvol_off = 6 * ne + 24 * nf;
    static constexpr int nexp = 6 * ne + 24 * nf + nv;
    static constexpr int nslots = 6 + (P >= 2 ? 4 : 0) + (P >= 3 ? 1 : 0);
};

__host__ __device__ inline int ndof_edge(int p) { return p; }
__host__ __device__ inline int ndof_face(int p) { return p * (p - 1); }
__host__ __device__ inline int ndof_volume(int p) { return p * (p - 1) * (p - 2) / 2; }
__host__ __device__ inline int ndof_element(int p) { return p * (p + 2) * (p + 3) / 2; }
__host__ __device__ inline int nslots_of(int p) { return 6 + (p >= 2 ? 4 : 0) + (p >= 3 ? 1 : 0); }

// slot (0..5 edges, 6..9 faces, 10 interior) and index inside the slot of local dof k
template <int P>
__device__ __forceinline__ void slot_of_local(int k, int &slot, int &d) {
    using O = Ord<P>;
    if (k < O::face_off) {
        slot = k / O::ne;
        d = k - slot * O::ne;
    } else if (O::nf > 0 && k < O::vol_off) {
        int kk = k - O::face_off;
        int f = kk / (O::nf > 0 ? O::nf : 1);
        slot = 6 + f;
        d = kk - f * O::nf;
    } else {
        slot = 10;
        d = k - O::vol_off;
    }
}

template <int P>
__device__ __forceinline__ int local_of_slot(int slot, int d) {
    using O = Ord<P>;
    if (slot < 6) return slot * O::ne + d;
    if (slot < 10) return O::face_off + (slot - 6) * O::nf + d;
    return O::vol_off + d;
}

// expanded (orientation-resolved) table index and edge-flip sign of (slot, d)
// code: bits 0..5 edge orientation, 3 bits per face from bit 6 (pg_element_geometry)
template <int P>
__device__ __forceinline__ int expanded_of_slot(int slot, int d, uint32_t code, double &sign) {
    using O = Ord<P>;
    sign = 1.0;
    if (slot < 6) {
        // OrientE (hvfem.py:791-822): swapping (s0,s1) negates the even-degree functions
        if (((code >> slot) & 1u) && ((d & 1) == 0)) sign = -1.0;
        return slot * O::ne + d;
    }
    if (slot < 10) {
        int f = slot - 6;
        int fo = (code >> (6 + 3 * f)) & 7u;
        return O::xface_off + (f * 6 + fo) * O::nf + d;
    }
    return O::xvol_off + d;
}

__device__ __forceinline__ int64_t global_entity(const int32_t *__restrict__ elemsE,
                                                 const int32_t *__restrict__ elemsF, int64_t nE, int64_t nF,
                                                 int64_t t, int slot) {
    if (slot < 6) return elemsE[t * 6 + slot];
    if (slot < 10) return nE + elemsF[t * 4 + (slot - 6)];
    return nE + nF + t;
}

__host__ __device__ inline int rows_of_entity(int64_t g, int64_t nE, int64_t nF, int p) {
    if (g < nE) return ndof_edge(p);
    if (g < nE + nF) return ndof_face(p);
    return ndof_volume(p);
}

// contraction of one table entry (SK[6], SM[6]) with the geometric factors
__device__ __forceinline__ void contract12(const double *__restrict__ tab, const double *g, double &k, double &m) {
    const double2 *t2 = reinterpret_cast<const double2 *>(tab);
    double2 a = __ldg(t2 + 0), b = __ldg(t2 + 1), c = __ldg(t2 + 2);
    double2 d = __ldg(t2 + 3), e = __ldg(t2 + 4), f = __ldg(t2 + 5);
    k = g[0] * a.x;
    k = fma(g[1], a.y, k);
    k = fma(g[2], b.x, k);
    k = fma(g[3], b.y, k);
    k = fma(g[4], c.x, k);
    k = fma(g[5], c.y, k);
    m = g[6] * d.x;
    m = fma(g[7], d.y, m);
    m = fma(g[8], e.x, m);
    m = fma(g[9], e.y, m);
    m = fma(g[10], f.x, m);
    m = fma(g[11], f.y, m);
}

// L2 eviction-priority hints: the matrix streams (values, column indices) are read once and must not
// push the x vector -- re-read ~40 times per entry -- out of the 126 MB L2 (ncu at C3 without hints:
// 33.3 GB of DRAM traffic against 28.9 GB algorithmic, i.e. x fetched ~9 times)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double2 ld_hint(const double2 *ptr, uint64_t policy) {
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(policy));
    return v;
}
__device__ __forceinline__ int32_t ld_hint(const int32_t *ptr, uint64_t policy) {
    int32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(policy));
    return v;
}

// the same, and the line is not allocated in L1: the matrix streams (20 MB per SM and pass) would otherwise
// push the x entries out of the 256 KB L1 that serves ~2/3 of the x gathers
__device__ __forceinline__ double2 ld_hint_na(const double2 *ptr, uint64_t policy) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(policy));
    return v;
}
__device__ __forceinline__ int32_t ld_hint_na(const int32_t *ptr, uint64_t policy) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(policy));
    return v;
}

// 256-bit loads (sm_100: LDG.E.256): a 2x1 block of complex values -- one 32-byte sector -- in ONE request.
// With two 16-byte loads per sector both requests miss L1 while the first is in flight, the L2 serves the
// first and evicts the evict-first line, and the second goes to DRAM again: ncu on the blocked SpMV at C3
// counted 1.41 G sector requests from L1 and 30.4 GB of DRAM reads for 24.5 GB of matrix.
struct __align__(32) double2x2 {
    double2 a, b;
};
__device__ __forceinline__ double2x2 ld256_stream(const double2 *ptr, uint64_t policy) {
    double2x2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
                 : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y) : "l"(ptr), "l"(policy));
    return v;
}
__device__ __forceinline__ double2x2 ld256(const double2 *ptr) {
    double2x2 v;
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y) : "l"(ptr));
    return v;
}

// HINT = 0: default policy for x, evict-first (ld.cs) for the streams; 1: explicit evict-first policy
// on the streams only; 2: evict-first on the streams and evict-last on x; 3: as 1, streams bypass L1
template <int HINT>
__device__ __forceinline__ double2 ld_stream(const double2 *p, uint64_t stream) {
    return HINT == 0 ? __ldcs(p) : HINT == 3 ? ld_hint_na(p, stream) : ld_hint(p, stream);
}
template <int HINT>
__device__ __forceinline__ int32_t ld_stream(const int32_t *p, uint64_t stream) {
    return HINT == 0 ? __ldcs(p) : HINT == 3 ? ld_hint_na(p, stream) : ld_hint(p, stream);
}
template <int HINT>
__device__ __forceinline__ double2 ld_keep(const double2 *p, uint64_t keep) {
    return HINT == 2 ? ld_hint(p, keep) : __ldg(p);
}
int spmv_hint_mode();  // PG_SPMV_HINTS environment variable (default 1), read once; pg_tune overrides
void set_spmv_hint_mode(int mode);

template <typename F>
int dispatch_order(int p, F &&f) {
    switch (p) {
        case 1: return f(Ord<1>());
        case 2: return f(Ord<2>());
        case 3: return f(Ord<3>());
        case 4: return f(Ord<4>());
        case 5: return f(Ord<5>());
        case 6: return f(Ord<6>());
        default:
            set_error("polynomial order %d outside 1..%d", p, PG_MAX_ORDER);
            return PG_EINVAL;
    }
}

}  // namespace pg

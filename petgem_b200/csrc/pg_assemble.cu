// Numeric assembly: fused element-matrix evaluation + deterministic row gather.
//
// Reference: the Python element loop solver.py:191-230 (computeElementalMatrices,
// Ae = K - i w mu M, MatSetValues(ADD_VALUES)) followed by MatAssembly
// (solver.py:233-235) and, optionally, MatZeroRowsColumns (solver.py:562).
//
// Design: owner-computes by ROW.  All rows of one mesh entity (edge / face /
// interior) share the same incident elements and the same column list, so a
// group of G lanes owns one entity: it zeroes a [rows x L] complex tile in shared
// memory, walks the entity's incident elements in ascending element index (the
// order PETSc would add them), evaluates exactly the local rows of Ae it needs
// with the 12-term geometric contraction, adds them at precomputed positions, and
// streams the finished rows to HBM with coalesced 16-byte stores.  Ae is never
// materialised, nothing is atomically updated, the result is bit-reproducible.
#include <cuda_pipeline.h>
#include <stdlib.h>

#include "pg_plan.cuh"

namespace pg {

struct AsmArgs {
    int64_t b0, b1;
    const int32_t *ent_order;
    const int64_t *row_base;
    const int32_t *inc_ptr;
    const IncRecord *rec;
    const int32_t *rowlen;
    const int32_t *selfpos;
    const int64_t *valoff;
    const uint8_t *bd_entity;  // null: no Dirichlet
    const EntHdr *hdr;         // owned entities in processing order
    const double *geo;
    const uint32_t *code;
    const double *table;
    const int32_t *itable;     // p >= 3: table in gather layout (int32 numerators; fp64 at p = 6)
    double inv_dk, inv_dm;     // 1/DK, 1/DM of the integer codes
    double mass_scale, diag;
    double2 *vals;
    int rc;         // rows per pass
    int bufstride;  // complex elements of shared tile per group
};

// integer denominators that make the reference tensors integral (basis.INTEGER_SCALES; asserted on the
// host and in tests/test_host.py); numerators fit 32 bits up to p = 5
__host__ __device__ inline double int_scale_k(int p) {
    return p == 3 ? 5040.0 : p == 4 ? 362880.0 : p == 5 ? 39916800.0 : 0.0;
}
__host__ __device__ inline double int_scale_m(int p) {
    return p == 3 ? 362880.0 : p == 4 ? 39916800.0 : p == 5 ? 6227020800.0 : 0.0;
}

// itable[Jx][q][Kx][4]: quad q = 0..2 of the 12 numerators of the pair (Jx, Kx); the lanes of a
// warp read consecutive column functions Kx, so each 16-byte quad load touches few 128-byte lines
__global__ void build_itable_kernel(int64_t nexp, int p, const double *__restrict__ table,
                                    int32_t *__restrict__ itable) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nexp * nexp * 12) return;
    const int c = (int)(i % 12);
    const int64_t pair = i / 12, jx = pair / nexp, kx = pair - jx * nexp;
    const double v = table[i] * (c < 6 ? int_scale_k(p) : int_scale_m(p));
    itable[(((jx * 3 + (c >> 2)) * nexp) + kx) * 4 + (c & 3)] = (int32_t)rint(v);
}

// p = 6 (numerators exceed 32 bits): the fp64 table in the same layout, dtable[Jx][q][Kx][2], q = 0..5
__global__ void build_dtable_kernel(int64_t nexp, const double *__restrict__ table, double *__restrict__ dtable) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nexp * nexp * 12) return;
    const int c = (int)(i % 12);
    const int64_t pair = i / 12, jx = pair / nexp, kx = pair - jx * nexp;
    dtable[(((jx * 6 + (c >> 1)) * nexp) + kx) * 2 + (c & 1)] = table[i];
}

// ---------------------------------------------------------------------------
// p >= 3: one warp per entity, table in L2
//
// * Lane l owns the local columns k = l, l + 32, ... (NC = ceil(n / 32) of them): their slot
//   and index inside the slot never change, and per incident element the expanded function
//   index, the orientation sign and the position in the entity's column list are worked out
//   once per column, not once per entry.
// * The table entry of a (row function, column function) pair is 12 exact integers (48 bytes,
//   p <= 5) widened with I2F, or 12 doubles (p = 6, numerators exceed 32 bits); all loads of a
//   batch of columns are issued before the first contraction.
// * The entity's rows are built in a shared tile of `bufstride` complex entries: as many rows
//   per pass as fit (all of them for most entities), records staged MC at a time.
// ---------------------------------------------------------------------------
template <int P>
struct Gen {
    static constexpr int n = Ord<P>::n;
    static constexpr int NC = (n + 31) / 32;        // columns per lane
    static constexpr int JB = NC < 3 ? NC : 3;      // columns per load batch
    static constexpr int MC = 8;                    // records staged per chunk
    static constexpr int GROUPS = 4;                // warps per block, one entity each
};

template <bool ITAB>
struct TabEntry;
template <>
struct TabEntry<true> {  // exact integer numerators
    int4 w[3];
    __device__ __forceinline__ void load(const AsmArgs &a, int64_t jx, int kx, int nexp) {
        const int4 *q = reinterpret_cast<const int4 *>(a.itable) + jx * 3 * nexp + kx;
        w[0] = __ldg(q), w[1] = __ldg(q + nexp), w[2] = __ldg(q + 2 * nexp);
    }
    __device__ __forceinline__ void contract(const double (&g)[12], double &k, double &m) const {
        k = g[0] * (double)w[0].x;
        k = fma(g[1], (double)w[0].y, k);
        k = fma(g[2], (double)w[0].z, k);
        k = fma(g[3], (double)w[0].w, k);
        k = fma(g[4], (double)w[1].x, k);
        k = fma(g[5], (double)w[1].y, k);
        m = g[6] * (double)w[1].z;
        m = fma(g[7], (double)w[1].w, m);
        m = fma(g[8], (double)w[2].x, m);
        m = fma(g[9], (double)w[2].y, m);
        m = fma(g[10], (double)w[2].z, m);
        m = fma(g[11], (double)w[2].w, m);
    }
};
template <>
struct TabEntry<false> {  // fp64 table
    double2 w[6];
    __device__ __forceinline__ void load(const AsmArgs &a, int64_t jx, int kx, int nexp) {
        const double2 *q = reinterpret_cast<const double2 *>(a.itable) + jx * 6 * nexp + kx;
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = __ldg(q + i * nexp);
    }
    __device__ __forceinline__ void contract(const double (&g)[12], double &k, double &m) const {
        k = g[0] * w[0].x;
        k = fma(g[1], w[0].y, k);
        k = fma(g[2], w[1].x, k);
        k = fma(g[3], w[1].y, k);
        k = fma(g[4], w[2].x, k);
        k = fma(g[5], w[2].y, k);
        m = g[6] * w[3].x;
        m = fma(g[7], w[3].y, m);
        m = fma(g[8], w[4].x, m);
        m = fma(g[9], w[4].y, m);
        m = fma(g[10], w[5].x, m);
        m = fma(g[11], w[5].y, m);
    }
};

// contraction of one loaded table entry and its add (or first-touch store) into the tile
template <bool ITAB>
__device__ __forceinline__ void add_entry(const TabEntry<ITAB> &e, const double (&gf)[12], double2 *dst, double sK,
                                          double sM, bool live, bool neg, bool first, bool zero) {
    double kk, mm;
    e.contract(gf, kk, mm);
    if (live) {
        double2 cur = make_double2(0.0, 0.0);
        if (!first) cur = *dst;
        cur.x = fma(neg ? -sK : sK, kk, cur.x);
        cur.y = fma(neg ? -sM : sM, mm, cur.y);
        *dst = cur;
    } else if (zero) {  // Dirichlet column touched first: explicit zero
        *dst = make_double2(0.0, 0.0);
    }
}

template <int P, bool ITAB>
__device__ __forceinline__ void assemble_body(const AsmArgs &a) {
    using O = Ord<P>;
    using S = Gen<P>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IncRecord *s_rec = reinterpret_cast<IncRecord *>(smem_raw);
    double2 *s_buf = reinterpret_cast<double2 *>(s_rec + S::GROUPS * S::MC);

    const int grp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    double2 *buf = s_buf + (size_t)grp * a.bufstride;
    IncRecord *rec = s_rec + grp * S::MC;

    int cks[S::NC], ckd[S::NC];  // slot and index inside the slot of the lane's columns (-1: none)
#pragma unroll
    for (int j = 0; j < S::NC; ++j) {
        const int k = lane + 32 * j;
        slot_of_local<P>(k < S::n ? k : 0, cks[j], ckd[j]);
        if (k >= S::n) cks[j] = -1;
    }
    const double scK = ITAB ? a.inv_dk : 1.0;
    const double scM = ITAB ? a.mass_scale * a.inv_dm : a.mass_scale;

    const int64_t nb = a.b1 - a.b0;
    for (int64_t i = blockIdx.x * (int64_t)S::GROUPS + grp; i < nb; i += (int64_t)gridDim.x * S::GROUPS) {
        EntHdr h;
        {
            const int4 *hp = reinterpret_cast<const int4 *>(a.hdr + i);
            reinterpret_cast<int4 *>(&h)[0] = __ldg(hp);
            reinterpret_cast<int4 *>(&h)[1] = __ldg(hp + 1);
        }
        const int m = h.m, L = h.L, r = h.rows;
        double2 *out = a.vals + h.valoff;

        if (a.bd_entity && h.bd) {
            // Dirichlet entity: identity rows (MatZeroRowsColumns puts `diag` on the diagonal)
            const int sp = h.selfpos;
            for (int it = lane; it < r * L; it += 32) {
                const int d = it / L, c = it - d * L;
                __stcs(out + it, make_double2(c == sp + d ? a.diag : 0.0, 0.0));
            }
            continue;
        }

        const int rc = min(r, a.bufstride / L);  // rows per pass (>= 1: bufstride >= max row length)
        for (int d0 = 0; d0 < r; d0 += rc) {
            const int rcur = min(rc, r - d0);
            // no zero fill: the element that first touches a column stores (firstmask), later ones add
            for (int c0 = 0; c0 < m; c0 += S::MC) {
                const int mc = min((int)S::MC, m - c0);
                __syncwarp();  // previous chunk consumed
                if (lane < 2 * mc)
                    reinterpret_cast<int4 *>(rec)[lane] =
                        __ldg(reinterpret_cast<const int4 *>(a.rec + h.inc0 + c0) + lane);
                __syncwarp();
                for (int ia = 0; ia < mc; ++ia) {
                    const IncRecord *rp = rec + ia;
                    const int64_t t = rp->elem;
                    const int rslot = rp->slot;
                    const unsigned bdm = a.bd_entity ? rp->bdmask : 0u;
                    const unsigned fm = rp->firstmask;
                    double gf[12];
                    {
                        const double2 *gp = reinterpret_cast<const double2 *>(a.geo + t * 12);
#pragma unroll
                        for (int q = 0; q < 6; ++q) {
                            const double2 v = __ldg(gp + q);
                            gf[2 * q] = v.x;
                            gf[2 * q + 1] = v.y;
                        }
                    }
                    const uint32_t cd = __ldg(a.code + t);
                    // the lane's columns in this element: expanded function, sign, tile position
                    int cxi[S::NC], cpos[S::NC];
                    bool cneg[S::NC], cfirst[S::NC], czero[S::NC];
#pragma unroll
                    for (int j = 0; j < S::NC; ++j) {
                        cxi[j] = -1, cpos[j] = 0, cneg[j] = cfirst[j] = czero[j] = false;
                        if (cks[j] >= 0) {
                            const bool bdc = (bdm >> cks[j]) & 1u;  // Dirichlet column: zero
                            cfirst[j] = (fm >> cks[j]) & 1u;
                            cpos[j] = rp->slotpos[cks[j]] + ckd[j];
                            czero[j] = bdc && cfirst[j];
                            if (!bdc) {
                                double s2;
                                cxi[j] = expanded_of_slot<P>(cks[j], ckd[j], cd, s2);
                                cneg[j] = s2 < 0.0;
                            }
                        }
                    }
                    for (int dd = 0; dd < rcur; ++dd) {
                        double s1;
                        const int Jx = expanded_of_slot<P>(rslot, d0 + dd, cd, s1);
                        const double sK = s1 * scK, sM = s1 * scM;
                        double2 *brow = buf + dd * L;
#pragma unroll
                        for (int j0 = 0; j0 < S::NC; j0 += S::JB) {
                            TabEntry<ITAB> e[S::JB];  // all loads of the batch in flight before the first use
#pragma unroll
                            for (int jj = 0; jj < S::JB; ++jj)
                                if (j0 + jj < S::NC) e[jj].load(a, Jx, max(cxi[j0 + jj], 0), O::nexp);
#pragma unroll
                            for (int jj = 0; jj < S::JB; ++jj)
                                if (j0 + jj < S::NC)
                                    add_entry(e[jj], gf, brow + cpos[j0 + jj], sK, sM, cxi[j0 + jj] >= 0,
                                              cneg[j0 + jj], cfirst[j0 + jj], czero[j0 + jj]);
                        }
                    }
                    __syncwarp();  // the next element may touch the same positions from other lanes
                }
            }
            double2 *o2 = out + (int64_t)d0 * L;
            for (int it = lane; it < rcur * L; it += 32) __stcs(o2 + it, buf[it]);
            __syncwarp();
        }
    }
}

// p = 5 would take 162 registers: capped at 128 (4 blocks per SM); the others fit on their own
template <int P, bool ITAB>
__global__ void __launch_bounds__(128) assemble_kernel(const AsmArgs a) {
    assemble_body<P, ITAB>(a);
}
template <int P, bool ITAB>
__global__ void __launch_bounds__(128, 4) assemble_kernel_capped(const AsmArgs a) {
    assemble_body<P, ITAB>(a);
}

// ---------------------------------------------------------------------------
// p = 1, 2: shared-memory integer-table kernel
//
// * For p <= 2 every row entity has the same number of rows R (1 or 2) and the
//   orientation-resolved function set collapses: at p = 2 the two functions of a face are,
//   for each of the six face orientations, two of the three functions
//   psi0 = l_c w_ab, psi1 = l_a w_bc, psi2 = l_b w_ca up to sign (w_xy = l_x grad l_y -
//   l_y grad l_x): 12 edge + 4 x 3 face = 24 functions (kFaceLut2).
// * The reference tensors are integrals of integer-coefficient polynomials over the master
//   tetrahedron, i.e. exact rationals: SK * DK and SM * DM are small integers (p=1: DK=6,
//   DM=120; p=2: DK=120, DM=5040, |n| <= 252; asserted on the host in basis.py and in
//   tests/test_host.py).  They are kept in shared memory as biased 16-bit integers, 32 bytes
//   per (row function, column) entry instead of 96, and widened to fp64 exactly with the
//   2^52 trick (one DADD); the 1/D and -omega*mu factors ride on the final sign multiply.
// * Columns are addressed by their LOCAL index k (orientation independent) and the face
//   variant selects a plane, so the 8 lanes of a group always hit 8 different 16-byte bank
//   groups: conflict-free table reads.
// * Eight lanes own one entity: R*n = 40 (6 at p = 1) entries per incident element, 5 per
//   lane over 3 distinct columns.  The element that first touches a column stores, later
//   ones read-modify-write (firstmask), so the tile is never zero-filled.
// ---------------------------------------------------------------------------
template <int P>
struct Small {
    static constexpr int R = P;                      // rows per entity (edges p, faces p(p-1): 1 / 2)
    static constexpr int n = Ord<P>::n;              // 6 / 20
    static constexpr int NX = (P == 1) ? 6 : 24;     // reduced function set
    static constexpr int NENT = (P == 1) ? 36 : 960; // table entries (plane 0: 24x24, planes 1,2: 24x8)
    static constexpr int G = 8;
    static constexpr int NCOL = (P == 1) ? 1 : 3;    // distinct columns per lane
    static constexpr int IPL = (P == 1) ? 1 : 5;     // items per lane
    static constexpr int MC = 8;                     // records staged per chunk
    static constexpr double DK = (P == 1) ? 6.0 : 120.0;
    static constexpr double DM = (P == 1) ? 120.0 : 5040.0;
    // 16-bit code = 0x3000 | ((n + BIAS) << SHIFT): fraction (n + BIAS) 2^(SHIFT-12), bias at 1/2
    static constexpr int BIAS = (P == 1) ? 16 : 256;   // |n| <= 8 (p=1), <= 252 (p=2)
    static constexpr int SHIFT = (P == 1) ? 7 : 3;
    static constexpr double UNIT = (P == 1) ? 7.105427357601002e-15 : 1.1368683772161603e-13;  // 2^-47 / 2^-43
};

// (orientation o, family d) -> base function 0..2 and sign bit, 3 bits each
constexpr uint64_t kFaceLut2 =
    (0ull << 0) | (1ull << 3) | (1ull << 6) | (2ull << 9) | (2ull << 12) | (0ull << 15) | (6ull << 18) |
    (5ull << 21) | (4ull << 24) | (6ull << 27) | (5ull << 30) | (4ull << 33);

// general expanded index of reduced function jr (used once, to fill the shared table)
template <int P>
__device__ __forceinline__ int expanded_of_reduced(int jr) {
    if (P == 1 || jr < 12) return jr;
    const int f = (jr - 12) / 3, b = (jr - 12) % 3;
    const int o = (b == 2) ? 1 : 0, fam = (b == 0) ? 0 : 1;
    return 12 + (f * 6 + o) * 2 + fam;
}

// reduced index of a ROW function (slot, d) under orientation code; neg = sign flip
template <int P>
__device__ __forceinline__ int reduced_row(int slot, int d, uint32_t code, unsigned &neg) {
    if (P == 1 || slot < 6) {
        neg = (((code >> slot) & 1u) && ((d & 1) == 0)) ? 1u : 0u;
        return slot * P + d;
    }
    const int f = slot - 6;
    const unsigned o = (code >> (6 + 3 * f)) & 7u;
    const unsigned e = (unsigned)(kFaceLut2 >> (6 * o + 3 * d)) & 7u;
    neg = e >> 2;
    return 12 + f * 3 + (int)(e & 3u);
}

struct ColDesc {     // one local column k handled by this lane: fixed for the whole kernel
    int k, slot, d;
    int recoff;      // byte offset of slotpos[slot] in an IncRecord
    unsigned ebit;   // edge column: orientation bit that flips the sign (0 if none)
    int fshift;      // face column: shift of its 3-bit orientation code, -1 for edge columns
};

// Table codes.  An integer numerator n in [-256, 255] is stored as the 16-bit code
// 0x3000 | ((n + 256) << 3); dropped into bits 8..23 of the high word 0x43000000 (one PRMT) it
// forms, with a zero low word, the double X = 2^52 (1 + (n+256)/512) = Xb + n 2^43, Xb = 1.5 2^52.
// The contraction sum_c g_c X_c is taken as is and the constant part sum_c g_c Xb (same operation
// order, so all-zero entries cancel bit-exactly) is subtracted once; 2^-43 rides on the final scale.
// Rounding: <= 3 ulp of 2^52 sum|g| against entries of size 2^43 |n| sum|g| -> ~2e-15 norm-relative.
__device__ __forceinline__ double code_lo(unsigned w, unsigned hi_const) {
    return __hiloint2double((int)__byte_perm(w, hi_const, 0x7104), 0);
}
__device__ __forceinline__ double code_hi(unsigned w, unsigned hi_const) {
    return __hiloint2double((int)__byte_perm(w, hi_const, 0x7324), 0);
}

template <int P>
struct SmallCtx {  // per-lane state of assemble_small_kernel that process_element needs
    const uint4 *tabA;
    const uint2 *tabB;
    double2 *buf;
    ColDesc col[Small<P>::NCOL];
    double invK, invM;
    unsigned hi_const;
    int row4, L;
    bool lane_valid, use_bd;
};

// one incident element: the lane's (up to) 5 entries of the local rows, added into the tile
template <int P>
__device__ __forceinline__ void process_element(const SmallCtx<P> &cx, const double2 (&g)[6], uint32_t cd,
                                                const IncRecord *recp) {
    using S = Small<P>;
    const unsigned char *rbytes = reinterpret_cast<const unsigned char *>(recp);
    const unsigned w24 = *reinterpret_cast<const unsigned *>(rbytes + 24);  // slotpos[10] | slot<<16
    const unsigned w28 = *reinterpret_cast<const unsigned *>(rbytes + 28);  // bdmask | firstmask<<16
    const int rslot = (w24 >> 16) & 0xff;
    const unsigned bdm = cx.use_bd ? (w28 & 0xffffu) : 0u;
    const unsigned fm = w28 >> 16;

    int jrow[S::R];
    unsigned jneg[S::R];
#pragma unroll
    for (int d = 0; d < S::R; ++d) jrow[d] = reduced_row<P>(rslot, d, cd, jneg[d]);

    // constant part of the contraction (all codes at the bias), same operation order as below
    const double xb = __hiloint2double(0x43380000, 0);
    double cK = g[0].x * xb;
    cK = fma(g[0].y, xb, cK);
    cK = fma(g[1].x, xb, cK);
    cK = fma(g[1].y, xb, cK);
    cK = fma(g[2].x, xb, cK);
    cK = fma(g[2].y, xb, cK);
    double cM = g[3].x * xb;
    cM = fma(g[3].y, xb, cM);
    cM = fma(g[4].x, xb, cM);
    cM = fma(g[4].y, xb, cM);
    cM = fma(g[5].x, xb, cM);
    cM = fma(g[5].y, xb, cM);

    int cent[S::NCOL], cmul[S::NCOL], cpos[S::NCOL];
    unsigned cneg[S::NCOL], cfirst[S::NCOL], cbd[S::NCOL];
#pragma unroll
    for (int j = 0; j < S::NCOL; ++j) {
        const ColDesc &cj = cx.col[j];
        cneg[j] = (cd & cj.ebit) ? 1u : 0u;
        cent[j] = cj.k;
        cmul[j] = (P == 1) ? 6 : 24;
        if (P >= 2 && cj.fshift >= 0) {
            const unsigned o = (cd >> cj.fshift) & 7u;
            const unsigned e = (unsigned)(kFaceLut2 >> (6 * o + 3 * cj.d)) & 7u;
            const int b = (int)(e & 3u);
            cneg[j] = e >> 2;
            if (b) {  // face variant planes: 8 wide, column (k - 8) & 7 keeps the bank phase k mod 8
                cent[j] = 576 + (b - 1) * 192 + ((cj.k - 8) & 7);
                cmul[j] = 8;
            }
        }
        cpos[j] = *reinterpret_cast<const uint16_t *>(rbytes + cj.recoff) + cj.d;
        cfirst[j] = (fm >> cj.slot) & 1u;
        cbd[j] = (bdm >> cj.slot) & 1u;
    }
#pragma unroll
    for (int q = 0; q < S::IPL; ++q) {
        const int j = (P == 1) ? 0 : (q < 2 ? 0 : (q < 4 ? 1 : 2));
        const int row = (P == 1) ? 0 : (q == 4 ? cx.row4 : (q & 1));
        const int jr = (S::R == 1) ? jrow[0] : (row ? jrow[S::R - 1] : jrow[0]);
        const unsigned neg = cneg[j] ^ ((S::R == 1) ? jneg[0] : (row ? jneg[S::R - 1] : jneg[0]));
        const int ent = cent[j] + jr * cmul[j];
        const uint4 w0 = cx.tabA[ent];
        const uint2 w1 = cx.tabB[ent];
        const unsigned hc = cx.hi_const;
        double kk = g[0].x * code_lo(w0.x, hc);
        kk = fma(g[0].y, code_hi(w0.x, hc), kk);
        kk = fma(g[1].x, code_lo(w0.y, hc), kk);
        kk = fma(g[1].y, code_hi(w0.y, hc), kk);
        kk = fma(g[2].x, code_lo(w0.z, hc), kk);
        kk = fma(g[2].y, code_hi(w0.z, hc), kk);
        double mm = g[3].x * code_lo(w0.w, hc);
        mm = fma(g[3].y, code_hi(w0.w, hc), mm);
        mm = fma(g[4].x, code_lo(w1.x, hc), mm);
        mm = fma(g[4].y, code_hi(w1.x, hc), mm);
        mm = fma(g[5].x, code_lo(w1.y, hc), mm);
        mm = fma(g[5].y, code_hi(w1.y, hc), mm);
        kk -= cK;
        mm -= cM;
        // +-2^-43/D (orientation signs), 0 for a Dirichlet column
        double sK = neg ? -cx.invK : cx.invK, sM = neg ? -cx.invM : cx.invM;
        if (cbd[j]) {
            sK = 0.0;
            sM = 0.0;
        }
        if (P == 1 && !cx.lane_valid) continue;
        double2 *dst = cx.buf + cpos[j] + row * cx.L;
        double2 cur = make_double2(0.0, 0.0);
        if (!cfirst[j]) cur = *dst;
        cur.x = fma(sK, kk, cur.x);
        cur.y = fma(sM, mm, cur.y);
        *dst = cur;
    }
}

template <int P>
__device__ __forceinline__ void load_geo(const AsmArgs &a, int64_t t, double2 (&g)[6], uint32_t &cd) {
    const double2 *gp = reinterpret_cast<const double2 *>(a.geo + t * 12);
#pragma unroll
    for (int q = 0; q < 6; ++q) g[q] = __ldg(gp + q);
    cd = __ldg(a.code + t);
}

// BULK variant (PG_ASM_BULK=1): finished tiles leave shared memory through the bulk-copy engine
// (cp.async.bulk shared -> global, SASS UBLKCP) instead of an LDS + ST.CS loop through the load/store pipe
// (86 % of l1tex data-pipe wavefronts, ~11 % of them this copy): the rows of an entity are contiguous in vals
// and in the tile, so one lane issues one copy of R*L*16 bytes per entity.  Measured at C3: LSU pipe 86 -> 74 %,
// kernel 8.27 -> 8.45 ms (the fence + wait + syncs cost more than the loop; the kernel is issue/latency bound
// at 4 warps per scheduler), so it is NOT the default.  Evidence: profiles/r2_assemble_small_bulk_ab.json.
__device__ __forceinline__ void bulk_store_tile(double2 *gdst, const double2 *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                 "cp.async.bulk.commit_group;"
                 :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int P, bool BULK>
__global__ void __launch_bounds__(512, 1) assemble_small_kernel(const AsmArgs a) {
    using S = Small<P>;
    constexpr int G = S::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: integer table, SoA: A [NENT] x 16 B (K0..K5, M0, M1) | B [NENT] x 8 B (M2..M5) |
    //         staged records [groups][2 (MC+1) + 1] (odd number of 32-byte records per group: the four
    //         groups of a warp read their records from disjoint banks) |
    //         tiles [groups][R*L] complex
    uint4 *s_tabA = reinterpret_cast<uint4 *>(smem_raw);
    uint2 *s_tabB = reinterpret_cast<uint2 *>(s_tabA + S::NENT);
    const int groups = blockDim.x / G;
    IncRecord *s_rec = reinterpret_cast<IncRecord *>(s_tabB + S::NENT);
    EntHdr *s_hdr = reinterpret_cast<EntHdr *>(s_rec + groups * (2 * (S::MC + 1) + 1));  // 3 headers per group
    double2 *s_buf = reinterpret_cast<double2 *>(s_hdr + groups * 3);

    {   // fill the integer table from the general fp64 table (L2 resident)
        unsigned short *tA = reinterpret_cast<unsigned short *>(s_tabA);
        unsigned short *tB = reinterpret_cast<unsigned short *>(s_tabB);
        for (int i = threadIdx.x; i < S::NENT * 12; i += blockDim.x) {
            const int e = i / 12, c = i - e * 12;
            int jr, kr;
            bool valid = true;
            if (P == 1) {
                jr = e / 6;
                kr = e - jr * 6;
            } else if (e < 576) {
                jr = e / 24;
                const int k = e - jr * 24;
                valid = valid && k < 20;
                kr = (k < 12) ? k : 12 + 3 * ((k - 12) >> 1);
            } else {
                const int e2 = e - 576, b = 1 + e2 / 192, r2 = e2 % 192;
                jr = r2 >> 3;
                const int k = 12 + (((r2 & 7) + 4) & 7);  // inverse of (k - 8) & 7 on k = 12..19
                kr = 12 + 3 * ((k - 12) >> 1) + b;
            }
            int nint = 0;
            if (valid) {
                const double v = __ldg(a.table + ((int64_t)expanded_of_reduced<P>(jr) * Ord<P>::nexp +
                                                  expanded_of_reduced<P>(kr)) * 12 + c);
                nint = (int)rint(v * (c < 6 ? S::DK : S::DM));
            }
            const unsigned short u = (unsigned short)(0x3000 | ((nint + S::BIAS) << S::SHIFT));
            if (c < 8) tA[e * 8 + c] = u; else tB[e * 4 + (c - 8)] = u;
        }
    }
    __syncthreads();

    const int grp = threadIdx.x / G;
    const int lane = threadIdx.x % G;
    const unsigned mask = ((1u << G) - 1u) << ((threadIdx.x & 31) / G * G);
    IncRecord *rec = s_rec + grp * (2 * (S::MC + 1) + 1);  // two staging buffers (double buffered)

    SmallCtx<P> cx;
    cx.tabA = s_tabA;
    cx.tabB = s_tabB;
    cx.buf = s_buf + (size_t)grp * a.bufstride;
#pragma unroll
    for (int j = 0; j < S::NCOL; ++j) {
        int k = (j == 0) ? lane : (j == 1) ? lane + 8 : 16 + (lane & 3);
        if (k >= S::n) k = S::n - 1;  // p=1: lanes 6,7 are idle (lane_valid)
        cx.col[j].k = k;
        slot_of_local<P>(k, cx.col[j].slot, cx.col[j].d);
        cx.col[j].recoff = 4 + 2 * cx.col[j].slot;
        const bool face = cx.col[j].slot >= 6;
        cx.col[j].ebit = (!face && (cx.col[j].d & 1) == 0) ? (1u << cx.col[j].slot) : 0u;
        cx.col[j].fshift = face ? 6 + 3 * (cx.col[j].slot - 6) : -1;
    }
    cx.lane_valid = lane < S::n;
    cx.row4 = lane >> 2;
    cx.invK = (1.0 / S::DK) * S::UNIT;           // 2^-(52 - 12 + SHIFT) / DK
    cx.invM = (a.mass_scale / S::DM) * S::UNIT;  // ... * (-omega mu) / DM
    cx.hi_const = 0x43000000u;
    cx.use_bd = a.bd_entity != nullptr;

    // Software pipeline over the group's entities e_0, e_1, ... (stride ngroups), all prefetches
    // by cp.async straight into shared memory (no registers held across the element loop):
    //   while e_k is processed, the header of e_{k+2} and the records of e_{k+1} are in flight,
    //   and the geometric factors of e_{k+1}'s first element are fetched before e_k's rows are
    //   streamed out.
    const int64_t nb = a.b1 - a.b0;
    const int64_t ngroups = (int64_t)gridDim.x * groups;
    int64_t i = blockIdx.x * (int64_t)groups + grp;
    EntHdr *hring = s_hdr + grp * 3;  // headers of e_k, e_{k+1}, e_{k+2} at slots k%3, (k+1)%3, (k+2)%3
    int hs = 0;                       // ring slot of the current entity
    int cur = 0;                      // staging buffer of the current entity's records
    double2 gA[6], gB[6];
    uint32_t cA = 0, cB = 0;

    // prologue: headers of e_0, e_1; records and first geometry of e_0
    if (lane < 2) {
        if (i < nb) __pipeline_memcpy_async(reinterpret_cast<char *>(hring) + 16 * lane,
                                            reinterpret_cast<const char *>(a.hdr + i) + 16 * lane, 16);
        if (i + ngroups < nb)
            __pipeline_memcpy_async(reinterpret_cast<char *>(hring + 1) + 16 * lane,
                                    reinterpret_cast<const char *>(a.hdr + i + ngroups) + 16 * lane, 16);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncwarp(mask);
    if (i < nb && !(cx.use_bd && hring[0].bd)) {
        if (lane < min((int)S::MC, (int)hring[0].m)) {
            const char *src = reinterpret_cast<const char *>(a.rec + hring[0].inc0 + lane);
            __pipeline_memcpy_async(rec + lane, src, 16);
            __pipeline_memcpy_async(reinterpret_cast<char *>(rec + lane) + 16, src + 16, 16);
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncwarp(mask);
        load_geo<P>(a, rec[0].elem, gA, cA);
    }
    for (; i < nb; i += ngroups) {
        const EntHdr *hc = hring + hs;
        const int hs1 = (hs == 2) ? 0 : hs + 1, hs2 = (hs1 == 2) ? 0 : hs1 + 1;
        const EntHdr *hn = hring + hs1;
        const int m = hc->m, L = hc->L;
        const bool bd = cx.use_bd && hc->bd;
        double2 *out = a.vals + hc->valoff;
        cx.L = L;
        // prefetch: records of e_{k+1} (its header arrived during e_{k-1}), header of e_{k+2}
        const bool next_ok = (i + ngroups < nb) && !(cx.use_bd && hn->bd);
        IncRecord *rnext = rec + (cur ^ 1) * (S::MC + 1);
        if (next_ok && lane < min((int)S::MC, (int)hn->m)) {
            const char *src = reinterpret_cast<const char *>(a.rec + hn->inc0 + lane);
            __pipeline_memcpy_async(rnext + lane, src, 16);
            __pipeline_memcpy_async(reinterpret_cast<char *>(rnext + lane) + 16, src + 16, 16);
        }
        if (lane < 2 && i + 2 * ngroups < nb)
            __pipeline_memcpy_async(reinterpret_cast<char *>(hring + hs2) + 16 * lane,
                                    reinterpret_cast<const char *>(a.hdr + i + 2 * ngroups) + 16 * lane, 16);
        __pipeline_commit();
        if (BULK) {  // the bulk copy of the previous entity has read the tile before it is written again
            if (lane == 0) bulk_store_wait_read();
            __syncwarp(mask);
        }

        if (bd) {
            const int sp = hc->selfpos;
            for (int it = lane; it < S::R * L; it += G) {
                const int d = it / L, c = it - d * L;
                __stcs(out + it, make_double2(c == sp + d ? a.diag : 0.0, 0.0));
            }
        } else {
            IncRecord *rc = rec + cur * (S::MC + 1);
            for (int c0 = 0; c0 < m; c0 += S::MC) {
                const int mc = min(S::MC, m - c0);
                if (c0 > 0) {  // rare: more than MC incident elements, fetch the next chunk synchronously
                    __syncwarp(mask);
                    if (lane < mc) {
                        const int4 *src = reinterpret_cast<const int4 *>(a.rec + hc->inc0 + c0 + lane);
                        int4 *dst = reinterpret_cast<int4 *>(rc + lane);
                        dst[0] = __ldg(src);
                        dst[1] = __ldg(src + 1);
                    }
                    __syncwarp(mask);
                    load_geo<P>(a, rc[0].elem, gA, cA);
                }
                for (int ia = 0; ia < mc; ia += 2) {
                    const bool has_b = ia + 1 < mc;
                    if (has_b) load_geo<P>(a, rc[ia + 1].elem, gB, cB);
                    process_element<P>(cx, gA, cA, rc + ia);
                    __syncwarp(mask);  // the next element may touch the same positions from other lanes
                    if (has_b) {
                        if (ia + 2 < mc) load_geo<P>(a, rc[ia + 2].elem, gA, cA);
                        process_element<P>(cx, gB, cB, rc + ia + 1);
                        __syncwarp(mask);
                    }
                }
            }
        }
        // everything prefetched at the top has landed long ago
        __pipeline_wait_prior(0);
        __syncwarp(mask);
        if (next_ok) load_geo<P>(a, rnext[0].elem, gA, cA);
        if (!bd) {
            if (BULK) {
                fence_proxy_async_smem();  // this lane's tile writes -> visible to the async proxy
                __syncwarp(mask);
                if (lane == 0) bulk_store_tile(out, cx.buf, (unsigned)(S::R * L * 16));
            } else {
                for (int it = lane; it < S::R * L; it += G) __stcs(out + it, cx.buf[it]);
                __syncwarp(mask);  // tile reusable
            }
        }
        hs = hs1;
        cur ^= 1;
    }
    if (BULK && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // copies complete before exit
}

template <int P>
static int launch_assemble_small(const pg_plan *pl, AsmArgs a, cudaStream_t st) {
    using S = Small<P>;
    const int L = pl->max_rowlen;
    const size_t table_bytes = (size_t)S::NENT * 24;
    const size_t per_group = (2 * (S::MC + 1) + 1) * sizeof(IncRecord) + 3 * sizeof(EntHdr) + (size_t)S::R * L * 16;
    const size_t budget = 227 * 1024 - 1024;
    PG_REQUIRE(table_bytes + 4 * per_group <= budget, PG_ERANGE,
               "pg_assemble: row length %d does not fit in shared memory", L);
    int groups = (int)std::min<size_t>(64, (budget - table_bytes) / per_group);
    groups -= groups % 4;  // whole warps
    const int threads = groups * S::G;
    a.rc = S::R;
    a.bufstride = S::R * L;
    const size_t smem = table_bytes + (size_t)groups * per_group;
    static const bool bulk = [] {
        const char *e = getenv("PG_ASM_BULK");
        return e ? atoi(e) != 0 : false;  // measured 2 % slower at C3 (profiles/r2_assemble_small_bulk_ab.json)
    }();
    auto kern = bulk ? assemble_small_kernel<P, true> : assemble_small_kernel<P, false>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nb = pl->b1 - pl->b0;
    const int64_t grid = std::min<int64_t>((nb + groups - 1) / groups, (int64_t)kNumSMs);
    kern<<<(unsigned)grid, threads, smem, st>>>(a);
    PG_LAUNCH_OK();
    return PG_OK;
}

template <int P, bool ITAB>
static int launch_assemble(const pg_plan *pl, AsmArgs a, cudaStream_t st) {
    using S = Gen<P>;
    const int L = pl->max_rowlen;
    const int rmax = std::max(ndof_edge(pl->p), std::max(ndof_face(pl->p), ndof_volume(pl->p)));
    // tile per warp: at least one row; aim at min(rmax, 3) rows (p <= 4) while >= 8 warps fit per SM
    const size_t budget = 216 * 1024;
    const size_t recs = S::MC * sizeof(IncRecord);
    const size_t want = (size_t)L * 16 * (pl->p <= 4 ? std::min(rmax, 3) : 1) + recs;
    int warps = (int)std::min<size_t>(32, budget / want);
    if (warps < 8) warps = (int)std::min<size_t>(8, budget / ((size_t)L * 16 + recs));
    warps -= warps % S::GROUPS;
    PG_REQUIRE(warps >= S::GROUPS, PG_ERANGE, "pg_assemble: row length %d does not fit in shared memory", L);
    a.bufstride = (int)((budget / warps - recs) / 16);
    a.rc = std::min(rmax, a.bufstride / L);
    const size_t smem = S::GROUPS * (recs + (size_t)a.bufstride * 16);
    void (*kern)(const AsmArgs);
    if constexpr (P == 5) kern = assemble_kernel_capped<P, ITAB>; else kern = assemble_kernel<P, ITAB>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * S::GROUPS, smem));
    PG_REQUIRE(occ >= 1, PG_ERANGE, "pg_assemble: kernel does not fit on an SM (smem %zu)", smem);
    const int64_t nb = pl->b1 - pl->b0;
    const int64_t grid = std::min<int64_t>((nb + S::GROUPS - 1) / S::GROUPS, (int64_t)kNumSMs * occ);
    kern<<<(unsigned)grid, 32 * S::GROUPS, smem, st>>>(a);
    PG_LAUNCH_OK();
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" int pg_assemble(const pg_plan *pl, const double *geo, const uint32_t *code, const double *table,
                           double mass_scale, int apply_dirichlet, double diag, double *vals, void *stream) {
    PG_REQUIRE(pl && geo && code && table, PG_EINVAL, "pg_assemble: null pointer");
    if (pl->b1 == pl->b0) return PG_OK;
    PG_REQUIRE(vals, PG_EINVAL, "pg_assemble: null output");
    PG_REQUIRE(!apply_dirichlet || pl->bd_entity, PG_EINVAL,
               "pg_assemble: apply_dirichlet without pg_plan_set_dirichlet");
    AsmArgs a;
    a.b0 = pl->b0, a.b1 = pl->b1;
    a.ent_order = pl->ent_order, a.row_base = pl->row_base, a.inc_ptr = pl->inc_ptr, a.rec = pl->rec;
    a.rowlen = pl->rowlen, a.selfpos = pl->selfpos, a.valoff = pl->valoff;
    a.hdr = pl->hdr;
    a.bd_entity = apply_dirichlet ? pl->bd_entity : nullptr;
    a.geo = geo, a.code = code, a.table = table;
    a.itable = nullptr, a.inv_dk = a.inv_dm = 0.0;
    if (pl->p >= 3) {
        // gather-friendly copy of the table (lazily built, cached in the plan for this table pointer):
        // exact integer numerators for p <= 5, fp64 for p = 6
        const int64_t nexp = pg_nexp(pl->p), nval = nexp * nexp * 12;
        const bool ints = pl->p <= 5;
        if (pl->itable_src != table) {
            if (!pl->itable)
                PG_CUDA_OK(cudaMalloc((void **)&pl->itable, nval * (ints ? sizeof(int32_t) : sizeof(double))));
            const unsigned nblk = (unsigned)((nval + 255) / 256);
            if (ints)
                build_itable_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(nexp, pl->p, table, pl->itable);
            else
                build_dtable_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(nexp, table,
                                                                            reinterpret_cast<double *>(pl->itable));
            PG_LAUNCH_OK();
            pl->itable_src = table;
        }
        a.itable = pl->itable;
        if (ints) {
            a.inv_dk = 1.0 / int_scale_k(pl->p);
            a.inv_dm = 1.0 / int_scale_m(pl->p);
        }
    }
    a.mass_scale = mass_scale, a.diag = diag;
    a.vals = reinterpret_cast<double2 *>(vals);
    a.rc = 1, a.bufstride = 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pl->p) {
        case 1: return launch_assemble_small<1>(pl, a, st);
        case 2: return launch_assemble_small<2>(pl, a, st);
        case 3: return launch_assemble<3, true>(pl, a, st);
        case 4: return launch_assemble<4, true>(pl, a, st);
        case 5: return launch_assemble<5, true>(pl, a, st);
        case 6: return launch_assemble<6, false>(pl, a, st);
    }
    set_error("pg_assemble: bad order %d", pl->p);
    return PG_EINVAL;
}

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
    const double *geo;
    const uint32_t *code;
    const double *table;
    double mass_scale, diag;
    double2 *vals;
    int rc;         // rows per pass
    int bufstride;  // complex elements of shared tile per group
};

template <int G>
__device__ __forceinline__ void group_sync(unsigned mask) {
    if (G == 32) __syncwarp(); else __syncwarp(mask);
}

template <int P, int G, int THREADS>
__global__ void __launch_bounds__(THREADS) assemble_kernel(const AsmArgs a) {
    using O = Ord<P>;
    constexpr int GROUPS = THREADS / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IncRecord *s_rec = reinterpret_cast<IncRecord *>(smem_raw);
    double2 *s_buf = reinterpret_cast<double2 *>(smem_raw + GROUPS * sizeof(IncRecord));

    const int grp = threadIdx.x / G;
    const int lane = threadIdx.x % G;
    const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    double2 *buf = s_buf + (size_t)grp * a.bufstride;
    IncRecord *rec = s_rec + grp;

    const int64_t ngroups = (int64_t)gridDim.x * GROUPS;
    for (int64_t b = a.b0 + blockIdx.x * (int64_t)GROUPS + grp; b < a.b1; b += ngroups) {
        const int32_t g = __ldg(a.ent_order + b);
        const int r = (int)(__ldg(a.row_base + b + 1) - __ldg(a.row_base + b));
        const int L = __ldg(a.rowlen + g);
        double2 *out = a.vals + __ldg(a.valoff + (b - a.b0));

        if (a.bd_entity && __ldg(a.bd_entity + g)) {
            // Dirichlet entity: identity rows (MatZeroRowsColumns puts `diag` on the diagonal)
            const int sp = __ldg(a.selfpos + g);
            for (int it = lane; it < r * L; it += G) {
                int d = it / L, c = it - d * L;
                out[it] = make_double2(c == sp + d ? a.diag : 0.0, 0.0);
            }
            continue;
        }

        const int32_t i0 = __ldg(a.inc_ptr + g), m = __ldg(a.inc_ptr + g + 1) - i0;
        for (int d0 = 0; d0 < r; d0 += a.rc) {
            const int rcur = min(a.rc, r - d0);
            for (int it = lane; it < rcur * L; it += G) buf[it] = make_double2(0.0, 0.0);
            for (int ia = 0; ia < m; ++ia) {
                group_sync<G>(mask);  // previous pass (or the zeroing) finished; s_rec reusable
                if (lane < 2)
                    reinterpret_cast<int4 *>(rec)[lane] = __ldg(reinterpret_cast<const int4 *>(a.rec + i0 + ia) + lane);
                group_sync<G>(mask);
                const int64_t t = rec->elem;
                const int rslot = rec->slot;
                const unsigned bdm = a.bd_entity ? rec->bdmask : 0u;
                double gf[12];
                {
                    const double2 *gp = reinterpret_cast<const double2 *>(a.geo + t * 12);
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        double2 v = __ldg(gp + i);
                        gf[2 * i] = v.x;
                        gf[2 * i + 1] = v.y;
                    }
                }
                const uint32_t cd = __ldg(a.code + t);
                for (int it = lane; it < rcur * O::n; it += G) {
                    const int dd = it / O::n, k = it - dd * O::n;
                    int ks, kd;
                    slot_of_local<P>(k, ks, kd);
                    if ((bdm >> ks) & 1u) continue;  // Dirichlet column: stays zero
                    double s1, s2;
                    const int Jx = expanded_of_slot<P>(rslot, d0 + dd, cd, s1);
                    const int Kx = expanded_of_slot<P>(ks, kd, cd, s2);
                    double kk, mm;
                    contract12(a.table + ((int64_t)Jx * O::nexp + Kx) * 12, gf, kk, mm);
                    const double s = s1 * s2;
                    double2 *dst = buf + dd * L + rec->slotpos[ks] + kd;
                    double2 cur = *dst;
                    cur.x += s * kk;
                    cur.y += s * (a.mass_scale * mm);
                    *dst = cur;
                }
            }
            group_sync<G>(mask);
            double2 *o2 = out + (int64_t)d0 * L;
            for (int it = lane; it < rcur * L; it += G) o2[it] = buf[it];
            group_sync<G>(mask);
        }
    }
}

template <int P, int G, int THREADS>
static int launch_assemble(const pg_plan *pl, AsmArgs a, cudaStream_t st) {
    constexpr int GROUPS = THREADS / G;
    const int L = pl->max_rowlen;
    int rmax = std::max(ndof_edge(pl->p), std::max(ndof_face(pl->p), ndof_volume(pl->p)));
    // shared tile budget per block: keep >= 2 blocks per SM when possible
    const size_t budget = 100 * 1024;
    size_t per_group = (budget - GROUPS * sizeof(IncRecord)) / GROUPS;
    int rc = (int)std::min<size_t>(rmax, per_group / ((size_t)L * 16));
    if (rc < 1) {
        // one row does not fit the default budget: take (almost) the whole SM
        const size_t big = 220 * 1024;
        per_group = (big - GROUPS * sizeof(IncRecord)) / GROUPS;
        rc = (int)std::min<size_t>(rmax, per_group / ((size_t)L * 16));
        PG_REQUIRE(rc >= 1, PG_ERANGE, "pg_assemble: row length %d does not fit in shared memory", L);
    }
    a.rc = rc;
    a.bufstride = rc * L;
    const size_t smem = GROUPS * sizeof(IncRecord) + (size_t)GROUPS * a.bufstride * 16;
    auto kern = assemble_kernel<P, G, THREADS>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    PG_REQUIRE(occ >= 1, PG_ERANGE, "pg_assemble: kernel does not fit on an SM (smem %zu)", smem);
    const int64_t nb = pl->b1 - pl->b0;
    int64_t grid = std::min<int64_t>((nb + GROUPS - 1) / GROUPS, (int64_t)kNumSMs * occ * 4);
    kern<<<(unsigned)grid, THREADS, smem, st>>>(a);
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
    a.bd_entity = apply_dirichlet ? pl->bd_entity : nullptr;
    a.geo = geo, a.code = code, a.table = table;
    a.mass_scale = mass_scale, a.diag = diag;
    a.vals = reinterpret_cast<double2 *>(vals);
    a.rc = 1, a.bufstride = 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pl->p) {
        case 1: return launch_assemble<1, 8, 256>(pl, a, st);
        case 2: return launch_assemble<2, 32, 256>(pl, a, st);
        case 3: return launch_assemble<3, 32, 128>(pl, a, st);
        case 4: return launch_assemble<4, 32, 128>(pl, a, st);
        case 5: return launch_assemble<5, 32, 128>(pl, a, st);
        case 6: return launch_assemble<6, 32, 128>(pl, a, st);
    }
    set_error("pg_assemble: bad order %d", pl->p);
    return PG_EINVAL;
}

/*
 * petgem_b200 -- C ABI of the B200-native PETGEM hot path.
 *
 * The reference (PETGEM v1.0, pure Python) has no FFI; its hot path is reached
 * through Python calls (kernel.py:64-73 -> petgem/solver.py -> petgem/hvfem.py and
 * petsc4py).  Each entry point below replaces one of those call sites with a
 * batched device kernel; the comment on each cites the reference interface it
 * stands in for (paths relative to /root/reference).  INTEGRATION.md shows the
 * ctypes binding a PETGEM maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every array pointer is a DEVICE pointer owned
 *    by the caller unless the name ends in _host; the library never frees caller
 *    memory and keeps no global state besides the last error string (thread local);
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *    asynchronous on that stream unless stated otherwise;
 *  - every function returns 0 on success or a negative code (PG_E*); it never calls
 *    exit().  The Python layer maps a failure to the reference convention
 *    Print.master(msg); exit(-1) (e.g. hvfem.py:284-286);
 *  - complex numbers are interleaved (re, im) doubles, i.e. numpy complex128 /
 *    PetscScalar in a --with-scalar-type=complex build;
 *  - indices are int32 (dof/entity ids) and int64 (row pointers / nnz offsets).
 *
 * There is no CPU fallback anywhere in this library.
 */
#ifndef PETGEM_B200_H
#define PETGEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_EINVAL -22   /* bad argument (order outside 1..6, null pointer, ...) */
#define PG_ENOMEM -12   /* device allocation failed */
#define PG_ECUDA -5     /* CUDA runtime error; see pg_last_error() */
#define PG_ERANGE -34   /* a size exceeds what the index types can hold */
#define PG_ETIMEDOUT -110 /* a peer GPU did not answer (pg_comm_status) */

#define PG_MAX_ORDER 6
#define PG_SLOTS 11     /* entity slots of a tetrahedron: 6 edges, 4 faces, 1 interior */

int pg_version(void);
const char *pg_last_error(void);

/* number of dofs per element / edge / face / interior for order p (hvfem.py:44-48) */
int pg_ndof_element(int p);
int pg_ndof_edge(int p);
int pg_ndof_face(int p);
int pg_ndof_volume(int p);
/* size of the orientation-expanded function set: 6p + 24p(p-1) + p(p-1)(p-2)/2 */
int pg_nexp(int p);

/* ---------------------------------------------------------------------------
 * a2 + a3: computeJacobian (hvfem.py:101-119) and computeElementOrientation
 * (hvfem.py:122-220), batched over elements; inputs are the rows of the scratch
 * files read at solver.py:193-211.
 *   nodes      [T,12] f64  xyz of the 4 vertices          (nodes.dat)
 *   elemsN     [T,4]  i32  global node ids                (meshConnectivity.dat)
 *   elemsE     [T,6]  i32  global edge ids                (edges.dat)
 *   edgesNodes [T,12] i32  (min,max) node pair per edge   (edgesNodes.dat)
 *   facesEdges [T,12] i32  3 global edges per local face  (facesEdges.dat)
 *   sigma      [T,2]  f64  (sigma_h, sigma_v)             (conductivityModel.dat)
 * outputs
 *   geo  [T,12] f64  gK[6] = sym(J J^T / detJ), gM[6] = sym(detJ J^-T diag(sh,sh,sv) J^-1)
 *                    packed (00,11,22,01,02,12); detJ is SIGNED (hvfem.py:265)
 *   code [T]    u32  bits 0..5 edge orientation, bits 6+3f..8+3f face code 0..5
 * --------------------------------------------------------------------------- */
int pg_element_geometry(int64_t T, const double *nodes, const int32_t *elemsN, const int32_t *elemsE,
                        const int32_t *edgesNodes, const int32_t *facesEdges, const double *sigma,
                        double *geo, uint32_t *code, void *stream);

/* ---------------------------------------------------------------------------
 * a5 + a6: the geometry-independent half of computeElementalMatrices: shape3DETet (hvfem.py:319-464)
 * and the quadrature (compute3DGaussPoints, hvfem.py:1055-1610) folded ONCE per order into the
 * reference-element contraction table the element kernels read:
 *   table [nexp,nexp,12] f64: per expanded pair (J,K) SK[6] then SM[6], packed (00,11,22,01,02,12),
 *   S^{ab}_{JK} = int_master N_J^a N_K^b (+ b<->a for a != b), curls for SK; exact conical Gauss-Jacobi
 *   rule of degree 2p+1.  pg_table_size(p) doubles; synchronous (scratch is freed on return).
 * --------------------------------------------------------------------------- */
int64_t pg_table_size(int p);
int pg_tables_init(int p, double *table, void *stream);
/* shape3DETet(X, Nord, NoriE, NoriF) (hvfem.py:319-464) at ONE master point, evaluated on the host by the
 * same code the device kernels run (a host utility like pg_ndof_element, used to pin the C++ basis against
 * the reference's golden shape functions without a GPU; no hot-path call goes through it):
 *   code as written by pg_element_geometry; xi_host [3]; N_host, C_host [n,3] (function-major). */
int pg_shape_functions_host(int p, uint32_t code, const double *xi_host, double *N_host, double *C_host);

/* ---------------------------------------------------------------------------
 * a4 (+a5, a6): computeElementalMatrices (hvfem.py:223-316), batched.
 *   table [nexp,nexp,12] f64: per expanded pair (J,K) SK[6] then SM[6]
 *         (pg_tables_init; built once per order)
 *   Me, Ke [T,n,n] f64 row-major (either may be NULL)
 * --------------------------------------------------------------------------- */
int pg_element_matrices(int64_t T, int p, const double *geo, const uint32_t *code, const double *table,
                        double *Me, double *Ke, void *stream);

/* The same Me, Ke as two n x n x 3*ngauss products per element on the FP64 tensor pipe (mma.sync.m8n8k4.f64,
 * SASS DMMA): Me = sign(detJ) U^T U, Ke = sign(detJ) V^T V with the mapped, weighted basis values at the Gauss
 * points as U, V (the literal form of the loops at hvfem.py:288-314).  Kept for the A/B against the table
 * contraction (tools/dmma_ab.py, profiles/r2_dmma_ab.json); pg_assemble does not use it.
 *   nodes [T,12], sigma [T,2], code [T] as for pg_element_geometry; phiN, phiC [nexp, ngauss, 3] f64: the
 *   orientation-expanded reference functions / curls at the Gauss points; weights [ngauss] (sum 1/6);
 *   work: pg_phi_gemm_workspace_doubles(T, p, ngauss) doubles; stage 0 = both, 1 = operands only, 2 = GEMMs only. */
int64_t pg_phi_gemm_workspace_doubles(int64_t T, int p, int ngauss);
int pg_element_matrices_phi_gemm(int64_t T, int p, const double *nodes, const double *sigma, const uint32_t *code,
                                 int ngauss, const double *phiN, const double *phiC, const double *weights,
                                 double *work, int stage, double *Me, double *Ke, void *stream);

/* solver.py:223-224: Ae = K - i*omega*mu*M, flattened row-major; mass_scale = -omega*mu.
 *   Ae [T,n,n] complex128 */
int pg_element_systems(int64_t T, int p, const double *geo, const uint32_t *code, const double *table,
                       double mass_scale, double *Ae, void *stream);

/* a7: computeConnectivityDOFS (hvfem.py:15-98): dofs [T,n] i32 */
int pg_connectivity_dofs(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                         int64_t nFaces, int32_t *dofs, void *stream);

/* ---------------------------------------------------------------------------
 * a8 + a9 (symbolic half): what createParallelMatrix + MatSetValues/MatAssembly
 * (parallel.py:150-177, solver.py:188-235) decide about the sparsity pattern:
 * union of element dof cliques, explicit zeros kept, columns ascending.
 *
 * The plan is an opaque host object that owns the device arrays of the symbolic
 * phase (entity incidence lists, per-entity column lists, slot positions).
 *   ent_order_host: NULL = reference numbering (edge dofs, face dofs, interior
 *       dofs; hvfem.py:50-71), else a permutation of the nEnt global entity ids
 *       (edges 0..nE-1, faces nE.., interiors nE+nF..) giving the internal row
 *       order; pg_plan_locality_order() builds the element-major one.
 *   row_begin,row_end: rows (in the numbering in use) owned by this process,
 *       PETSc-style contiguous block; pass 0, -1 for all rows.  Must fall on
 *       entity boundaries (pg_plan_entity_aligned_split helps).
 * Synchronous (host decisions depend on device counts).
 * --------------------------------------------------------------------------- */
typedef struct pg_plan pg_plan;

int pg_plan_create(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                   int64_t nFaces, const int32_t *ent_order_host, int64_t row_begin, int64_t row_end,
                   pg_plan **plan, void *stream);
void pg_plan_destroy(pg_plan *plan);

/* element-major entity order (entities sorted by first incident element, then slot) */
int pg_plan_locality_order(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                           int64_t nFaces, int32_t *ent_order_host, void *stream);

/* same, but "first" is taken in the caller's element traversal: elem_rank [T] i32 (device) = position of
 * each element in that traversal (e.g. a Morton / Hilbert curve through the centroids), NULL = element id.
 * A space-filling traversal keeps the x entries a row touches close together in memory (SpMV re-fetches)
 * and makes contiguous row blocks compact 3-D chunks (multi-GPU halo). */
int pg_plan_ranked_order(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                         int64_t nFaces, const int32_t *elem_rank, int32_t *ent_order_host, void *stream);

int64_t pg_plan_num_dofs(const pg_plan *plan);      /* N (global) */
int64_t pg_plan_num_entities(const pg_plan *plan);  /* entities that carry dofs */
int64_t pg_plan_local_rows(const pg_plan *plan);    /* row_end - row_begin */
int64_t pg_plan_row_begin(const pg_plan *plan);
int64_t pg_plan_nnz(const pg_plan *plan);           /* nnz of the owned rows */
int64_t pg_plan_contributions(const pg_plan *plan); /* sum over owned rows of element contributions */
int pg_plan_max_row_length(const pg_plan *plan);
/* smallest element range [*t_begin, *t_end) containing every element incident to an owned row: the
 * only elements whose geometry (pg_element_geometry) this process needs */
int pg_plan_element_range(const pg_plan *plan, int64_t *t_begin, int64_t *t_end);

/* CSR of the owned rows: rowptr [local_rows+1] i64 (starting at 0), colidx [nnz] i32 (global cols) */
int pg_plan_csr(const pg_plan *plan, int64_t *rowptr, int32_t *colidx, void *stream);
/* perm [N] i32: row/col index (numbering in use) of every reference dof id */
int pg_plan_dof_permutation(const pg_plan *plan, int32_t *perm, void *stream);
/* smallest entity-aligned row >= row (numbering in use); host, synchronous */
int64_t pg_plan_entity_aligned_row(const pg_plan *plan, int64_t row);

/* ---------------------------------------------------------------------------
 * a10 (fused form): tell the plan which entities are Dirichlet so that
 * pg_assemble can apply MatZeroRowsColumns (solver.py:562) while it assembles.
 * In PETGEM the boundary dofs are always ALL dofs of boundary edges and boundary
 * faces (mesh.py:280-321), so the set is given per entity:
 *   bd_entity [nEnt] u8 (global entity ids: edges, then faces, then interiors);
 *   NULL clears it.  elemsE/elemsF as given to pg_plan_create.
 * --------------------------------------------------------------------------- */
int pg_plan_set_dirichlet(pg_plan *plan, const int32_t *elemsE, const int32_t *elemsF, const uint8_t *bd_entity,
                          void *stream);

/* ---------------------------------------------------------------------------
 * a1 + a4 + a9 (numeric half): the element loop solver.py:191-230 fused with the
 * scatter-add: every owned row gathers its element contributions in ascending
 * element order (the order MatSetValues(ADD_VALUES) is called in), no atomics,
 * bit-reproducible.  vals [nnz] complex128 in pg_plan_csr order.
 *   apply_dirichlet != 0: rows/columns of the entities given to
 *       pg_plan_set_dirichlet are zeroed and `diag` put on their diagonal, i.e.
 *       the result equals assembly followed by A.zeroRowsColumns(bd, diag).
 * --------------------------------------------------------------------------- */
int pg_assemble(const pg_plan *plan, const double *geo, const uint32_t *code, const double *table,
                double mass_scale, int apply_dirichlet, double diag, double *vals, void *stream);

/* a10: A.zeroRowsColumns(bd) (solver.py:562; petsc4py default diag = 1.0) on an
 * assembled CSR block; bd_mask [N] u8 over global columns; row_begin = first owned row. */
int pg_zero_rows_columns(int64_t local_rows, int64_t row_begin, const int64_t *rowptr, const int32_t *colidx,
                         const uint8_t *bd_mask, double diag, double *vals, void *stream);

/* ---------------------------------------------------------------------------
 * a11: the arithmetic inside KSP.solve (solver.py:584-590): MatMult and the
 * Vec kernels GMRES/BiCGStab call.  x is the FULL (gathered) vector, y the owned
 * block.  Scalars live in device memory so a Krylov cycle needs no host sync.
 * --------------------------------------------------------------------------- */
int pg_spmv(int64_t local_rows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
            const double *x, double *y, void *stream);
/* y = dscale .* (A x): MatMult fused with the PCJACOBI application (dscale = 1/diag, may be NULL) */
int pg_spmv_scaled(int64_t local_rows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
                   const double *x, const double *dscale, double *y, void *stream);
/* Entity-blocked MatMult for p = 2 (same result as pg_spmv_scaled on the plan's CSR): reads the plan's
 * per-entity column lists (4 B per 2x2 block instead of 16 B of colidx).
 *   colstart: NULL = the plan's own (global columns of the numbering in use), or a remapped copy of
 *             pg_plan_column_starts() indexing a [own | halo] vector in multi-GPU runs */
int pg_spmv_blocked(const pg_plan *plan, const int32_t *colstart, const double *vals, const double *x,
                    const double *dscale, double *y, void *stream);
/* the same for the entities [ent_begin, ent_end) of the plan's processing order only, and the interior range of a
 * row block whose columns are numbered [own | halo]: every entity of [*ent_begin_host, *ent_end_host) has all its
 * columns below n_own, so its rows can be multiplied while the halo is still in flight (pg_comm_push), the two
 * remaining ranges after pg_comm_wait */
int pg_spmv_blocked_range(const pg_plan *plan, int64_t ent_begin, int64_t ent_end, const int32_t *colstart,
                          const double *vals, const double *x, const double *dscale, double *y, void *stream);
int pg_plan_halo_split(const pg_plan *plan, const int32_t *colstart, int64_t n_own, int64_t *ent_begin_host,
                       int64_t *ent_end_host, void *stream);
int64_t pg_plan_num_column_entities(const pg_plan *plan);
int pg_plan_column_starts(const pg_plan *plan, int32_t *colstart, void *stream);
/* diag [local_rows] complex128 of the owned block (for PCJACOBI) */
int pg_csr_diagonal(int64_t local_rows, int64_t row_begin, const int64_t *rowptr, const int32_t *colidx,
                    const double *vals, double *diag, void *stream);

/* y += alpha x ; alpha read from device memory (complex) */
int pg_zaxpy(int64_t n, const double *alpha, const double *x, double *y, void *stream);
/* y = x + beta y */
int pg_zaypx(int64_t n, const double *beta, const double *x, double *y, void *stream);
/* w = a x + b y + c z (any coefficient pointer NULL = term skipped); w may alias x,y,z */
int pg_zaxpbypcz(int64_t n, const double *a, const double *x, const double *b, const double *y, const double *c,
                 const double *z, double *w, void *stream);
/* x *= alpha, alpha complex on device; if inv_real != 0: x *= 1/Re(alpha) (normalisation) */
int pg_zscal(int64_t n, const double *alpha, int inv_real, double *x, void *stream);
/* z = x .* y (PCJACOBI apply with y = 1/diag) */
int pg_zpointwise_mult(int64_t n, const double *x, const double *y, double *z, void *stream);
/* out[0] = sum conj(x_i) y_i  (VecDot(y,x) convention: PETSc conjugates the 2nd arg);
 * deterministic two-stage tree, work = pg_reduce_workspace_bytes() */
int pg_zdotc(int64_t n, const double *x, const double *y, double *out, void *work, void *stream);
/* out[0] = sum x_i y_i (VecTDot, no conjugation: the inner product of COCG on the complex symmetric A) */
int pg_zdotu(int64_t n, const double *x, const double *y, double *out, void *work, void *stream);
/* VecMDot: out[i] = sum conj(V_i) . w for k vectors V_i = V + i*ldv (complex elements), one pass over w */
int pg_zmdotc(int64_t n, int k, const double *V, int64_t ldv, const double *w, double *out, void *work,
              void *stream);
/* VecMAXPY: w += sum_i scale*alpha[i] V_i (alpha complex on device), one pass over w */
int pg_zmaxpy(int64_t n, int k, const double *alpha, double scale, const double *V, int64_t ldv, double *w,
              void *stream);
/* VecMAXPY followed by VecNorm of the result, fused: out[0] = sum |w_i|^2 after the update */
int pg_zmaxpy_nrm2sq(int64_t n, int k, const double *alpha, double scale, const double *V, int64_t ldv, double *w,
                     double *out, void *work, void *stream);
/* out[0] = a/b (negated when negate != 0), out[1] = -out[0]; device scalars: the Krylov coefficients
 * (alpha = rho / p^T A p, beta = rho_new / rho) never visit the host */
int pg_zdiv(const double *a, const double *b, int negate, double *out, void *stream);
/* y = alpha x, or y = x / Re(alpha) when inv_real != 0 (VecCopy + VecScale fused) */
int pg_zcopy_scaled(int64_t n, const double *alpha, int inv_real, const double *x, double *y, void *stream);
/* out[0] = sum |x_i|^2 (real, stored as complex with zero imaginary part) */
int pg_dznrm2sq(int64_t n, const double *x, double *out, void *work, void *stream);
int64_t pg_reduce_workspace_bytes(int kmax);

/* ---------------------------------------------------------------------------
 * Several right-hand sides at once: K sources (or the two MT polarizations) share A, so the K
 * solves that the reference runs one after the other (solver.py:584-590, one KSP.solve per
 * right-hand side) advance in lockstep and the matrix is streamed once per iteration for all of
 * them.  Vectors are INTERLEAVED: X[(i*k + r)] (complex) = entry i of right-hand side r,
 * k = 1, 2, 4 or 8 (pad with zero right-hand sides).  Per-right-hand-side scalars are arrays of
 * k complex numbers in device memory.  work = pg_reduce_workspace_bytes(2*k).
 * --------------------------------------------------------------------------- */
/* Y = dscale .* (A X) for k right-hand sides (dscale may be NULL); X is the FULL interleaved block */
int pg_spmm(int64_t local_rows, const int64_t *rowptr, const int32_t *colidx, const double *vals, int k,
            const double *X, const double *dscale, double *Y, void *stream);
/* the same through the plan's 2x2 entity blocks (p = 2, k = 2, 4, 8): see pg_spmv_blocked */
int pg_spmm_blocked(const pg_plan *plan, const int32_t *colstart, const double *vals, int k, const double *X,
                    const double *dscale, double *Y, void *stream);
/* pg_spmv_blocked / pg_spmm_blocked (k = 1, 4, 8) fused with the dot product COCG / COCR take right after it:
 * out[r] = sum_rows X[row,r] Y[row,r] (unconjugated; X[row] = the own entries of the multiplied vector), summed
 * in a fixed order.  work = pg_spmv_dot_workspace_bytes(plan, k); arrays 32-byte aligned (PG_EINVAL otherwise:
 * fall back to the separate calls). */
int64_t pg_spmv_dot_workspace_bytes(const pg_plan *plan, int k);
int pg_spmm_blocked_dot(const pg_plan *plan, const int32_t *colstart, const double *vals, int k, const double *X,
                        const double *dscale, double *Y, double *out, void *work, void *stream);
int pg_spmm_blocked_range(const pg_plan *plan, int64_t ent_begin, int64_t ent_end, const int32_t *colstart,
                          const double *vals, int k, const double *X, const double *dscale, double *Y, void *stream);
/* Y[:,r] += alpha[r] X[:,r] */
int pg_zbaxpy(int64_t n, int k, const double *alpha, const double *X, double *Y, void *stream);
/* Y[:,r] = X[:,r] + beta[r] Y[:,r] */
int pg_zbaypx(int64_t n, int k, const double *beta, const double *X, double *Y, void *stream);
/* Y[i,r] = d[i] X[i,r] (PCJACOBI apply for every right-hand side) */
int pg_zbscale_rows(int64_t n, int k, const double *d, const double *X, double *Y, void *stream);
/* out[r] = sum_i X[i,r] Y[i,r] (no conjugation) */
int pg_zbdotu(int64_t n, int k, const double *X, const double *Y, double *out, void *work, void *stream);
/* out[r] = sum_i X[i,r] w[i] Y[i,r] (no conjugation; w NULL: no weight) */
int pg_zbdotu_w(int64_t n, int k, const double *X, const double *Y, const double *w, double *out, void *work,
                void *stream);
/* COCR, the element-wise passes of an iteration: X += alpha P, RT -= alpha dinv .* AP (alpha2 = {alpha[k],
 * -alpha[k]}, dinv NULL: no preconditioner) and P = RT + beta P, AP = ART + beta AP */
int pg_cocr_update(int64_t n, int k, const double *alpha2, const double *P, const double *AP, const double *dinv,
                   double *X, double *RT, void *stream);
int pg_cocr_direction(int64_t n, int k, const double *beta, const double *RT, const double *ART, double *P,
                      double *AP, void *stream);
/* pg_cocr_direction, and in the same pass out[r] = sum_i AP[i,r] w[i] AP[i,r] of the NEW A p (w = D^-1 or NULL):
 * the dot product that opens the next COCR iteration, without re-reading A p and D^-1 */
int pg_cocr_direction_dot(int64_t n, int k, const double *beta, const double *RT, const double *ART, const double *w,
                          double *P, double *AP, double *out, void *work, void *stream);
/* out[r] = sum_i |X[i,r]|^2 */
int pg_zbnrm2sq(int64_t n, int k, const double *X, double *out, void *work, void *stream);
/* out[r] = a[r] / b[r] (0 where b[r] == 0), out[k + r] = -out[r] */
int pg_zbdiv(int k, const double *a, const double *b, double *out, void *stream);
/* fused COCG update: X += alpha P, R -= alpha Q (alpha2 = {alpha[k], -alpha[k]} from pg_zbdiv),
 * Z = dinv .* R (dinv NULL: Z = R), out[r] = R[:,r]^T Z[:,r], out[k + r] = |Z[:,r]|^2 */
int pg_cocg_step(int64_t n, int k, const double *alpha2, const double *P, const double *Q, const double *dinv,
                 double *X, double *R, double *Z, double *out, void *work, void *stream);

/* The whole solve as one call on one GPU (full matrix), zero initial guess, optional Jacobi preconditioning,
 * convergence on ||M^-1 r|| <= rtol ||M^-1 b|| (KSP defaults with left preconditioning; solver.py:584-590):
 *   PG_KSP_COCG  -ksp_type cg -ksp_cg_type symmetric   (A is complex symmetric)
 *   PG_KSP_COCR  -ksp_type cr (conjugate-orthogonal conjugate residuals)
 *   PG_KSP_BCGS  -ksp_type bcgs
 *   PG_KSP_GMRES -ksp_type gmres, restart = -ksp_gmres_restart (PETSc default 30), classical Gram-Schmidt
 * COCG / COCR check the residual every check_every iterations (coefficients stay on the device); BiCGStab
 * and GMRES read their scalars every iteration.  work = pg_krylov_workspace_bytes(n, method, restart) bytes
 * of device memory; iterations and rel_residual are host outputs; returns PG_OK also when maxit is reached
 * (compare rel_residual with rtol). */
#define PG_KSP_COCG 0
#define PG_KSP_COCR 1
#define PG_KSP_BCGS 2
#define PG_KSP_GMRES 3
int64_t pg_krylov_workspace_bytes(int64_t n, int method, int restart);
int pg_krylov_solve(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *b,
                    double *x, int method, int restart, int jacobi, double rtol, int maxit, int check_every,
                    void *work, int *iterations, double *rel_residual, void *stream);

/* ---------------------------------------------------------------------------
 * f1: the callers on either side of the solve, on the device.
 * pg_locate_points: containing element of each point, lowest element index whose barycentric coordinates
 *   are all >= -tol, -1 if none (postprocessing.py:532-539 / preprocessing.py:414-420 use
 *   Delaunay.find_simplex on the mesh): points [npts,3] f64 -> pt_elem [npts] i32.  Synchronous.
 * pg_interpolate_fields: fieldInterpolator (postprocessing.py:479-616): for every point the basis of its
 *   element is evaluated at the point, E = sum_j x_j N_j, H = sum_j x_j curl N_j / (i omega mu):
 *   fields [npts,6] complex128 (Ex,Ey,Ez,Hx,Hy,Hz); x [N] complex128; perm NULL: x in the reference dof
 *   numbering, else perm[ref dof] = index into x (pg_plan_dof_permutation).  NaN where pt_elem < 0.
 * pg_csem_rhs: the dipole right-hand side (solver.py:247-316): b[dof_j] += i omega mu (moment . N_j(x_src))
 *   for the dofs of the source element; position/moment are HOST xyz triples (moment = current * length *
 *   rotation(azimuth, dip), hvfem.py:2303-2344); b holds the rows [row_begin, row_begin+local_rows) of the
 *   numbering in use (perm as above).  Synchronous.
 * --------------------------------------------------------------------------- */
int pg_locate_points(int64_t T, const double *nodes, int64_t npts, const double *points, double tol,
                     int32_t *pt_elem, void *stream);
int pg_interpolate_fields(int64_t npts, const double *points, const int32_t *pt_elem, int p, const double *nodes,
                          const uint32_t *code, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges,
                          int64_t nFaces, const int32_t *perm, const double *x, double omega, double mu, double *fields,
                          void *stream);
int pg_csem_rhs(int p, int64_t source_elem, const double *position_host, const double *moment_host, const double *nodes,
                const uint32_t *code, const int32_t *elemsE, const int32_t *elemsF, int64_t nEdges, int64_t nFaces,
                const int32_t *perm, int64_t row_begin, int64_t local_rows, double omega, double mu, double *b,
                void *stream);

/* ---------------------------------------------------------------------------
 * Gradient-space (Hiptmair) preconditioner, the GPU-friendly stand-in for the -pc_type sor / asm /
 * gamg of the reference's option files (examples/case1..5 petsc.opts, read by KSP.setFromOptions at
 * solver.py:586-589):  M^-1 = D^-1 + G diag(G^T A G)^-1 G^T, G = discrete gradient of the H1 vertex
 * (and, p >= 2, quadratic edge) functions in the Nedelec dofs, a REAL sparse matrix with <= 3 entries
 * per row.  Complex symmetric, so COCG / COCR stay valid.  petgem_b200/gradient.py builds G.
 * --------------------------------------------------------------------------- */
/* Y[i,:] = s[i] * (sum_j vals[j] X[colidx[j],:]) + a[i] * Z[i,:] for a real CSR matrix (rowptr i32 [rows+1],
 * colidx i32, vals f64) and k interleaved complex vectors (k = 1, 2, 4, 8).  s, a: complex [rows] or NULL
 * (s NULL: factor 1; a NULL: no Z term). */
int pg_rcsr_apply(int64_t rows, const int32_t *rowptr, const int32_t *colidx, const double *vals, int k,
                  const double *X, const double *s, const double *a, const double *Z, double *Y, void *stream);
/* out[k] = sum_{i < a_rows} sum_j R[k,i] A[i,j] R[k,j]: diagonal of R A R^T restricted to the owned rows of A
 * (R = G^T as real CSR with columns sorted, indexed like the columns of A; A complex CSR, a_rows owned rows).
 * On several GPUs the partial sums of the ranks add up to diag(G^T A G). */
int pg_galerkin_diagonal(int64_t rows, const int32_t *r_rowptr, const int32_t *r_colidx, const double *r_vals,
                         int64_t a_rows, const int64_t *a_rowptr, const int32_t *a_colidx, const double *a_vals,
                         double *out, void *stream);
/* d[i] = 1/d[i] (complex), 0 where d[i] == 0 or mask[i] == 0 (mask u8 [n] or NULL) */
int pg_masked_reciprocal(int64_t n, const uint8_t *mask, double *d, void *stream);

/* ---------------------------------------------------------------------------
 * e: the two collectives of a multi-GPU Krylov iteration, over peer memory (NVLink / NVSwitch) instead of
 * a library: the VecScatter inside MatMult and the MPI_Allreduce inside VecDot / VecNorm of KSP.solve
 * (solver.py:584-590) on the row-partitioned objects of createParallelMatrix / createParallelVector
 * (parallel.py:150-203).  One process per GPU; every process maps the buffers of the others (CUDA IPC).
 * All calls are plain kernel launches with device-resident sequence counters (CUDA-graph capturable);
 * waits are bounded by the timeout given to pg_comm_create and raise a sticky error (pg_comm_status).
 *
 * pg_ipc_*: device memory that other processes of the box can map.  handle_host: PG_IPC_HANDLE_BYTES bytes
 *   the caller ships to the peers by whatever transport it has (MPI, torch.distributed, a file).
 * pg_comm_create: ctrl_host[r] = the control block (pg_comm_ctrl_bytes() zeroed bytes from pg_ipc_alloc) of
 *   rank r as mapped in THIS process (own pointer at [rank]).
 * pg_comm_allreduce: out[0..k) = sum over ranks of in[0..k) (k complex scalars, k <= PG_COMM_MAX_REDUCE), added
 *   in rank order on every rank (bit-identical everywhere); out may alias in.
 * pg_comm_push: halo push on channel `chan`: for every destination rank d and entry e of its segment
 *   [seg_host[d], seg_host[d+1]):  dst_host[d][(e - seg_host[d])*k + r] = x[send_idx[e]*k + r]  (complex,
 *   k interleaved right-hand sides), dst_host[d] being a pointer INTO rank d's mapped buffer; then a flag on d.
 *   Before writing to d it waits for d's acknowledgement of the previous push on the channel.
 * pg_comm_wait / pg_comm_ack: reader side of a channel, bracketing the kernels that read the pushed entries:
 *   wait until every rank of from_mask (bit r = rank r sends to me) has pushed; ack tells them the entries
 *   have been consumed.  Every push on a channel must be matched by one wait + ack on each destination.
 * --------------------------------------------------------------------------- */
#define PG_IPC_HANDLE_BYTES 64
#define PG_COMM_MAX_RANKS 16
#define PG_COMM_MAX_CHANNELS 64
#define PG_COMM_MAX_REDUCE 64
typedef struct pg_comm pg_comm;

/* Halo set-up of a row block [row_begin, row_end) (what MatSetUpMultiply / VecScatterCreate do for PETSc's MPIAIJ):
 * pg_halo_columns: the sorted set of GLOBAL columns of the block that other ranks own -> ext [<= nnz] i32 (device),
 *   *n_ext_host entries.  Their owners follow from the row cuts (binary search on the host); every rank sends
 *   each owner its part of the list (one all-to-all of the caller's transport) and receives, per destination, the
 *   rows it must push: send_idx = received global rows - row_begin, segments in rank order (pg_comm_push).
 * pg_halo_remap: colidx (or the plan's column-entity starts, pg_plan_column_starts) -> the [own | halo] numbering:
 *   owned column c -> c - row_begin, external column -> (row_end - row_begin) + its position in ext.
 * Synchronous. */
int pg_halo_columns(int64_t nnz, const int32_t *colidx, int64_t row_begin, int64_t row_end, int32_t *ext,
                    int64_t *n_ext_host, void *stream);
int pg_halo_remap(int64_t nnz, int32_t *colidx, int64_t row_begin, int64_t row_end, const int32_t *ext, int64_t n_ext,
                  void *stream);
int pg_ipc_alloc(int64_t bytes, void **ptr, void *handle_host);
int pg_ipc_open(const void *handle_host, void **ptr);
int pg_ipc_close(void *ptr);
int pg_ipc_free(void *ptr);
int64_t pg_comm_ctrl_bytes(void);
int pg_comm_create(int rank, int world, void *const *ctrl_host, double timeout_s, pg_comm **comm);
void pg_comm_destroy(pg_comm *comm);
int pg_comm_allreduce(pg_comm *comm, int k, const double *in, double *out, void *stream);
int pg_comm_push(pg_comm *comm, int chan, int k, const double *x, const int32_t *send_idx, const int64_t *seg_host,
                 void *const *dst_host, void *stream);
int pg_comm_wait(pg_comm *comm, int chan, uint32_t from_mask, void *stream);
int pg_comm_ack(pg_comm *comm, int chan, uint32_t from_mask, void *stream);
/* PG_OK, or PG_ETIMEDOUT once any wait of this rank has timed out (synchronises the stream) */
int pg_comm_status(pg_comm *comm, void *stream);

/* CUDA graph of a batch of the calls above (the launches of the Krylov iterations between two host
 * checks of the residual): begin capture on a NON-default stream, issue the calls, end -> executable
 * graph, launch it any number of times. */
int pg_graph_begin(void *stream);
int pg_graph_end(void *stream, void **graph_exec);
int pg_graph_launch(void *graph_exec, void *stream);
void pg_graph_destroy(void *graph_exec);

/* L2 residency of the vector MatMult gathers from (x is re-read ~20 times per entry, the matrix streams once).
 * pg_l2_persist: access-policy window on `stream` over [ptr, ptr+bytes): lines of the window are kept in the
 *   persisting set-aside of the L2 (all of them when the window fits pg_l2_persist_capacity(), else the
 *   fraction hit_ratio, 0 = capacity / bytes), everything else streams; ptr NULL clears the window and resets
 *   the persisting lines.  Applies to the kernels launched on the stream afterwards (captured graphs keep it).
 * pg_l2_fetch_granularity: cudaLimitMaxL2FetchGranularity (32, 64 or 128 bytes; scattered 32-byte gathers).
 * pg_tune_spmv_hints: per-load eviction hints of the MatMult kernels (0: ld.cs streams; 1: evict-first
 *   streams, the default; 2: evict-first streams + evict-last x; 3: as 1 and the streams are not allocated in L1; -1: PG_SPMV_HINTS /
 *   default).  pg_tune_spmv_chunk: consecutive 32-entity tiles per thread block of pg_spmv_blocked. */
int pg_l2_persist(const void *ptr, int64_t bytes, double hit_ratio, void *stream);
int64_t pg_l2_persist_capacity(void);
int pg_l2_fetch_granularity(int bytes);
int pg_tune_spmv_hints(int mode);
int pg_tune_spmv_chunk(int chunk);
/* schedule of pg_spmm_blocked (k = 4, 8): 1 = the column-entity indices of a row entity are loaded up front and
 * handed out by shuffle (no index -> gather dependency per step), 0 = plain loop, -1 = PG_SPMM_PF / default */
int pg_tune_spmm_prefetch(int mode);

#ifdef __cplusplus
}
#endif
#endif /* PETGEM_B200_H */

/* A caller of the C ABI that is not Python: reads the rows of the scratch files of a PETGEM case (raw binary
 * dumps of nodes.dat, meshConnectivity.dat, edges.dat, ... written by tests/test_c_abi.py), and runs the
 * whole hot path through include/petgem_b200.h: reference-element tables, element geometry, symbolic plan,
 * fused assembly with Dirichlet rows, the CSEM right-hand side, the Krylov solve and the receiver
 * interpolation (kernel.py:64-73 -> Solver.assembly / Solver.run, solver.py:131-598; postprocessing.py:479-616).
 * Output: fields.bin [npts][6] complex128 and a line "iterations rel_residual N nnz" on stdout.
 *
 *   cc -I include -I $CUDA/include tests/c_abi/solve_case.c -L petgem_b200 -lpetgem_b200 -L $CUDA/lib64 -lcudart -lm
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "petgem_b200.h"

#define CK(call)                                                                            \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != 0) {                                                                     \
            fprintf(stderr, "%s:%d %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc_, pg_last_error()); \
            exit(2);                                                                        \
        }                                                                                   \
    } while (0)
#define CU(call)                                                                            \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            exit(3);                                                                        \
        }                                                                                   \
    } while (0)

static void *load(const char *dir, const char *name, size_t bytes) { /* file -> device */
    char path[4096];
    snprintf(path, sizeof path, "%s/%s.bin", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(4); }
    void *h = malloc(bytes ? bytes : 1);
    if (fread(h, 1, bytes, f) != bytes) { fprintf(stderr, "%s: short read\n", path); exit(4); }
    fclose(f);
    void *d = NULL;
    CU(cudaMalloc(&d, bytes ? bytes : 16));
    CU(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    free(h);
    return d;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <case dir> <method 0..3>\n", argv[0]); return 1; }
    const char *dir = argv[1];
    const int method = atoi(argv[2]);
    char path[4096];
    snprintf(path, sizeof path, "%s/manifest.txt", dir);
    FILE *mf = fopen(path, "r");
    if (!mf) { perror(path); return 4; }
    long long T, nNodes, nE, nF, npts, nbd;
    int p;
    double omega, mu, src[3], moment[3], rtol;
    if (fscanf(mf, "%lld %lld %lld %lld %d %lld %lld %lf %lf %lf %lf %lf %lf %lf %lf %lf", &T, &nNodes, &nE, &nF, &p, &npts,
               &nbd, &omega, &mu, &src[0], &src[1], &src[2], &moment[0], &moment[1], &moment[2], &rtol) != 16) {
        fprintf(stderr, "bad manifest\n");
        return 4;
    }
    fclose(mf);
    (void)nNodes;
    CU(cudaSetDevice(0));

    /* rows of the scratch files (solver.py:193-211) */
    double *nodes = load(dir, "nodes", (size_t)T * 12 * 8), *sigma = load(dir, "sigma", (size_t)T * 2 * 8);
    int32_t *elemsN = load(dir, "elemsN", (size_t)T * 4 * 4), *elemsE = load(dir, "elemsE", (size_t)T * 6 * 4);
    int32_t *edgesNodes = load(dir, "edgesNodes", (size_t)T * 12 * 4), *facesEdges = load(dir, "facesEdges", (size_t)T * 12 * 4);
    int32_t *elemsF = load(dir, "elemsF", (size_t)T * 4 * 4);
    const long long nEnt = nE + (p >= 2 ? nF : 0) + (p >= 3 ? T : 0);
    uint8_t *bd_entity = load(dir, "bd_entity", (size_t)nEnt);
    double *points = load(dir, "receivers", (size_t)npts * 3 * 8);

    /* a5 + a6: reference-element tables, once per order */
    double *table;
    CU(cudaMalloc((void **)&table, (size_t)pg_table_size(p) * 8));
    CK(pg_tables_init(p, table, NULL));
    /* a2 + a3 */
    double *geo;
    uint32_t *code;
    CU(cudaMalloc((void **)&geo, (size_t)T * 12 * 8));
    CU(cudaMalloc((void **)&code, (size_t)T * 4));
    CK(pg_element_geometry(T, nodes, elemsN, elemsE, edgesNodes, facesEdges, sigma, geo, code, NULL));
    /* a8 + a9 symbolic, a10 fused */
    pg_plan *plan = NULL;
    CK(pg_plan_create(T, p, elemsE, elemsF, nE, nF, NULL, 0, -1, &plan, NULL));
    CK(pg_plan_set_dirichlet(plan, elemsE, elemsF, bd_entity, NULL));
    const int64_t N = pg_plan_num_dofs(plan), nnz = pg_plan_nnz(plan);
    int64_t *rowptr;
    int32_t *colidx;
    double *vals, *b, *x;
    CU(cudaMalloc((void **)&rowptr, (size_t)(N + 1) * 8));
    CU(cudaMalloc((void **)&colidx, (size_t)nnz * 4));
    CU(cudaMalloc((void **)&vals, (size_t)nnz * 16));
    CU(cudaMalloc((void **)&b, (size_t)N * 16));
    CU(cudaMalloc((void **)&x, (size_t)N * 16));
    CK(pg_plan_csr(plan, rowptr, colidx, NULL));
    /* a1 + a4 + a9 numeric: Ae = K - i omega mu M scattered into the CSR, Dirichlet rows -> identity */
    CK(pg_assemble(plan, geo, code, table, -omega * mu, 1, 1.0, vals, NULL));

    /* right-hand side (solver.py:247-316) and b.setValues(bd, 0) (solver.py:565-567) */
    CU(cudaMemset(b, 0, (size_t)N * 16));
    double *srcd;
    int32_t *src_elem_d, src_elem;
    CU(cudaMalloc((void **)&srcd, 24));
    CU(cudaMalloc((void **)&src_elem_d, 4));
    CU(cudaMemcpy(srcd, src, 24, cudaMemcpyHostToDevice));
    CK(pg_locate_points(T, nodes, 1, srcd, 1e-12, src_elem_d, NULL));
    CU(cudaMemcpy(&src_elem, src_elem_d, 4, cudaMemcpyDeviceToHost));
    if (src_elem < 0) { fprintf(stderr, "source outside the mesh\n"); return 5; }
    CK(pg_csem_rhs(p, src_elem, src, moment, nodes, code, elemsE, elemsF, nE, nF, NULL, 0, N, omega, mu, b, NULL));
    {
        uint8_t *bdh = malloc((size_t)nEnt);
        double zero[2 * 64] = {0};
        CU(cudaMemcpy(bdh, bd_entity, (size_t)nEnt, cudaMemcpyDeviceToHost));
        for (long long g = 0; g < nEnt; ++g) {
            if (!bdh[g]) continue;
            /* dof ids of entity g (hvfem.py:50-71) */
            const long long rows = g < nE ? pg_ndof_edge(p) : g < nE + nF ? pg_ndof_face(p) : pg_ndof_volume(p);
            const long long first = g < nE ? g * pg_ndof_edge(p)
                                    : g < nE + nF ? nE * pg_ndof_edge(p) + (g - nE) * pg_ndof_face(p)
                                                  : nE * pg_ndof_edge(p) + nF * pg_ndof_face(p) + (g - nE - nF) * pg_ndof_volume(p);
            CU(cudaMemcpy(b + 2 * first, zero, (size_t)rows * 16, cudaMemcpyHostToDevice));
        }
        free(bdh);
    }

    /* a11: the solve */
    const int restart = 30;
    void *work;
    CU(cudaMalloc(&work, (size_t)pg_krylov_workspace_bytes(N, method, restart)));
    int its = 0;
    double rel = 0.0;
    CK(pg_krylov_solve(N, rowptr, colidx, vals, b, x, method, restart, 1, rtol, 100000, 10, work, &its, &rel, NULL));

    /* true residual ||b - A x|| / ||b|| through pg_spmv (host reduction: this is a test program) */
    double *y;
    CU(cudaMalloc((void **)&y, (size_t)N * 16));
    CK(pg_spmv(N, rowptr, colidx, vals, x, y, NULL));
    double *hy = malloc((size_t)N * 16), *hb = malloc((size_t)N * 16);
    CU(cudaMemcpy(hy, y, (size_t)N * 16, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hb, b, (size_t)N * 16, cudaMemcpyDeviceToHost));
    double rr = 0.0, bb = 0.0;
    for (int64_t i = 0; i < 2 * N; ++i) {
        rr += (hb[i] - hy[i]) * (hb[i] - hy[i]);
        bb += hb[i] * hb[i];
    }

    /* f1: receiver fields (postprocessing.py:479-616) */
    int32_t *pt_elem;
    double *fields;
    CU(cudaMalloc((void **)&pt_elem, (size_t)npts * 4));
    CU(cudaMalloc((void **)&fields, (size_t)npts * 6 * 16));
    CK(pg_locate_points(T, nodes, npts, points, 1e-12, pt_elem, NULL));
    CK(pg_interpolate_fields(npts, points, pt_elem, p, nodes, code, elemsE, elemsF, nE, nF, NULL, x, omega, mu, fields, NULL));
    double *hf = malloc((size_t)npts * 6 * 16);
    CU(cudaMemcpy(hf, fields, (size_t)npts * 6 * 16, cudaMemcpyDeviceToHost));
    snprintf(path, sizeof path, "%s/fields.bin", dir);
    FILE *of = fopen(path, "wb");
    fwrite(hf, 16, (size_t)npts * 6, of);
    fclose(of);
    printf("%d %.6e %lld %lld %.6e\n", its, rel, (long long)N, (long long)nnz, sqrt(rr / bb));
    pg_plan_destroy(plan);
    return 0;
}

#!/usr/bin/env python3
"""Time the SpMV variants (CSR / entity-blocked, cache-hint modes via PG_SPMV_HINTS) on a p=2 box."""
import argparse
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--morton", type=int, default=0)
ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity in bytes (32, 64, 128; 0 = default)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
if args.l2_fetch:
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    val = ctypes.c_size_t(0)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(args.l2_fetch))  # cudaLimitMaxL2FetchGranularity
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print("cudaLimitMaxL2FetchGranularity -> rc %d, now %d" % (rc, val.value))
tab = bench.build_case(args.m, 2)
rows = bench.host_rows(tab)
el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"], rows["elemsF"],
                 rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
from petgem_b200.device import morton_element_rank  # noqa: E402
rank = morton_element_rank(rows["nodes"]) if args.morton else None
plan = AssemblyPlan(el, 2, order="locality", elem_rank=rank)
plan.set_dirichlet(bench.bd_entities(tab, 2, plan.nEnt))
g, c = el.geometry()
vals = plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True)
torch.cuda.synchronize()
ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ea.record()
for _ in range(5):
    plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True, out=vals)
eb.record()
torch.cuda.synchronize()
asm_ms = ea.elapsed_time(eb) / 5
rowptr, colidx = plan.csr()
x = torch.randn(plan.N, dtype=torch.complex128, device=dev)
out = {}
for name, A in (("csr", CSRMatrix(rowptr, colidx, vals, plan.N)),
                ("blocked", CSRMatrix(rowptr, colidx, vals, plan.N, plan=plan, blocked=True))):
    y = A.mult(x)
    for _ in range(3):
        A.mult(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        A.mult(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    out[name] = (ms, (20.0 * plan.nnz + 40.0 * plan.N) / ms / 1e6)
print("pf=%s l2fetch=%d " % (os.environ.get("PG_SPMM_PF", "0"), args.l2_fetch), end="")
print("hints=%s morton=%d m=%d nnz=%d assemble %.3f ms:" % (os.environ.get("PG_SPMV_HINTS", "1"), args.morton, args.m,
                                                            plan.nnz, asm_ms),
      " ".join("%s %.3f ms %.0f GB/s" % (k, v[0], v[1]) for k, v in out.items()))

# several right-hand sides per pass over the matrix: CSR (pg_spmm) and entity-blocked (pg_spmm_blocked)
from petgem_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

L = lib()
for k in (2, 4, 8):
    X = torch.randn((plan.N, k), dtype=torch.complex128, device=dev)
    Y1, Y2 = torch.empty_like(X), torch.empty_like(X)
    res = {}
    for name in ("csr", "blocked"):
        def run(Y):
            if name == "csr":
                check(L.pg_spmm(plan.N, ptr(rowptr), ptr(colidx), ptr(vals), k, ptr(X), None, ptr(Y), stream_ptr()))
            else:
                check(L.pg_spmm_blocked(plan._h, None, ptr(vals), k, ptr(X), None, ptr(Y), stream_ptr()))
        Y = Y1 if name == "csr" else Y2
        for _ in range(3):
            run(Y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run(Y)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / args.reps
    err = float((Y1 - Y2).abs().max() / Y1.abs().max())
    print("k=%d: csr %.3f ms (%.3f/rhs)  blocked %.3f ms (%.3f/rhs)  max rel diff %.1e"
          % (k, res["csr"], res["csr"] / k, res["blocked"], res["blocked"] / k, err))
    del X, Y1, Y2

#!/usr/bin/env python3
"""Time the SpMV variants (CSR / entity-blocked, cache-hint modes via PG_SPMV_HINTS) on a p=2 box."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0)
tab = bench.build_case(args.m, 2)
rows = bench.host_rows(tab)
el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"], rows["elemsF"],
                 rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
plan = AssemblyPlan(el, 2, order="locality")
plan.set_dirichlet(bench.bd_entities(tab, 2, plan.nEnt))
g, c = el.geometry()
vals = plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True)
rowptr, colidx = plan.csr()
x = torch.randn(plan.N, dtype=torch.complex128, device=dev)
out = {}
for name, A in (("csr", CSRMatrix(rowptr, colidx, vals, plan.N)), ("blocked", CSRMatrix(rowptr, colidx, vals, plan.N, plan=plan))):
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
print("hints=%s m=%d nnz=%d :" % (os.environ.get("PG_SPMV_HINTS", "0"), args.m, plan.nnz),
      " ".join("%s %.3f ms %.0f GB/s" % (k, v[0], v[1]) for k, v in out.items()))

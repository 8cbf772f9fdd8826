#!/bin/bash
python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData
dev = torch.device("cuda", 0)
tab = bench.build_case(94, 2); rows = bench.host_rows(tab)
el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"], rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
g, c = el.geometry()
plan = AssemblyPlan(el, 2, order="locality"); plan.set_dirichlet(bench.bd_entities(tab, 2, plan.nEnt))
vals = plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True)
rowptr, colidx = plan.csr()
A = CSRMatrix(rowptr, colidx, vals, plan.N, plan=plan)
for k in (4, 8):
    X = torch.randn((plan.N, k), dtype=torch.complex128, device=dev); Y = torch.empty_like(X)
    for _ in range(3): A.mult_multi(X, Y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): A.mult_multi(X, Y)
    e1.record(); torch.cuda.synchronize()
    print("k=%d: %.3f ms" % (k, e0.elapsed_time(e1) / 20))
    del X, Y
PY

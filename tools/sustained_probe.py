#!/usr/bin/env python3
"""Does a memory-bound kernel slow down when it runs for seconds instead of a 100 ms burst?  (The Krylov loop of the
C3 solve runs ~12 s; its per-iteration time is ~0.8 ms above the sum of its kernels timed one by one.)
Times the C3 SpMV in bursts of 20 while the loop keeps the GPU busy, with SM/memory clocks and power sampled."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData  # noqa: E402

dev = torch.device("cuda", 0)
tab = bench.build_case(94, 2)
rows = bench.host_rows(tab)
el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"], rows["elemsF"],
                 rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
g, c = el.geometry()
plan = AssemblyPlan(el, 2, order="locality")
plan.set_dirichlet(bench.bd_entities(tab, 2, plan.nEnt))
A = CSRMatrix(*plan.csr(), plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True), plan.N, plan=plan)
x = torch.ones(plan.N, dtype=torch.complex128, device=dev)
y = torch.empty_like(x)
for _ in range(5):
    A.mult(x, y)
torch.cuda.synchronize()
time.sleep(3.0)  # cool down
out = []
t_start = time.time()
while time.time() - t_start < 12.0:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        A.mult(x, y)
    e1.record()
    torch.cuda.synchronize()
    q = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu",
                        "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    out.append({"t_s": round(time.time() - t_start, 2), "spmv_ms": round(e0.elapsed_time(e1) / 20, 4), "sm_mem_power_temp": q})
print(json.dumps({"first": out[:3], "last": out[-3:], "n": len(out),
                  "min_ms": min(o["spmv_ms"] for o in out), "max_ms": max(o["spmv_ms"] for o in out)}))

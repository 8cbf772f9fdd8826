#!/usr/bin/env python3
"""Device-time breakdown of one Krylov iteration at C3 on N GPUs (nsys is not available; ncu must not wrap a
multi-rank run): every C-ABI call of the COCR loop is bracketed with CUDA events on the launching stream
(eager launches, PG_CUDA_GRAPH=0), aggregated per entry point over the timed iterations, max over ranks.
The waits (pg_comm_wait, pg_comm_allreduce) contain the time a rank spends waiting for its peers.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/iteration_timeline.py [--iters 60] [--pc jacobi]
One JSON object on stdout (rank 0)."""
import argparse
import collections
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PG_CUDA_GRAPH"] = "0"
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from petgem_b200 import _lib, device, gradient, krylov, peer  # noqa: E402
from petgem_b200.device import CSRMatrix, ElementData  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=94)
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--pc", default="jacobi,hiptmair")
ap.add_argument("--out", default="", help="also write the JSON object to this file (rank 0)")
args = ap.parse_args()
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

REAL = _lib.lib()
RECORD = []
ACTIVE = [False]


class _Proxy:
    """The ctypes library with every pg_* call bracketed by CUDA events while ACTIVE."""

    def __getattr__(self, name):
        fn = getattr(REAL, name)
        if not name.startswith("pg_"):
            return fn

        def timed(*a):
            if not ACTIVE[0]:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            RECORD.append((name, e0, e1))
            return rc
        return timed


proxy = _Proxy()
for mod in (krylov, peer, device, gradient):
    mod.lib = lambda: proxy

p = 2
tab = bench.build_case(args.m, p)
rows = bench.host_rows(tab)
el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"], rows["elemsF"],
                 rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
plan, row_begins = bench.make_plan(el, p, "locality", world, rank)
plan.set_dirichlet(bench.bd_entities(tab, p, plan.nEnt))
rowptr, colidx = plan.csr()
g, c = el.geometry(plan.element_range)
vals = plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True, diag=1.0)
A = CSRMatrix(rowptr, colidx, vals, plan.N, plan.row_begin, plan=plan)
ctx = krylov.DistContext(row_begins, plan.N) if world > 1 else None
b = bench.csem_rhs_device(tab, p, plan, dev)
out = {"n_gpus": world, "transport": ctx.transport if ctx else None, "tets": int(tab["elemsN"].shape[0]),
       "rows_per_rank": int(plan.local_rows), "iterations_timed": args.iters, "per_iteration_us": {}}
for pc in args.pc.split(","):
    op = krylov.Operator(A, pc=pc, ctx=ctx, halo="p2p" if world > 1 else "auto")
    krylov.cocr(op, b, rtol=1e-30, maxit=20)  # warm-up: allocations, IPC mappings
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    RECORD.clear()
    ACTIVE[0] = True
    t0 = time.time()
    krylov.cocr(op, b, rtol=1e-30, maxit=args.iters)
    torch.cuda.synchronize()
    wall = time.time() - t0
    ACTIVE[0] = False
    agg = collections.OrderedDict()
    for name, e0, e1 in RECORD:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1) * 1e3
    names = list(agg)
    t = torch.tensor([agg[k][1] for k in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_it = {k: {"calls_per_iteration": round(agg[k][0] / args.iters, 2), "us_per_iteration": round(float(v) / args.iters, 2)}
              for k, v in zip(names, t.tolist())}
    out["per_iteration_us"][pc] = {"entry_points": per_it,
                                   "sum_us": round(sum(v["us_per_iteration"] for v in per_it.values()), 1),
                                   "wall_us_per_iteration_eager": round(wall / args.iters * 1e6, 1)}
if rank == 0:
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(json.dumps(out) + "\n")
    os.write(1, (json.dumps(out) + "\n").encode())
sys.stderr.write("rank %d done\n" % rank)
sys.stderr.flush()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

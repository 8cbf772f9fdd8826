#!/usr/bin/env python3
"""Time-to-solution probe on the synthetic CSEM box (not the bench contract): assembly + Krylov solve
for a list of preconditioners.  usage: solve_case.py --m 94 --p 2 --pc hiptmair,jacobi [--ksp cr]
Under torchrun the rows are partitioned like bench.py does."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=94)
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--pc", default="hiptmair,jacobi")
    ap.add_argument("--ksp", default="cr")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--maxit", type=int, default=40000)
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--k", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from petgem_b200 import krylov
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tab = bench.build_case(args.m, args.p)
    rows = bench.host_rows(tab)
    el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"],
                     rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
    plan, row_begins = bench.make_plan(el, args.p, "locality", world, rank)
    plan.set_dirichlet(bench.bd_entities(tab, args.p, plan.nEnt))
    rowptr, colidx = plan.csr()
    g, c = el.geometry(plan.element_range)
    vals = plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True)
    A = CSRMatrix(rowptr, colidx, vals, plan.N, plan.row_begin, plan=plan)
    ctx = krylov.DistContext(row_begins, plan.N) if world > 1 else None
    b = bench.csem_rhs_device(tab, args.p, plan, dev)
    out = {"m": args.m, "p": args.p, "N": int(plan.N), "world": world, "ksp": args.ksp}
    for pc in args.pc.split(","):
        torch.cuda.synchronize()
        t0 = time.time()
        op = krylov.Operator(A, pc=pc, ctx=ctx, halo="p2p" if world > 1 else "auto")
        torch.cuda.synchronize()
        t_setup = time.time() - t0
        t0 = time.time()
        if args.ksp in ("cr", "cg"):
            B = b.reshape(-1, 1)
            if args.k > 1:
                B = torch.stack([bench.csem_rhs_device(tab, args.p, plan, dev, src=(1750.0 + dx, 1750.0, -975.0))
                                 for dx in np.linspace(-600, 600, args.k)], dim=1).contiguous()
            res = krylov.cocg_multi(op, B, rtol=args.rtol, maxit=args.maxit, max_seconds=args.seconds,
                                    method="cocr" if args.ksp == "cr" else "cocg")
            conv = bool(np.all(res.converged))
            rel = float(np.max(res.residuals[-1] / res.residuals[0]))
            x = res.x
            r = B - (op.matmat(x, torch.empty_like(x)) if B.shape[1] > 1 else op.matvec(x.reshape(-1), torch.empty_like(x).reshape(-1)).reshape(-1, 1))
            tr = torch.linalg.vector_norm(r, dim=0) ** 2
            bn = torch.linalg.vector_norm(B, dim=0) ** 2
            if world > 1:
                dist.all_reduce(tr)
                dist.all_reduce(bn)
            true_rel = float((tr / bn).sqrt().max())
        else:
            fn = {"gmres": krylov.gmres, "bcgs": krylov.bicgstab, "tfqmr": krylov.tfqmr}[args.ksp]
            res = fn(op, b, rtol=args.rtol, maxit=args.maxit)
            conv, rel, true_rel = bool(res.converged), res.residuals[-1] / res.residuals[0], None
        torch.cuda.synchronize()
        dt = time.time() - t0
        out[pc] = {"setup_s": t_setup, "iterations": res.iterations, "converged": conv, "reason": res.reason,
                   "seconds": dt, "ms_per_iteration": 1e3 * dt / max(res.iterations, 1), "rel_residual": rel,
                   "true_rel_residual": true_rel}
        del op, res
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

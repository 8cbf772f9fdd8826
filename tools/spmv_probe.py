#!/usr/bin/env python3
"""Where does the SpMV's DRAM traffic beyond the matrix stream come from?  (round 2)

Times the p = 2 entity-blocked MatMult at C3 (m = 94) on ONE GPU for the whole matrix and for the row block
rank `--rank` of `--world` would own (its x is [own | halo], 64 MB: smaller than the L2), under
  * the per-load eviction hints 0 / 1 / 2 (pg_tune_spmv_hints),
  * cudaLimitMaxL2FetchGranularity 32 / 64 / 128,
  * an access-policy window that keeps x in the persisting part of the L2 (pg_l2_persist),
  * the floor: every gather pointed at ONE x entry (no x traffic at all).
One JSON line per configuration on stdout."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from petgem_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=94)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=3)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--k", type=int, default=4, help="right-hand sides of the SpMM probes")
ap.add_argument("--only", default="both", choices=["both", "block", "whole"])
ap.add_argument("--quick", action="store_true", help="baseline MatMult only (for ncu captures)")
ap.add_argument("--quick-floor", action="store_true", help="with --quick: the x-free variant")
ap.add_argument("--quick-k", type=int, default=1, help="with --quick: right-hand sides (1 = MatMult)")
ap.add_argument("--persist", action="store_true", help="also the access-policy-window runs (measured slower)")
ap.add_argument("--fetch", action="store_true", help="also the L2 fetch granularity runs (measured neutral)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
L = lib()
tab = bench.build_case(args.m, 2)
rows = bench.host_rows(tab)
el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"], rows["elemsF"],
                 rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
g, c = el.geometry()
print(json.dumps({"persisting_l2_capacity_bytes": int(L.pg_l2_persist_capacity())}), flush=True)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def probe(label, world, rank):
    full = AssemblyPlan(el, 2, order="locality")
    N = full.N
    if world == 1:
        plan, lo, hi = full, 0, N
    else:
        cuts = [0] + [full.entity_aligned_row(N * r // world) for r in range(1, world)] + [N]
        lo, hi = cuts[rank], cuts[rank + 1]
        order = full.order_host
        del full
        torch.cuda.empty_cache()
        plan = AssemblyPlan(el, 2, order=order, row_range=(lo, hi))
    plan.set_dirichlet(bench.bd_entities(tab, 2, plan.nEnt))
    vals = plan.assemble(g, c, bench.OMEGA, bench.MU, apply_dirichlet=True)
    rowptr, colidx = plan.csr()
    n = hi - lo
    cs = plan.column_starts().to(torch.int64)
    if world > 1:  # [own | halo] numbering of the column entities, as krylov.DistContext.build_halo does
        cc = colidx.to(torch.int64)
        outside = (cc < lo) | (cc >= hi)
        ext = torch.unique(cc[outside])
        o2 = (cs < lo) | (cs >= hi)
        cs = torch.where(o2, n + torch.searchsorted(ext, cs), cs - lo)
        nx = n + int(ext.numel())
        del cc, outside
    else:
        nx = N
    cs32 = cs.to(torch.int32).contiguous()
    A = CSRMatrix(rowptr, colidx, vals, nx, lo, plan=plan, colstart=cs32, blocked=True)
    x = torch.randn(nx, dtype=torch.complex128, device=dev)
    y = torch.empty(n, dtype=torch.complex128, device=dev)
    nnz = plan.nnz
    alg = 20.0 * nnz + 40.0 * n
    base = {"case": label, "rows": n, "nnz": nnz, "x_bytes": nx * 16}

    def emit(cfg, ms, **kw):
        d = dict(base, config=cfg, ms=round(ms, 4), gbs_survey_formula=round(alg / ms / 1e6, 1) if ms > 0 else None)
        d.update(kw)
        print(json.dumps(d), flush=True)

    if args.quick:
        if args.quick_floor:
            A = CSRMatrix(rowptr, colidx, vals, nx, lo, plan=plan, colstart=(cs32 & 0x1FFE).contiguous(), blocked=True)
        if args.quick_k > 1:
            Xq = torch.randn((nx, args.quick_k), dtype=torch.complex128, device=dev)
            Yq = torch.empty((n, args.quick_k), dtype=torch.complex128, device=dev)
            emit("spmm k=%d" % args.quick_k, timed(lambda: A.mult_multi(Xq, Yq), 2))
            return
        emit("hints=1" + (" floor" if args.quick_floor else ""), timed(lambda: A.mult(x, y), 2))
        return
    # reference points: a pure read stream over the matrix values, and a copy of them
    wk = torch.empty((L.pg_reduce_workspace_bytes(1) // 16,), dtype=torch.complex128, device=dev)
    o1 = torch.zeros((1,), dtype=torch.complex128, device=dev)
    ms = timed(lambda: check(L.pg_dznrm2sq(nnz, ptr(vals), ptr(o1), ptr(wk), stream_ptr())), args.reps)
    emit("pure read stream over vals (pg_dznrm2sq)", ms, read_gbs=round(16.0 * nnz / ms / 1e6, 1))
    if world > 1:
        v2 = torch.empty_like(vals)
        ms = timed(lambda: v2.copy_(vals), args.reps)
        emit("copy of vals (torch)", ms, read_plus_write_gbs=round(32.0 * nnz / ms / 1e6, 1))
        del v2
    for mode in (1, 4, 1, 4, 0, 2, 3):
        check(L.pg_tune_spmv_hints(mode))
        emit("hints=%d" % mode, timed(lambda: A.mult(x, y), args.reps))
    for mode in (4,):
        check(L.pg_tune_spmv_hints(mode))
        for chunk in (2, 4):
            check(L.pg_tune_spmv_chunk(chunk))
            emit("hints=%d chunk=%d" % (mode, chunk), timed(lambda: A.mult(x, y), args.reps))
        check(L.pg_tune_spmv_chunk(1))
    check(L.pg_tune_spmv_hints(1))
    if args.fetch:
        for gran in (32, 128, 64):
            check(L.pg_l2_fetch_granularity(gran))
            emit("l2_fetch=%d" % gran, timed(lambda: A.mult(x, y), args.reps))
    side = torch.cuda.Stream()
    if args.persist:  # x kept in the persisting part of the L2 (side stream: the window is a stream attribute)
        with torch.cuda.stream(side):
            for mode in (1, 2):
                check(L.pg_tune_spmv_hints(mode))
                check(L.pg_l2_persist(ptr(x), nx * 16, 0.0, stream_ptr()))
                emit("persist window on x, hints=%d" % mode, timed(lambda: A.mult(x, y), args.reps))
                check(L.pg_l2_persist(None, 0, 0.0, stream_ptr()))
        check(L.pg_tune_spmv_hints(1))
    # floor: (almost) no x traffic: every gather lands in the first 128 KB of x (spread over 4096 sector pairs;
    # a single entry would serialise on one L2 slice: measured 0.91 ms against 0.64 ms for the real pattern)
    A0 = CSRMatrix(rowptr, colidx, vals, nx, lo, plan=plan, colstart=(cs32 & 0x1FFE).contiguous(), blocked=True)
    for mode in (1, 4):
        check(L.pg_tune_spmv_hints(mode))
        emit("floor (all gathers inside 128 KB of x), hints=%d" % mode, timed(lambda: A0.mult(x, y), args.reps))
    check(L.pg_tune_spmv_hints(1))
    # four right-hand sides
    k = args.k
    X = torch.randn((nx, k), dtype=torch.complex128, device=dev)
    Y = torch.empty((n, k), dtype=torch.complex128, device=dev)
    for pf, mode in ((0, 1), (1, 1), (1, 4), (1, 3)):
        check(L.pg_tune_spmm_prefetch(pf))
        check(L.pg_tune_spmv_hints(mode))
        emit("spmm k=%d prefetch=%d hints=%d" % (k, pf, mode), timed(lambda: A.mult_multi(X, Y), args.reps))
        if pf == 0:
            Yref = Y.clone()
    emit("spmm prefetch parity", 0.0, max_rel_diff=float((Y - Yref).abs().max() / Yref.abs().max()))
    emit("spmm k=%d floor (prefetch=1 hints=3)" % k, timed(lambda: A0.mult_multi(X, Y), args.reps))
    check(L.pg_tune_spmv_hints(1))


if args.only in ("both", "block"):
    probe("row block %d of %d" % (args.rank, args.world), args.world, args.rank)
    torch.cuda.empty_cache()
if args.only in ("both", "whole"):
    probe("whole matrix", 1, 0)

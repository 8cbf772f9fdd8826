#!/usr/bin/env python3
"""Benchmark of the PETGEM hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--m 94] [--p 2] [--extras c4,c5]

Workload (BASELINE.json configs[2], the configuration the metric "at p=2" is quoted on;
it fits one B200): synthetic layered-earth CSEM box, m=94 -> 4 983 504 tets, p=2,
31.8 M dofs, 1.4e9 nnz.  A *step* is one numeric assembly pass over the whole mesh:
per-element geometry + orientation kernel, then the fused element-matrix +
deterministic row-gather kernel with the Dirichlet condition applied (everything
Solver.assembly + zeroRowsColumns do to A).  Inputs are resident in HBM for `value`;
`e2e` repeats the step through the public API from pinned host arrays (H2D of the
per-element rows, D2H of ||A||_F^2 reduced on the device) inside the timed region.

Also reported (same JSON line): SpMV GB/s, Krylov iteration times and the solve to a TRUE
relative residual of 1e-8 (`solve`), time-to-solution on configs[1] (`tts`), the roofline of the
dominant kernel, the CPU baselines (oracle port on the host cores, bounded samples: assembly,
SpMV, GMRES time-to-solution), a `parity` block (multi-GPU runs: the NCCL path checked against
a one-GPU evaluation on rank 0) and the two other named configurations as `c4` (m=69, p=3, MT,
two polarizations) and `c5` (m=32, p=6), each with its own roofline.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FREQ = 2.0
OMEGA, MU = 2.0 * np.pi * FREQ, 4e-7 * np.pi
METRIC = "elements assembled/s (p=%d fused element-matrix + CSR assembly)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                mx = float(f[1])
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        return out


def build_case(m, p, vti=1.0):
    from petgem_b200 import synthetic

    t0 = time.time()
    nodes, elemsN = synthetic.kuhn_box(m)
    tab = synthetic.mesh_tables(nodes, elemsN)
    tab["sigma"] = synthetic.layered_sigma(nodes, elemsN, vti_ratio=vti)
    tab["host_prep_s"] = time.time() - t0
    return tab


def host_rows(tab):
    """Per-element scratch rows (nodes.dat, meshConnectivity.dat, edges.dat, edgesNodes.dat,
    facesEdges.dat, faces.dat, conductivityModel.dat) as contiguous host arrays."""
    elemsN, T = tab["elemsN"], tab["elemsN"].shape[0]
    return dict(
        nodes=np.ascontiguousarray(tab["nodes"][elemsN].reshape(T, 12)),
        elemsN=elemsN.astype(np.int32),
        elemsE=tab["elemsE"].astype(np.int32),
        edgesNodes=np.ascontiguousarray(tab["edgesNodes"][tab["elemsE"]].reshape(T, 12).astype(np.int32)),
        facesEdges=np.ascontiguousarray(tab["facesE"][tab["elemsF"]].reshape(T, 12).astype(np.int32)),
        elemsF=tab["elemsF"].astype(np.int32),
        sigma=np.ascontiguousarray(tab["sigma"]),
    )


def bd_entities(tab, p, nEnt):
    bd = np.zeros(nEnt, dtype=np.uint8)
    bd[tab["bEdges"]] = 1
    if p >= 2:
        bd[tab["nEdges"] + tab["bFaces"]] = 1
    return bd


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import petgem_oracle as oracle

    p, chunk = args
    out = 0.0
    for (coords, nodesEle, edgesEle, en, fe, sig) in chunk:
        Ae = oracle.element_system(coords, nodesEle, edgesEle, en, fe, sig, p, OMEGA, MU)
        out += abs(Ae[0, 0])
    return out


def cpu_sample(tab, p, nsample, seed=0):
    rng = np.random.default_rng(seed)
    T = tab["elemsN"].shape[0]
    sel = np.sort(rng.choice(T, size=min(nsample, T), replace=False))
    items = []
    for t in sel:
        items.append((tab["nodes"][tab["elemsN"][t]], tab["elemsN"][t], tab["elemsE"][t],
                      tab["edgesNodes"][tab["elemsE"][t]], tab["facesE"][tab["elemsF"][t]], tab["sigma"][t]))
    return sel, items


def cpu_assembly_rate(tab, p, nsample, pool, cores):
    """elements/s of the oracle port: element systems (process pool over the host cores) +
    the scipy-style scatter-add of the sampled cliques (single process, like PETSc's per-rank
    MatSetValues)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import petgem_oracle as oracle

    sel, items = cpu_sample(tab, p, nsample)
    chunks = [(p, items[i::cores]) for i in range(cores)]
    t0 = time.time()
    pool.map(_cpu_worker, chunks)
    t_elem = time.time() - t0
    n = p * (p + 2) * (p + 3) // 2
    dofs, *_, N = oracle.compute_connectivity_dofs(tab["elemsE"][sel], tab["elemsF"][sel], p)
    Ae = np.zeros((sel.size, n, n), dtype=np.complex128)
    t0 = time.time()
    oracle.assemble_global(Ae, dofs, N)
    t_asm = time.time() - t0
    return sel.size / (t_elem + t_asm), t_elem, t_asm


def run_reference_arm(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the reference
    itself is pure Python and cannot travel) on all host cores, same metric/config."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = args.p
    cores = os.cpu_count() or 1
    tab = build_case(args.m, p)  # the bench's own mesh (C3: m = 94); each step assembles a random sample of its elements
    nsample = args.cpu_sample or {1: 20000, 2: 6000, 3: 1500, 4: 400, 5: 120, 6: 60}[p]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_assembly_rate(tab, p, max(nsample // 8, cores), pool, cores)
        t0 = time.time()
        done = 0
        for _ in range(args.steps):
            cpu_assembly_rate(tab, p, nsample, pool, cores)
            done += nsample
        dt = time.time() - t0
    val = done / dt
    line = {
        "impl": "reference", "metric": METRIC % p, "value": val, "unit": "elements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic layered-earth CSEM box, m=%d -> %d tets, p=%d%s; each step assembles a "
                               "bounded random sample of its elements"
                               % (args.m, tab["elemsN"].shape[0], p, " (BASELINE configs[2])" if args.m == 94 else ""),
                   "tets": int(tab["elemsN"].shape[0]), "p": p},
        "cpu_baseline": {"value": val, "unit": "elements/s", "cores": cores, "kind": "port",
                         "sample": "%d elements of the m=%d mesh per step: oracle element_system in a %d-process pool + "
                                   "scatter-add of their cliques" % (nsample, args.m, cores),
                         # context: the unmodified reference (pure Python, cannot travel to the GPU box) needs
                         # 43 ms per element and core at p = 2 (SURVEY probe of computeElementalMatrices), ~4x this port
                         "reference_python_ms_per_element_p2": 43.0},
        "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# time-to-solution per CSEM source on BASELINE configs[1] (~1 M tets, p=1, single source, 1 GPU)
# ---------------------------------------------------------------------------------------------
def csem_rhs_device(tab, p, plan, dev, src=(1750.0, 1750.0, -975.0)):
    """b for a unit x-directed dipole at the centroid of the element nearest to `src` (default: the
    box centre) (solver.py:247-316), in the plan's numbering, owned rows only; Dirichlet rows zeroed."""
    import torch

    from petgem_b200 import hvfem

    src = np.asarray(src, dtype=np.float64)
    cen = tab["nodes"][tab["elemsN"]].mean(axis=1)
    te = int(np.argmin(((cen - src) ** 2).sum(axis=1)))
    Xe = tab["nodes"][tab["elemsN"][te]]
    J, Ji = hvfem.computeJacobian(Xe)
    eo, fo = hvfem.computeElementOrientation(tab["elemsE"][te], tab["elemsN"][te],
                                             tab["edgesNodes"][tab["elemsE"][te]], tab["facesE"][tab["elemsF"][te]])
    basis_, _ = hvfem.computeBasisFunctions(eo, fo, J, Ji, p, np.array([0.25, 0.25, 0.25]))
    de = hvfem.dofs_of_elements(tab["elemsE"][te], tab["elemsF"][te], [te], tab["nEdges"], tab["nFaces"], p)[0]
    rhs = 1j * OMEGA * MU * (np.array([1.0, 0.0, 0.0]) @ basis_[:, :, 0])
    perm = plan.dof_permutation()
    b = torch.zeros((plan.local_rows,), dtype=torch.complex128, device=dev)
    gi = perm[torch.as_tensor(de, device=dev)].to(torch.int64) - plan.row_begin
    ok = (gi >= 0) & (gi < plan.local_rows)
    b[gi[ok]] = torch.as_tensor(rhs, device=dev)[ok]
    return b


def time_to_solution(dev, m=55, p=1, maxit=20000):
    """Assembly + Krylov solve to rtol 1e-8 (examples/case1/petsc.opts: gmres, rtol 1e-8; Jacobi or the
    Hiptmair preconditioner instead of SOR) for one source, wall clock with a device synchronize on both
    sides.  Convergence is tested on the preconditioned residual like PETSc does unless the entry says
    "true residual" (-ksp_norm_type unpreconditioned)."""
    import torch

    from petgem_b200 import krylov
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    tab = build_case(m, p)
    rows = host_rows(tab)
    el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"],
                     rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    plan = AssemblyPlan(el, p, order="locality")
    plan.set_dirichlet(bd_entities(tab, p, plan.nEnt))
    rowptr, colidx = plan.csr()
    torch.cuda.synchronize()
    t_sym = time.time() - t0
    t0 = time.time()
    g, c = el.geometry()
    vals = plan.assemble(g, c, OMEGA, MU, apply_dirichlet=True, diag=1.0)
    torch.cuda.synchronize()
    t_asm = time.time() - t0
    A = CSRMatrix(rowptr, colidx, vals, plan.N, plan=plan)
    b = csem_rhs_device(tab, p, plan, dev)
    out = {"config": "synthetic layered-earth CSEM box, m=%d -> %d tets, p=%d, %d dofs (BASELINE configs[1])"
                     % (m, tab["elemsN"].shape[0], p, plan.N),
           "symbolic_s": t_sym, "assembly_s": t_asm}
    for name, opts in (("gmres(30)+jacobi", {"ksp_type": "gmres", "pc_type": "jacobi"}),
                       ("gmres(30)+hiptmair", {"ksp_type": "gmres", "pc_type": "hiptmair"}),
                       ("bcgs+hiptmair", {"ksp_type": "bcgs", "pc_type": "hiptmair"}),
                       ("cocr+jacobi", {"ksp_type": "cr", "pc_type": "jacobi"}),
                       ("cocr+hiptmair", {"ksp_type": "cr", "pc_type": "hiptmair"}),
                       ("cocr+hiptmair, true residual 1e-8", {"ksp_type": "cr", "pc_type": "hiptmair",
                                                             "ksp_norm_type": "unpreconditioned"})):
        opts.update({"ksp_rtol": 1e-8, "ksp_max_it": maxit})
        torch.cuda.synchronize()
        t0 = time.time()
        res = krylov.solve(A, b, opts)
        torch.cuda.synchronize()
        dt = time.time() - t0
        r = b - A.mult(res.x)
        out[name] = {"seconds": dt, "iterations": res.iterations, "converged": bool(res.converged),
                     "true_rel_residual": float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(b)),
                     "time_to_solution_s": t_asm + dt}
    # four sources of a tow line sharing A, solved in lockstep: one pass over the matrix per iteration
    B = torch.stack([csem_rhs_device(tab, p, plan, dev, src=(1750.0 + dx, 1750.0, -975.0))
                     for dx in (-600.0, -200.0, 200.0, 600.0)], dim=1).contiguous()
    torch.cuda.synchronize()
    t0 = time.time()
    X, results = krylov.solve_multi(A, B, {"ksp_type": "cr", "pc_type": "hiptmair", "ksp_rtol": 1e-8,
                                           "ksp_max_it": maxit, "ksp_norm_type": "unpreconditioned"})
    torch.cuda.synchronize()
    dt = time.time() - t0
    Rm = B - A.mult_multi(X)
    out["cocr+hiptmair, 4 sources in lockstep, true residual 1e-8"] = {
        "seconds": dt, "seconds_per_source": dt / 4, "iterations": results[0].iterations,
        "converged": bool(results[0].converged.all()),
        "true_rel_residual_max": float((torch.linalg.vector_norm(Rm, dim=0) / torch.linalg.vector_norm(B, dim=0)).max()),
        "time_to_solution_per_source_s": (t_asm + dt) / 4}
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def make_plan(el, p, order, world, rank):
    """Symbolic plan of this rank: the whole matrix on one GPU, else a PETSc-style contiguous row block
    (entity aligned) -> (plan, row_begins)."""
    import torch

    from petgem_b200.device import AssemblyPlan

    if world == 1:
        return AssemblyPlan(el, p, order=order), [0]
    probe = AssemblyPlan(el, p, order=order)  # global plan: entity-aligned PETSc-style split
    N = probe.N
    row_begins = [0] + [probe.entity_aligned_row(N * r // world) for r in range(1, world)]
    order_host = probe.order_host
    del probe
    torch.cuda.empty_cache()
    rb = row_begins[rank]
    re_ = row_begins[rank + 1] if rank + 1 < world else N
    plan = AssemblyPlan(el, p, order=order_host if order_host is not None else "reference", row_range=(rb, re_))
    return plan, row_begins


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def fixed_vector(lo, hi, dev, a=0.37, b=0.11):
    """Partition-independent test vector: entry i (global row in the numbering in use) is
    cos(a i) + i sin(b i)."""
    import torch

    i = torch.arange(lo, hi, dtype=torch.float64, device=dev)
    return torch.complex(torch.cos(a * i), torch.sin(b * i))


def checksums(A, op_j, op_h, plan, dev, dist=None):
    """Partition-independent checksums of the assembled block and of the operators built on it:
    ||A||_F^2, w^H (A x), w^H (M_hiptmair^-1 x) with fixed x, w (all-reduced over the ranks)."""
    import torch

    lo, hi = plan.row_begin, plan.row_begin + plan.local_rows
    x, w = fixed_vector(lo, hi, dev), fixed_vector(lo, hi, dev, 0.05, 0.23)
    y = op_j.matvec(x, torch.empty_like(x))
    out = [torch.sum(A.vals.real ** 2) + torch.sum(A.vals.imag ** 2) + 0j, torch.sum(torch.conj(w) * y)]
    if op_h is not None:
        z = op_h.precond(x, torch.empty_like(x))
        out.append(torch.sum(torch.conj(w) * z))
    t = torch.stack([torch.as_tensor(v, dtype=torch.complex128, device=dev) for v in out])
    if dist is not None:
        dist.all_reduce(torch.view_as_real(t))
    return [complex(v) for v in t.cpu().numpy()]


def parity_block(el, tab, p, order, dev, rank, world, dist, dist_sums):
    """Multi-GPU parity: rank 0 re-evaluates the whole problem on its own GPU (one plan over all rows, no
    NCCL) and compares with the all-reduced checksums of the row-partitioned NCCL path."""
    import torch

    from petgem_b200 import krylov
    from petgem_b200.device import AssemblyPlan, CSRMatrix

    res = None
    if rank == 0:
        plan1 = AssemblyPlan(el, p, order=order)
        plan1.set_dirichlet(bd_entities(tab, p, plan1.nEnt))
        g, c = el.geometry()
        vals1 = plan1.assemble(g, c, OMEGA, MU, apply_dirichlet=True, diag=1.0)
        A1 = CSRMatrix(*plan1.csr(), vals1, plan1.N, plan=plan1)
        ref = checksums(A1, krylov.Operator(A1, pc="jacobi"), krylov.Operator(A1, pc="hiptmair"), plan1, dev)
        names = ["frobenius_norm_sq", "spmv_checksum", "hiptmair_checksum"]
        rel = {n: abs(a - b) / abs(b) for n, a, b in zip(names, dist_sums, ref)}
        res = {"against": "one-GPU evaluation of the same problem on rank 0 (no NCCL)", "ranks": world,
               "relative_difference": rel, "tolerance": 1e-11, "pass": bool(max(rel.values()) <= 1e-11)}
        del A1, vals1, plan1
    torch.cuda.empty_cache()
    dist.barrier()
    return res


MT_LAYERS = ((0.0, 2.0 / 7.0, 1e-10), (2.0 / 7.0, 1.0, 0.01))  # air over earth (examples/case4/params.yaml:10-11)


def roofline_of(plan_stats, asm_ms, p, peak, peak_src, traffic=None):
    """Assembly roofline: 16 B per CSR value written + 4 n^2 B of slot map + (96+16+4n+4) B of
    element inputs per element visit; plan_stats = (nnz, contributions) of the rows assembled in asm_ms."""
    n = p * (p + 2) * (p + 3) // 2
    nnz, contributions = plan_stats
    visits = contributions / float(n * n)
    alg_bytes = 16.0 * nnz + visits * (4.0 * n * n + 96 + 16 + 4 * n + 4)
    achieved = alg_bytes / (asm_ms * 1e-3) / 1e9
    kname = ("assemble_small_kernel<%d>" if p <= 2 else "assemble_kernel<%d>") % p
    out = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": asm_ms}
    if p >= 3:
        # What actually bounds the p >= 3 kernel: every contribution contracts one table entry of 12 numbers read
        # from the L2-resident reference-element table (int32 numerators, 48 B; fp64 at p = 6, 96 B) -- 3 to 7
        # bytes of L2 traffic per byte of CSR written.  The L2 -> SM cap is ~6300 B/clk for the whole chip
        # (B300_MICROARCH.md, LTS throughput; 9.9 TB/s at the ~1.57 GHz these kernels run at, 12.4 TB/s at 1.965).
        tb = contributions * (96.0 if p == 6 else 48.0)
        out["l2_table"] = {"bytes_per_launch": tb, "achieved_tbs": tb / (asm_ms * 1e-3) / 1e12,
                           "cap_tbs_at_1570_mhz": 6300 * 1.57e9 / 1e12,
                           "frac_of_cap_at_1570_mhz": tb / (asm_ms * 1e-3) / (6300 * 1.57e9),
                           "note": "L2 -> SM table traffic (bytes per contribution x contributions) / kernel time"}
    return out


def run_c4(args, dev, world, rank, dist, peak, peak_src):
    """BASELINE configs[3]: synthetic 3-D MT model, m=69 -> 1 971 054 tets, p=3, two polarizations sharing A
    (no Dirichlet rows in MT mode, solver.py:552): assembly rate + roofline, two-vector SpMM, and a bounded
    lockstep COCR solve with the Hiptmair preconditioner."""
    import torch

    from petgem_b200 import krylov, mt, synthetic
    from petgem_b200 import mesh as pmesh
    from petgem_b200.device import CSRMatrix, ElementData
    from petgem_b200.preprocessing import boundary_element_rows

    p, m = 3, args.c4_m
    t0 = time.time()
    nodes, elemsN = synthetic.kuhn_box(m)
    tab = synthetic.mesh_tables(nodes, elemsN)
    tab["sigma"] = synthetic.layered_sigma(nodes, elemsN, layers=MT_LAYERS)
    host_s = time.time() - t0
    T = elemsN.shape[0]
    rows = host_rows(tab)
    el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"],
                     rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
    omega = OMEGA  # 2 Hz like examples/case4/params.yaml
    t0 = time.time()
    plan, row_begins = make_plan(el, p, "locality", world, rank)
    rowptr, colidx = plan.csr()
    torch.cuda.synchronize()
    symbolic_s = time.time() - t0
    vals = torch.empty((plan.nnz,), dtype=torch.complex128, device=dev)
    gbuf = el.geometry(plan.element_range)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def total(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            return float(t.item())
        return float(v)

    steps = max(3, min(args.steps, 5))
    for _ in range(3):
        g, c = el.geometry(plan.element_range, out=gbuf)
        plan.assemble(g, c, omega, MU, out=vals)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ev[0].record()
    for i in range(steps):
        g, c = el.geometry(plan.element_range, out=gbuf)
        kev[i][0].record()
        plan.assemble(g, c, omega, MU, out=vals)
        kev[i][1].record()
    ev[1].record()
    barrier()
    total_ms = maxr(ev[0].elapsed_time(ev[1])) / steps
    asm_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    out = {"config": "synthetic 3-D MT box (air over 0.01 S/m earth), m=%d -> %d tets, p=3, N=%d dofs, two "
                     "polarizations (BASELINE configs[3])" % (m, T, plan.N),
           "tets": T, "dofs": int(plan.N), "nnz": int(total(plan.nnz)), "n_gpus": world,
           "elements_per_s": T / (total_ms * 1e-3), "ms_per_step": total_ms, "steps": steps,
           "roofline": roofline_of((plan.nnz, plan.contributions), asm_ms, p, peak, peak_src),
           "setup": {"host_mesh_s": host_s, "symbolic_s": symbolic_s}}
    # MatMult for the two polarizations at once (CSR SpMM, k = 2): bytes = 20 nnz + 2 * 40 rows
    A = CSRMatrix(rowptr, colidx, vals, plan.N, plan.row_begin, plan=plan)
    ctx = krylov.DistContext(row_begins, plan.N) if world > 1 else None
    op = krylov.Operator(A, pc="hiptmair", ctx=ctx, halo="p2p" if world > 1 else "auto")
    X2 = torch.ones((plan.local_rows, 2), dtype=torch.complex128, device=dev)
    Y2 = torch.empty_like(X2)
    for _ in range(3):
        op.matmat(X2, Y2)
    barrier()
    ev[0].record()
    for _ in range(steps):
        op.matmat(X2, Y2)
    ev[1].record()
    barrier()
    mm_ms = maxr(ev[0].elapsed_time(ev[1])) / steps
    mm_bytes = 20.0 * out["nnz"] + 2 * 40.0 * plan.N
    out["spmm_two_polarizations"] = {"ms": mm_ms, "ms_per_rhs": mm_ms / 2, "gbs": mm_bytes / (mm_ms * 1e-3) / 1e9,
                                     "frac_of_peak": mm_bytes / (mm_ms * 1e-3) / 1e9 / (peak * world),
                                     "kernel": "spmm_kernel<2> (CSR)", "includes_halo_exchange": world > 1}
    del X2, Y2
    if not args.no_solve:
        # right-hand sides of the two polarizations from the 1-D layered-earth excitation (solver.py:318-512)
        t0 = time.time()
        bFacesN, bFaces, _ = pmesh.computeBoundaryFaces(tab["elemsF"], tab["facesN"])
        plane = pmesh.computeFacePlane(tab["nodes"], bFaces, bFacesN)
        bElems, _ = pmesh.computeBoundaryElements(tab["elemsF"], bFaces, tab["nFaces"])
        from petgem_b200 import hvfem
        dofs_b = hvfem.dofs_of_elements(tab["elemsE"][bElems], tab["elemsF"][bElems], bElems, tab["nEdges"],
                                        tab["nFaces"], p)
        brow = boundary_rows_sparse(tab, dofs_b, bFaces, bElems, plane)
        z = tab["nodes"][:, 2]
        rhs = mt.mt_rhs(brow, float(z.max()), float(z.min()), p, omega, MU, ["x", "y"], plan.N, n_nodes_1d=200001)
        perm = plan.dof_permutation().to(torch.int64)
        lo, hi = plan.row_begin, plan.row_begin + plan.local_rows
        B = torch.zeros((plan.N, 2), dtype=torch.complex128, device=dev)
        for i in range(2):
            B[perm, i] = torch.as_tensor(rhs[i], device=dev)
        B = B[lo:hi].contiguous()
        rhs_s = time.time() - t0
        barrier()
        t0 = time.time()
        res = krylov.cocg_multi(op, B, rtol=1e-8, maxit=args.c4_maxit, max_seconds=args.c4_seconds, method="cocr")
        barrier()
        dt = time.time() - t0
        out["solve"] = {"method": "cocr+hiptmair, two polarizations in lockstep", "iterations": res.iterations,
                        "converged": bool(np.all(res.converged)), "reason": res.reason, "seconds": dt,
                        "ms_per_iteration": 1e3 * dt / max(res.iterations, 1),
                        "rel_residual": [float(v) for v in res.residuals[-1] / np.maximum(res.residuals[0], 1e-300)],
                        "host_rhs_s": rhs_s}
    return out


def boundary_rows_sparse(tab, dofs_b, bFaces, bElems, plane):
    """Rows of boundaryElements.dat (preprocessing.py:326-367) for the boundary elements only."""
    t = np.asarray(bElems, dtype=np.int64)
    nb, n = t.size, dofs_b.shape[1]
    rows = np.zeros((nb, 53 + n), dtype=np.float64)
    rows[:, 0:4] = tab["elemsN"][t]
    rows[:, 4:16] = tab["nodes"][tab["elemsN"][t]].reshape(nb, 12)
    rows[:, 16:20] = tab["elemsF"][t]
    rows[:, 20:32] = tab["facesE"][tab["elemsF"][t]].reshape(nb, 12)
    rows[:, 32:38] = tab["elemsE"][t]
    rows[:, 38:50] = tab["edgesNodes"][tab["elemsE"][t]].reshape(nb, 12)
    rows[:, 50] = plane
    rows[:, 51] = bFaces
    rows[:, 52] = tab["sigma"][t, 0]
    rows[:, 53:] = dofs_b
    return rows


def run_c5(args, dev, world, rank, dist, peak, peak_src):
    """BASELINE configs[4]: high-order stress test, m=32 -> 196 608 tets, p=6 (216 dofs per element,
    ~8e9 nonzeros = 128 GB of CSR values): assembly dominated.  The matrix does not fit one GPU next to the
    plan, so the rows are assembled block by block into one reused buffer (PETSc-style contiguous row blocks,
    the same owner-computes partition the multi-GPU run uses): `world * nsub` blocks, `nsub` per rank."""
    import torch

    from petgem_b200.device import AssemblyPlan, ElementData

    p, m = 6, args.c5_m
    t0 = time.time()
    tab = build_case(m, p)
    T = tab["elemsN"].shape[0]
    rows = host_rows(tab)
    el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"],
                     rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
    nsub = max(1, 8 // world)
    nblocks = world * nsub
    t0 = time.time()
    probe = AssemblyPlan(el, p, order="locality")
    N = probe.N
    cuts = [0] + [probe.entity_aligned_row(N * r // nblocks) for r in range(1, nblocks)] + [N]
    order_host = probe.order_host
    del probe
    torch.cuda.empty_cache()
    bd = bd_entities(tab, p, tab["nEdges"] + tab["nFaces"] + T)
    geo = el.geometry()
    steps = max(2, min(args.steps, 3))
    ms_local, nnz_local, contrib_local = 0.0, 0, 0
    buf = None
    for blk in range(rank * nsub, (rank + 1) * nsub):
        plan = AssemblyPlan(el, p, order=order_host, row_range=(cuts[blk], cuts[blk + 1]))
        plan.set_dirichlet(bd)
        if buf is None or buf.numel() < plan.nnz:
            buf = None
            torch.cuda.empty_cache()
            buf = torch.empty((int(plan.nnz * 1.05),), dtype=torch.complex128, device=dev)
        vals = buf[: plan.nnz]
        plan.assemble(geo[0], geo[1], OMEGA, MU, apply_dirichlet=True, out=vals)  # warm-up (+ table build)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            plan.assemble(geo[0], geo[1], OMEGA, MU, apply_dirichlet=True, out=vals)
        b.record()
        torch.cuda.synchronize()
        ms_local += a.elapsed_time(b) / steps
        nnz_local += plan.nnz
        contrib_local += plan.contributions
        del plan
    symbolic_and_run_s = time.time() - t0
    t = torch.tensor([ms_local, float(nnz_local), float(contrib_local)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
    ms = float(tmax[0])
    roof = roofline_of((float(t[1]) / world, float(t[2]) / world), ms, p, peak, peak_src)
    return {"config": "synthetic layered-earth CSEM box, m=%d -> %d tets, p=6, N=%d dofs (BASELINE configs[4])"
                      % (m, T, N), "tets": T, "dofs": int(N), "nnz": int(t[1]), "n_gpus": world,
            "row_blocks": nblocks, "blocks_per_gpu": nsub, "elements_per_s": T / (ms * 1e-3), "ms_per_pass": ms,
            "steps": steps, "roofline": roof, "wall_s_including_symbolic": symbolic_and_run_s,
            "mode": "assemble-and-discard: every row block is assembled into one reused buffer"}


# ---------------------------------------------------------------------------------------------
# CPU baselines beside the GPU numbers (rank 0, N=1): SpMV and GMRES time-to-solution
# ---------------------------------------------------------------------------------------------
_CPU_SPMV = {}


def _cpu_spmv_worker(args):
    c, reps = args
    A, x = _CPU_SPMV["chunks"][c], _CPU_SPMV["x"]
    acc = 0.0
    for _ in range(reps):
        acc += float(abs((A @ x)[0]))
    return acc


def cpu_spmv_baseline(A, cores, rows_target=1500000, reps=4):
    """scipy complex128 CSR MatMult on the host cores: a slab of the first rows of the GPU-assembled matrix,
    row blocks over one process per core, same byte formula as the GPU number (20 nnz + 40 rows)."""
    import multiprocessing as mp

    import scipy.sparse as sp

    R = int(min(A.rows, rows_target))
    rp = A.rowptr[: R + 1].cpu().numpy()
    nnz = int(rp[-1])
    ci = A.colidx[:nnz].cpu().numpy()
    vv = A.vals[:nnz].cpu().numpy()
    ncols = int(ci.max()) + 1
    cuts = np.linspace(0, R, cores + 1).astype(np.int64)
    chunks = []
    for c in range(cores):
        a, b = cuts[c], cuts[c + 1]
        chunks.append(sp.csr_matrix((vv[rp[a]:rp[b]], ci[rp[a]:rp[b]], rp[a:b + 1] - rp[a]), shape=(b - a, ncols)))
    _CPU_SPMV["chunks"] = chunks
    _CPU_SPMV["x"] = np.ones(ncols, dtype=np.complex128)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_spmv_worker, [(c, 1) for c in range(cores)])
        t0 = time.time()
        pool.map(_cpu_spmv_worker, [(c, reps) for c in range(cores)])
        dt = time.time() - t0
    _CPU_SPMV.clear()
    nbytes = (20.0 * nnz + 40.0 * R) * reps
    return {"gbs": nbytes / dt / 1e9, "cores": cores, "kind": "port",
            "sample": "scipy CSR A@x on the first %d rows (%d nnz) of the GPU-assembled matrix, row blocks over %d "
                      "processes, %d repetitions in %.2f s" % (R, nnz, cores, reps, dt)}


def cpu_gmres_tts(A, b, label, budget_s=20.0):
    """Reference-equivalent solver settings on the CPU: GMRES(30), left Jacobi, rtol 1e-8 (examples/case1/
    petsc.opts with Jacobi for SOR) through the oracle's KSPGMRES restatement, bounded by wall time."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import petgem_oracle as oracle

    As = A.to_scipy().tocsr()
    bh = b.cpu().numpy()
    dinv = 1.0 / As.diagonal()
    t0 = time.time()
    _, its0, _ = oracle.gmres(lambda v: As @ v, bh, rtol=1e-8, maxit=30, pc=lambda v: dinv * v)
    per_it = (time.time() - t0) / max(its0, 1)
    maxit = int(max(30, min(100000, budget_s / per_it)))
    t0 = time.time()
    x, its, hist = oracle.gmres(lambda v: As @ v, bh, rtol=1e-8, maxit=maxit, pc=lambda v: dinv * v)
    dt = time.time() - t0
    rel = hist[-1] / hist[0]
    return {"system": label, "solver": "gmres(30)+jacobi (oracle port of KSPGMRES), 1 core", "iterations": its,
            "seconds": dt, "converged": bool(rel <= 1e-8), "rel_residual": float(rel), "dofs": int(As.shape[0])}


def box_system(dev, m, p=1):
    """(A, b) of the synthetic CSEM box (Dirichlet applied), reference numbering, on the GPU."""
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    tab = build_case(m, p)
    rows = host_rows(tab)
    el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"],
                     rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
    plan = AssemblyPlan(el, p, order="reference")
    plan.set_dirichlet(bd_entities(tab, p, plan.nEnt))
    g, c = el.geometry()
    A = CSRMatrix(*plan.csr(), plan.assemble(g, c, OMEGA, MU, apply_dirichlet=True), plan.N, plan=plan)
    return A, csem_rhs_device(tab, p, plan, dev), tab["elemsN"].shape[0]


def c1_system(dev):
    """BASELINE configs[0] geometry: the reference's own test mesh (tests/data/test_mesh.msh, recorded in
    tests/golden/test_mesh_topology.npz) with the case1 physics at p=1 -> (A, b) on the GPU."""
    import torch

    from petgem_b200 import hvfem
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    topo = dict(np.load(os.path.join(ROOT, "tests", "golden", "test_mesh_topology.npz")))
    sig = np.array([1.0, 0.01, 1.0, 3.3333])[topo["tags"] - 1]
    el = ElementData.from_mesh(topo["nodes"], topo["elemsN"], topo["elemsE"], topo["edgesNodes"], topo["elemsF"],
                               topo["facesE"], np.stack([sig, sig], axis=1), device=dev)
    plan = AssemblyPlan(el, 1, order="reference")
    bd = np.zeros(plan.nEnt, dtype=np.uint8)
    bd[topo["bEdges"]] = 1
    plan.set_dirichlet(bd)
    g, c = el.geometry()
    A = CSRMatrix(*plan.csr(), plan.assemble(g, c, OMEGA, MU, apply_dirichlet=True), plan.N, plan=plan)
    tab = dict(nodes=topo["nodes"], elemsN=topo["elemsN"], elemsE=topo["elemsE"], edgesNodes=topo["edgesNodes"],
               elemsF=topo["elemsF"], facesE=topo["facesE"], nEdges=topo["edgesNodes"].shape[0],
               nFaces=topo["facesE"].shape[0])
    b = csem_rhs_device(tab, 1, plan, dev)
    b[torch.as_tensor(topo["boundary_dofs_p1"], device=dev)] = 0
    return A, b


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # C-level prints (e.g. "NCCL version ...") must not pollute the JSON line
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--m", type=int, default=94, help="hexes per box side (T = 6 m^3); 94 = C3")
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--solve-maxit", type=int, default=40000)
    ap.add_argument("--solve-seconds", type=float, default=120.0, help="wall-time bound of the C3 solve")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-tts", action="store_true")
    ap.add_argument("--tts-m", type=int, default=55, help="box size of the time-to-solution case (55 = C2)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--order", default="locality", choices=["locality", "reference"])
    ap.add_argument("--extras", default="c4,c5", help="other named configurations to measure (c4,c5 or none)")
    ap.add_argument("--c4-m", type=int, default=69)
    ap.add_argument("--c4-maxit", type=int, default=3000)
    ap.add_argument("--c4-seconds", type=float, default=25.0)
    ap.add_argument("--c5-m", type=int, default=32)
    ap.add_argument("--jacobi-seconds", type=float, default=12.0, help="bound of the Jacobi comparison solve")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from petgem_b200 import krylov
    from petgem_b200._lib import lib
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    p = args.p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- setup (untimed): mesh, symbolic phase ------------------------------------------------
    tab = build_case(args.m, p)
    T = tab["elemsN"].shape[0]
    rows = host_rows(tab)
    el = ElementData(rows["nodes"], rows["elemsN"], rows["elemsE"], rows["edgesNodes"], rows["facesEdges"],
                     rows["elemsF"], rows["sigma"], tab["nEdges"], tab["nFaces"], device=dev)
    t0 = time.time()
    plan, row_begins = make_plan(el, p, args.order, world, rank)
    plan.set_dirichlet(bd_entities(tab, p, plan.nEnt))
    rowptr, colidx = plan.csr()
    torch.cuda.synchronize()
    symbolic_s = time.time() - t0
    vals = torch.empty((plan.nnz,), dtype=torch.complex128, device=dev)
    erange = plan.element_range  # elements incident to the owned rows (all of them on one GPU)
    gbuf = el.geometry(erange)

    def step():
        g, c = el.geometry(erange, out=gbuf)
        plan.assemble(g, c, OMEGA, MU, apply_dirichlet=True, diag=1.0, out=vals)

    # ---- timed region: K assembly steps ----------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        g, c = el.geometry(erange, out=gbuf)
        kev[i][0].record()
        plan.assemble(g, c, OMEGA, MU, apply_dirichlet=True, diag=1.0, out=vals)
        kev[i][1].record()
    ev[1].record()
    barrier()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    asm_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    value = T * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel (assemble_kernel), algorithmic bytes per SURVEY 8(d):
    # 16 B per CSR value written + 4 n^2 B of slot map + (96+16+4n+4) B of element inputs
    n = plan.n
    t_local = plan.contributions / float(n * n)  # element visits of this rank (redundant ones included)
    alg_bytes = 16.0 * plan.nnz + t_local * (4.0 * n * n + 96 + 16 + 4 * n + 4)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (asm_ms * 1e-3) / 1e9
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "r2_traffic_c3.json")
    if p == 2 and args.m == 94 and world == 1 and os.path.exists(tpath):
        t = json.load(open(tpath))["assemble_small_kernel<2>"]
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    kname = ("assemble_small_kernel<%d>" if p <= 2 else "assemble_kernel<%d>") % p
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": asm_ms}

    # ---- SpMV (Dirichlet-applied A), inputs larger than L2 ----------------------------------------
    A = CSRMatrix(rowptr, colidx, vals, plan.N, plan.row_begin, plan=plan)  # p=2: entity-blocked MatMult
    ctx = krylov.DistContext(row_begins, plan.N) if world > 1 else None
    op = krylov.Operator(A, pc="jacobi", ctx=ctx)
    if op.xchg is not None:  # peer transport: x lives in front of its halo, as in the Krylov drivers
        xg = op.halo_vector(None, tag="bench")[: plan.local_rows]
        xg.fill_(1.0)
    else:
        xg = torch.ones((plan.local_rows,), dtype=torch.complex128, device=dev)
    yg = torch.empty_like(xg)
    for _ in range(3):
        op.matvec(xg, yg)
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        op.matvec(xg, yg)
    ev[1].record()
    barrier()
    spmv_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / args.steps
    nnz_total, rows_total = plan.nnz, plan.local_rows
    if world > 1:
        t = torch.tensor([plan.nnz, plan.local_rows], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        nnz_total, rows_total = int(t[0]), int(t[1])
    spmv_bytes = 20.0 * nnz_total + 40.0 * rows_total
    spmv = {"ms": spmv_ms, "gbs": spmv_bytes / (spmv_ms * 1e-3) / 1e9,
            "frac_of_peak": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / (peak * world), "nnz": nnz_total, "rows": rows_total,
            "includes_halo_exchange": world > 1,
            "transport": (ctx.transport if ctx is not None else None), "halo_mode": op.mode}
    # four right-hand sides per pass over the matrix (sources sharing A): bytes = 20 nnz + 4 * 40 rows
    if op.mode in ("single", "p2p"):
        if op.xchg is not None:
            Xm = op.halo_vector(4, tag="bench")[: plan.local_rows]
            Xm.fill_(1.0)
        else:
            Xm = torch.ones((plan.local_rows, 4), dtype=torch.complex128, device=dev)
        Ym = torch.empty_like(Xm)
        for _ in range(3):
            op.matmat(Xm, Ym)
        barrier()
        ev[0].record()
        for _ in range(args.steps):
            op.matmat(Xm, Ym)
        ev[1].record()
        barrier()
        mm_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / args.steps
        mm_bytes = 20.0 * nnz_total + 4 * 40.0 * rows_total
        spmv["four_rhs"] = {"ms": mm_ms, "ms_per_rhs": mm_ms / 4, "gbs": mm_bytes / (mm_ms * 1e-3) / 1e9,
                            "frac_of_peak": mm_bytes / (mm_ms * 1e-3) / 1e9 / (peak * world)}
        del Xm, Ym

    # ---- bounded Krylov run: time per GMRES iteration, time-to-solution if it converges -------------
    solve = None
    if not args.no_solve:
        b = csem_rhs_device(tab, p, plan, dev)
        barrier()
        t0 = time.time()
        res = krylov.gmres(op, b, rtol=1e-8, restart=30, maxit=min(args.solve_maxit, 90))  # 3 restart cycles
        barrier()
        dt = time.time() - t0
        solve = {"gmres(30)+jacobi": {"iterations": res.iterations, "ms_per_iteration": 1e3 * dt / max(res.iterations, 1),
                                      "rel_residual": res.residuals[-1] / res.residuals[0]}}
        # COCG (A is complex symmetric): per-iteration cost from a bounded run
        barrier()
        t0 = time.time()
        res = krylov.cocg(op, b, rtol=1e-8, maxit=min(args.solve_maxit, 300))
        barrier()
        dt = time.time() - t0
        solve["cocg+jacobi"] = {"iterations": res.iterations, "ms_per_iteration": 1e3 * dt / max(res.iterations, 1),
                                "rel_residual": res.residuals[-1] / res.residuals[0]}
        # the solve proper: COCR with the Hiptmair (gradient-space) preconditioner to a TRUE relative residual
        # ||b - A x|| <= 1e-8 ||b|| (-ksp_norm_type unpreconditioned), bounded by iterations and wall time
        barrier()
        t0 = time.time()
        op_h = krylov.Operator(A, pc="hiptmair", ctx=ctx, halo="p2p" if world > 1 else "auto")
        barrier()
        pc_setup_s = time.time() - t0
        t0 = time.time()
        res = krylov.cocr(op_h, b, rtol=1e-8, maxit=args.solve_maxit, max_seconds=args.solve_seconds,
                          norm_type="unpreconditioned")
        barrier()
        dt = time.time() - t0
        r = b - op.matvec(res.x, torch.empty_like(b))
        rr = torch.stack([torch.sum(r.real ** 2 + r.imag ** 2), torch.sum(b.real ** 2 + b.imag ** 2)])
        if world > 1:
            dist.all_reduce(rr)
        solve["cocr+hiptmair"] = {"rtol": 1e-8, "norm": "unpreconditioned (true residual)", "iterations": res.iterations,
                                  "converged": bool(res.converged), "reason": res.reason,
                                  "preconditioned_rel_residual": res.residuals[-1] / res.residuals[0],
                                  "true_rel_residual": float((rr[0] / rr[1]).sqrt()),
                                  "seconds": dt, "pc_setup_s": pc_setup_s,
                                  "ms_per_iteration": 1e3 * dt / max(res.iterations, 1),
                                  "time_to_solution_s": total_ms / args.steps * 1e-3 + pc_setup_s + dt}
        # multi-GPU parity of the NCCL path (assembly, halo SpMV, distributed preconditioner)
        if world > 1:
            sums = checksums(A, op, op_h, plan, dev, dist)
        del res
        # the round-1 solver for comparison (point Jacobi; 12 300 iterations to converge at C3): bounded run
        barrier()
        t0 = time.time()
        res = krylov.cocr(op, b, rtol=1e-8, maxit=args.solve_maxit, max_seconds=args.jacobi_seconds)
        barrier()
        dt = time.time() - t0
        solve["cocr+jacobi"] = {"rtol": 1e-8, "iterations": res.iterations, "converged": bool(res.converged),
                                "reason": res.reason, "rel_residual": res.residuals[-1] / res.residuals[0],
                                "seconds": dt, "ms_per_iteration": 1e3 * dt / max(res.iterations, 1)}
        if op.mode in ("single", "p2p"):
            # cost of an iteration with four sources in lockstep (bounded run, not a solve)
            del res
            B4 = torch.stack([csem_rhs_device(tab, p, plan, dev, src=(1750.0 + dx, 1750.0, -975.0))
                              for dx in (-600.0, -200.0, 200.0, 600.0)], dim=1).contiguous()
            barrier()
            t0 = time.time()
            resm = krylov.cocg_multi(op_h, B4, rtol=1e-8, maxit=min(args.solve_maxit, 200), method="cocr")
            barrier()
            dt = time.time() - t0
            solve["cocr+hiptmair, 4 sources in lockstep"] = {
                "iterations": resm.iterations, "ms_per_iteration": 1e3 * dt / max(resm.iterations, 1),
                "ms_per_iteration_per_source": 1e3 * dt / max(resm.iterations, 1) / 4}
            del resm, B4

    # ---- time-to-solution on configs[1] (1 GPU only) ---------------------------------------------------
    tts = None
    if world == 1 and not args.no_solve and not args.no_tts:
        tts = time_to_solution(dev, m=args.tts_m, p=1)

    # ---- e2e through the public API with host buffers -----------------------------------------------
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in rows.items()}
    # the step's result read back by the host: ||A||_F^2 of the assembled block (a 16-byte metric, like a
    # loss), reduced on the device by pg_dznrm2sq
    from petgem_b200._lib import check, ptr, stream_ptr
    fro_host = torch.empty((1,), dtype=torch.complex128).pin_memory()
    fro_dev = torch.zeros((1,), dtype=torch.complex128, device=dev)
    fro_work = torch.empty((lib().pg_reduce_workspace_bytes(1) // 16,), dtype=torch.complex128, device=dev)

    # each rank needs (and copies) only the rows of the elements incident to its matrix rows
    t0e, t1e = erange
    dev_rows = {"nodes": el.nodes, "elemsN": el.elemsN, "elemsE": el.elemsE, "edgesNodes": el.edgesNodes,
                "facesEdges": el.facesEdges, "elemsF": el.elemsF, "sigma": el.sigma}
    h2d = sum(int(t[t0e:t1e].numel() * t.element_size()) for t in pinned.values())

    # Inputs are double buffered: the host->device copy of step i+1 (copy stream) overlaps the kernels of
    # step i; every step still copies all of its inputs and reads back its own result.
    import copy as _copy
    sets = [dev_rows, {k: torch.empty_like(v) for k, v in dev_rows.items()}]
    els = [el, _copy.copy(el)]
    for k, v in sets[1].items():
        setattr(els[1], k, v)
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i, overlap=True, only=None):
        s = i & 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            if not overlap:
                copy_stream.wait_stream(cur)     # serial variant: the copy starts after the previous step
            copy_stream.wait_event(consumed[s])  # the geometry kernel of step i-2 has read this buffer set
            for k, t in pinned.items():
                if only is None or k in only:
                    sets[s][k][t0e:t1e].copy_(t[t0e:t1e], non_blocking=True)
            copied[s].record(copy_stream)
        cur.wait_event(copied[s])
        g, c = els[s].geometry(erange, out=gbuf)
        consumed[s].record(cur)
        plan.assemble(g, c, OMEGA, MU, apply_dirichlet=True, diag=1.0, out=vals)
        check(lib().pg_dznrm2sq(plan.nnz, ptr(vals), ptr(fro_dev), ptr(fro_work), stream_ptr()), "pg_dznrm2sq")
        fro_host.copy_(fro_dev, non_blocking=True)

    ksteps = max(4, args.steps // 2)
    e2e_serial_ms = None
    for overlap in (False, True):
        for i in range(2):
            e2e_step(i, overlap)
        barrier()
        ev[0].record()
        for i in range(ksteps):
            e2e_step(i, overlap)
        ev[1].record()
        barrier()
        ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / ksteps
        if overlap:
            e2e_ms = ms
        else:
            e2e_serial_ms = ms
    # the loop a survey / inversion actually runs on a fixed mesh: only the conductivity model changes between
    # steps (a new frequency changes a scalar), so only sigma travels; geometry + assembly + read-back as above
    for k, v in sets[1].items():
        if k != "sigma":
            v.copy_(dev_rows[k])
    for i in range(2):
        e2e_step(i, True, only=("sigma",))
    barrier()
    ev[0].record()
    for i in range(ksteps):
        e2e_step(i, True, only=("sigma",))
    ev[1].record()
    barrier()
    e2e_sigma_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / ksteps
    h2d_sigma = int(pinned["sigma"][t0e:t1e].numel() * pinned["sigma"].element_size())
    # the sampler ran over the timed assembly steps, the SpMV / Krylov section and the e2e steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:  # bytes copied by all ranks together
        t = torch.tensor([h2d, h2d_sigma], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        h2d, h2d_sigma = int(t[0].item()), int(t[1].item())
    e2e = {"value": T / (e2e_ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms, "ms_per_step_without_copy_overlap": e2e_serial_ms,
           "pipeline": "inputs double buffered: H2D of step i+1 overlaps the kernels of step i",
           "result": "squared Frobenius norm of the assembled matrix: %.17g" % fro_host[0].real.item(),
           "outside_the_timed_region": "host mesh tables (setup.host_mesh_s) and the symbolic phase "
                                       "(setup.symbolic_s), once per mesh; the assembled matrix stays on the device",
           "sigma_only": {"value": T / (e2e_sigma_ms * 1e-3), "unit": "elements/s", "ms_per_step": e2e_sigma_ms,
                          "h2d_bytes_per_step": h2d_sigma, "d2h_bytes_per_step": 16,
                          "what": "fixed mesh resident in HBM, a new conductivity model uploaded every step "
                                  "(the multi-frequency / inversion loop); same kernels and read-back"}}

    # ---- multi-GPU parity: the NCCL path against a one-GPU evaluation on rank 0 ----------------------------
    parity = None
    if world > 1 and not args.no_solve:
        parity = parity_block(el, tab, p, args.order, dev, rank, world, dist, sums)
        if parity is not None:
            parity["transport"] = ctx.transport  # "peer": pg_comm kernels over NVLink; "nccl": torch.distributed

    # ---- CPU baselines (rank 0, N=1 only) --------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import multiprocessing as mp

        cores = os.cpu_count() or 1
        nsample = args.cpu_sample or {1: 40000, 2: 12000, 3: 3000, 4: 800, 5: 240, 6: 120}[p]
        small = build_case(16, p)
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_assembly_rate(small, p, cores * 8, pool, cores)
            rate, t_elem, t_asm = cpu_assembly_rate(small, p, nsample, pool, cores)
        cpu = {"value": rate, "unit": "elements/s", "cores": cores, "kind": "port",
               "sample": "%d elements of the same mesh family: oracle element_system on %d processes (%.1f s) + "
                         "scatter-add of their cliques (%.1f s)" % (nsample, cores, t_elem, t_asm),
               # the real reference is ~4x slower per core than this port: SURVEY probe of the unmodified
               # computeElementalMatrices, 43 ms per element at p = 2 on one core
               "reference_python_ms_per_element_p2": 43.0,
               "petsc4py_mpirun_on_gpu_box": "absent (profiles/r2_probe_petsc4py_mpirun.txt): the north_star's "
                                             "mpirun petsc4py arm cannot run; oracle port instead"}
        cpu["spmv"] = cpu_spmv_baseline(A, cores)
        if not args.no_solve:
            A1, b1 = c1_system(dev)
            cpu["tts"] = [cpu_gmres_tts(A1, b1, "C1: reference test mesh, 9453 tets, p=1 (case1 physics)", 12.0)]
            # the same two systems on the GPU (same solver settings) so that the pair can be read side by side
            for (Ax, bx, lab) in ((A1, b1, "C1"),):
                torch.cuda.synchronize()
                t0 = time.time()
                rg = krylov.solve(Ax, bx, {"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-8})
                torch.cuda.synchronize()
                cpu["tts"][-1]["gpu_same_settings"] = {"iterations": rg.iterations, "seconds": time.time() - t0,
                                                       "converged": bool(rg.converged)}
            del A1, b1
            A2, b2, t2 = box_system(dev, 24, 1)
            cpu["tts"].append(cpu_gmres_tts(A2, b2, "slice of C2: synthetic box m=24, %d tets, p=1" % t2, 15.0))
            torch.cuda.synchronize()
            t0 = time.time()
            rg = krylov.solve(A2, b2, {"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-8,
                                       "ksp_max_it": cpu["tts"][-1]["iterations"] if not cpu["tts"][-1]["converged"]
                                       else 100000})
            torch.cuda.synchronize()
            cpu["tts"][-1]["gpu_same_settings"] = {"iterations": rg.iterations, "seconds": time.time() - t0,
                                                   "converged": bool(rg.converged)}
            del A2, b2

    # ---- the other named configurations (BASELINE configs[3], configs[4]); C3 is released first ------------
    n_dofs, nnz_own = int(plan.N), plan.nnz
    del A, op, vals, gbuf, plan, el, els, sets, dev_rows, pinned
    if not args.no_solve:
        del op_h, b
    xg = yg = None
    torch.cuda.empty_cache()
    extras = {}
    for name in [e for e in args.extras.split(",") if e and e != "none"]:
        try:
            extras[name] = {"c4": run_c4, "c5": run_c5}[name](args, dev, world, rank, dist if world > 1 else None,
                                                            peak, peak_src)
        except Exception as err:  # an extra configuration must not take the headline line down with it
            extras[name] = {"error": "%s: %s" % (type(err).__name__, err)}
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC % p, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic layered-earth CSEM box, m=%d -> %d tets, p=%d, N=%d dofs "
                                   "(BASELINE configs[2])" % (args.m, T, p, n_dofs),
                       "tets": T, "p": p, "dofs": n_dofs, "nnz": int(nnz_total), "order": args.order,
                       "partition": "PETSc-style contiguous row blocks, entity aligned" if world > 1 else "single GPU",
                       "l2": "inputs (%.1f GB) and output (%.1f GB) larger than L2; no flush needed"
                             % (T * 0.364e-6 * 1e3 / 1e3, nnz_own * 16e-9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 2 * args.steps,
            "clocks": clocks, "spmv": spmv, "solve": solve, "tts": tts, "parity": parity,
            "setup": {"host_mesh_s": tab["host_prep_s"], "symbolic_s": symbolic_s},
        }
        line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

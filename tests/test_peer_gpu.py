"""Peer-memory transport (csrc/pg_comm.cu, petgem_b200/peer.py) on ONE GPU: two (three) processes share
cuda:0, map each other's buffers over CUDA IPC and run the halo push / flag wait / all-reduce kernels
against each other (the GPU time-slices between the processes, so this is slow but exercises exactly the
code of the multi-GPU runs).  torch.distributed (gloo) only carries the set-up.  Compared with the
single-process results: MatMult with the pushed halo, the distributed Hiptmair preconditioner, and COCR
solves (CUDA-graph replay included) of the case1 system on the reference's test mesh (solver.py:584-590)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _system(topo, p, row_range=None, order=None):
    """case1 matrix (Dirichlet fused) on the reference's test mesh, whole or one row block."""
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    sig = np.array([1.0, 0.01, 1.0, 3.3333])[topo["tags"] - 1]
    el = ElementData.from_mesh(topo["nodes"], topo["elemsN"], topo["elemsE"], topo["edgesNodes"], topo["elemsF"],
                               topo["facesE"], np.stack([sig, sig], axis=1))
    geo, code = el.geometry()
    plan = AssemblyPlan(el, p, order="locality" if order is None else order, row_range=row_range)
    bd = np.zeros(plan.nEnt, dtype=np.uint8)
    bd[topo["bEdges"]] = 1
    if p >= 2:
        bd[topo["edgesNodes"].shape[0] + topo["bFaces"]] = 1
    plan.set_dirichlet(bd)
    vals = plan.assemble(geo, code, 2 * np.pi * 2.0, 4e-7 * np.pi, apply_dirichlet=True)
    A = CSRMatrix(*plan.csr(), vals, plan.N, plan.row_begin, plan=plan)
    return A, plan


def _worker(rank, world, port, p, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PG_TRANSPORT"] = "peer"
    import sys

    sys.path.insert(0, ROOT)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from petgem_b200 import krylov

        topo = dict(np.load(os.path.join(ROOT, "tests", "golden", "test_mesh_topology.npz"), allow_pickle=False))
        # --- the whole problem in this process (reference for the distributed run) ---
        Afull, full = _system(topo, p)
        N = full.N
        cuts = [0] + [full.entity_aligned_row(N * r // world) for r in range(1, world)] + [N]
        rng = np.random.default_rng(5)
        xg = torch.as_tensor(rng.normal(size=N) + 1j * rng.normal(size=N), device=dev)
        b = torch.zeros((N,), dtype=torch.complex128, device=dev)
        free = torch.nonzero(~Afull.dirichlet_mask.to(torch.bool)).reshape(-1)
        b[free[free.numel() // 2]] = 1.0j
        b[free[free.numel() // 3]] = 0.5
        yfull = Afull.mult(xg)
        op1 = krylov.Operator(Afull, pc="hiptmair")
        zfull = op1.precond(xg, torch.empty_like(xg))
        ref = krylov.cocr(op1, b, rtol=1e-10, maxit=4000)
        assert ref.converged
        lo, hi = cuts[rank], cuts[rank + 1]

        # --- one row block per process, peer transport ---
        A, plan = _system(topo, p, row_range=(lo, hi), order=full.order_host)
        ctx = krylov.DistContext(cuts[:-1], N)
        assert ctx.peer is not None and ctx.transport == "peer"
        # all-reduce kernel: many rounds, values that identify rank and round
        t = torch.zeros((5,), dtype=torch.complex128, device=dev)
        for i in range(40):
            t[:] = torch.arange(5, device=dev) * (rank + 1) + 1j * i
            ctx.allreduce(t)
            exp = np.arange(5) * (world * (world + 1) // 2) + 1j * i * world
            assert np.array_equal(t.cpu().numpy(), exp), (i, t)
        op = krylov.Operator(A, pc="hiptmair", ctx=ctx, halo="p2p")
        assert op.mode == "p2p" and op.xchg is not None
        # MatMult with the pushed halo: scratch vector (copy) and resident vector (no copy), twice each
        y = torch.empty((hi - lo,), dtype=torch.complex128, device=dev)
        tol = 1e-13 * yfull.abs().max().item()
        for _ in range(2):
            op.matvec(xg[lo:hi].contiguous(), y)
            assert (y - yfull[lo:hi]).abs().max().item() <= tol
        res_v = op.halo_vector(1)[: hi - lo]
        res_v[:, 0] = xg[lo:hi]
        for _ in range(2):
            op.matmat(res_v, y.reshape(-1, 1))
            assert (y - yfull[lo:hi]).abs().max().item() <= tol
        X4 = torch.stack([xg, 2 * xg, -xg, 1j * xg], dim=1).contiguous()
        Y4 = torch.empty((hi - lo, 4), dtype=torch.complex128, device=dev)
        op.matmat(X4[lo:hi].contiguous(), Y4)
        assert (Y4[:, 3] - 1j * yfull[lo:hi]).abs().max().item() <= 4 * tol
        # distributed Hiptmair application (interface sums over peer memory)
        z = op.precond(xg[lo:hi].contiguous(), torch.empty_like(y))
        assert (z - zfull[lo:hi]).abs().max().item() <= 1e-11 * zfull.abs().max().item()
        # COCR solve: graph replay over the peer kernels; same iteration count (+- one check interval) and x
        for graph in ("1", "0"):
            os.environ["PG_CUDA_GRAPH"] = graph
            res = krylov.cocr(op, b[lo:hi].contiguous(), rtol=1e-10, maxit=4000)
            assert res.converged, (res.reason, res.iterations)
            assert abs(res.iterations - ref.iterations) <= 10, (res.iterations, ref.iterations)
            err = (res.x - ref.x[lo:hi]).abs().max().item()
            assert err <= 1e-7 * ref.x.abs().max().item(), err
        # two right-hand sides in lockstep
        B2 = torch.stack([b, 2j * b], dim=1)[lo:hi].contiguous()
        os.environ["PG_CUDA_GRAPH"] = "1"
        rm = krylov.cocg_multi(op, B2, rtol=1e-10, maxit=4000, method="cocr")
        assert rm.converged.all()
        assert (rm.x[:, 1] - 2j * ref.x[lo:hi]).abs().max().item() <= 2e-7 * ref.x.abs().max().item()
        # interior-first MatMult: halo push on a second stream while the halo-free rows are multiplied
        os.environ["PG_HALO_OVERLAP"] = "force"
        op2 = krylov.Operator(A, pc="jacobi", ctx=ctx, halo="p2p")
        os.environ["PG_HALO_OVERLAP"] = "0"
        if p == 2:
            assert op2._split is not None and 0 <= op2._split[0] <= op2._split[1] <= op2._split[2]
        y2 = torch.empty_like(y)
        for _ in range(3):
            op2.matvec(xg[lo:hi].contiguous(), y2)
            assert (y2 - yfull[lo:hi]).abs().max().item() <= tol
        op2.matmat(X4[lo:hi].contiguous(), Y4)
        assert (Y4[:, 3] - 1j * yfull[lo:hi]).abs().max().item() <= 4 * tol
        os.environ["PG_CUDA_GRAPH"] = "1"
        ref_j = krylov.cocr(krylov.Operator(Afull, pc="jacobi"), b, rtol=1e-8, maxit=20000)
        res_j = krylov.cocr(op2, b[lo:hi].contiguous(), rtol=1e-8, maxit=20000)
        assert res_j.converged and abs(res_j.iterations - ref_j.iterations) <= 10, (res_j.iterations, ref_j.iterations)
        assert (res_j.x - ref_j.x[lo:hi]).abs().max().item() <= 1e-6 * ref_j.x.abs().max().item()
        ctx.peer.status()
        torch.cuda.synchronize()
        op2.close()
        op.close()
        out[rank] = int(res.iterations)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p", [(2, 2), (3, 1)])
def test_peer_transport_ranks_sharing_one_gpu(world, p):
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), p, out), nprocs=world, join=True)
    assert len(out) == world and len(set(out.values())) == 1, dict(out)

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: PETSc-style row-block ownership,
column remapping into the padded all-gather buffer, and the all-reduced dot products.  The SpMV
itself is emulated with scipy here; the CUDA kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, cuts, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from petgem_b200.krylov import DistContext

        rng = np.random.default_rng(seed)
        A = sp.random(N, N, density=0.05, random_state=seed, format="csr") + sp.identity(N, format="csr")
        A = (A + 1j * A.T).tocsr()
        A.sort_indices()
        x = rng.normal(size=N) + 1j * rng.normal(size=N)
        ctx = DistContext(cuts[:-1], N)
        assert ctx.sizes == [cuts[r + 1] - cuts[r] for r in range(world)]
        lo, hi = cuts[rank], cuts[rank + 1]
        Aloc = A[lo:hi]
        col_local = ctx.remap_columns(torch.from_numpy(Aloc.indices.astype(np.int32))).numpy()
        # every remapped column points at the right entry of the padded gather buffer
        send = torch.zeros(ctx.max_rows, dtype=torch.complex128)
        full = torch.zeros(world * ctx.max_rows, dtype=torch.complex128)
        ctx.gather(torch.from_numpy(x[lo:hi].copy()), send, full)
        assert np.array_equal(full.numpy()[col_local], x[Aloc.indices])
        Ah = sp.csr_matrix((Aloc.data, col_local, Aloc.indptr), shape=(hi - lo, world * ctx.max_rows))
        y = Ah @ full.numpy()
        assert np.abs(y - (A @ x)[lo:hi]).max() < 1e-12
        # neighbour halo: [own | received] buffer and remapped columns give the same product
        col_h, send_idx, send_splits, recv_splits = ctx.build_halo(torch.from_numpy(Aloc.indices.astype(np.int32)))
        nloc, next_ = hi - lo, int(sum(recv_splits))
        sendbuf = torch.zeros(int(sum(send_splits)), dtype=torch.complex128)
        xbuf = torch.zeros(nloc + next_, dtype=torch.complex128)
        ctx.exchange(torch.from_numpy(x[lo:hi].copy()), send_idx, send_splits, recv_splits, sendbuf, xbuf)
        assert np.array_equal(xbuf.numpy()[col_h.numpy()], x[Aloc.indices])
        Ap = sp.csr_matrix((Aloc.data, col_h.numpy(), Aloc.indptr), shape=(nloc, nloc + next_))
        assert np.abs(Ap @ xbuf.numpy() - (A @ x)[lo:hi]).max() < 1e-12
        # the same halo for an interleaved block of k right-hand sides (lockstep multi-source solve)
        from petgem_b200.krylov import _exchange_multi

        for k in (2, 4):
            Xk = rng.normal(size=(N, k)) + 1j * rng.normal(size=(N, k))
            sb = torch.zeros((int(sum(send_splits)), k), dtype=torch.complex128)
            xb = torch.zeros((nloc + next_, k), dtype=torch.complex128)
            _exchange_multi(ctx, torch.from_numpy(Xk[lo:hi].copy()), send_idx, send_splits, recv_splits, sb, xb)
            assert np.array_equal(xb.numpy()[col_h.numpy()], Xk[Aloc.indices])
            assert np.abs(Ap @ xb.numpy() - (A @ Xk)[lo:hi]).max() < 1e-12
        # decisions taken from per-rank clocks are agreed on (max over ranks), and the set-up helpers reduce
        assert ctx.transport == "nccl" and ctx.peer is None  # no CUDA here: torch.distributed collectives
        assert ctx.agree(rank == 1) is True and ctx.agree(False) is False
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        assert ctx._sum(t).item() == 3.0
        # all-reduced dot product equals the global one
        d = torch.tensor([np.vdot(x[lo:hi], y)], dtype=torch.complex128)
        ctx.allreduce(d)
        assert abs(d.item() - np.vdot(x, A @ x)) < 1e-10
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_row_block_halo_logic_world2():
    world, N = 2, 257
    cuts = [0, 130, N]  # uneven blocks -> padding exercised
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), N, cuts, 3, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}

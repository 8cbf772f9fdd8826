"""Krylov drivers (restarted GMRES, BiCGStab, TFQMR, COCG) over the C-ABI vector kernels.

Stands in for ``ksp.solve(b, x)`` at ``petgem/solver.py:584-590`` with the PETSc
defaults that apply silently there: GMRES restart 30, LEFT
preconditioning, classical Gram-Schmidt (one VecMDot + one VecMAXPY per
iteration), zero initial guess, convergence on the preconditioned residual norm
relative to ||M^-1 b||, maxit 10000.  Host code is Python; every vector operation
is a kernel from include/petgem_b200.h working on device scalars.  With a process
group, rows are owned PETSc-style by contiguous blocks: before the SpMV the x halo
is exchanged over NCCL (packed neighbour entries, all-gather as the fallback) and
the dot products are all-reduced.  Right-hand sides that share A (several sources,
the MT polarizations) advance in lockstep through ``solve_multi`` / ``cocg_multi``:
interleaved [n, k] blocks, one pass over the matrix per iteration for all k.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .device import CSRMatrix

_C128 = torch.complex128


class DistContext:
    """Row-block ownership across ranks (PETSc-style contiguous blocks, entity aligned).

    transport: how the two per-iteration collectives travel.  "peer" (default on CUDA): kernels over
    IPC-mapped peer memory (peer.py / csrc/pg_comm.cu: halo push + flag wait, one-kernel all-reduce, CUDA
    graph capturable); "nccl": torch.distributed collectives (all_to_all_single + all_reduce), kept as the
    A/B baseline and for the CPU (gloo) tests of the host logic.  PG_TRANSPORT overrides the default."""

    def __init__(self, row_begins, N, group=None, transport=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.row_begins = [int(r) for r in row_begins] + [int(N)]
        self.N = int(N)
        self.sizes = [self.row_begins[i + 1] - self.row_begins[i] for i in range(self.world)]
        self.max_rows = max(self.sizes)
        self._staged = dist.get_backend(group) == "gloo"  # gloo moves host tensors only
        transport = transport or os.environ.get("PG_TRANSPORT", "auto")
        if transport == "auto":
            transport = "peer" if (torch.cuda.is_available() and self.world > 1) else "nccl"
        if transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        self.transport = transport
        self.peer = None
        if transport == "peer" and self.world > 1:
            from .peer import PeerComm

            self.peer = PeerComm(dist, group)

    # set-up collectives (torch.distributed; staged through the host when the backend is gloo)
    def _a2a(self, out, inp, out_splits=None, in_splits=None):
        if self._staged and inp.is_cuda:
            o = torch.empty(out.shape, dtype=out.dtype)
            self.dist.all_to_all_single(o, inp.cpu(), output_split_sizes=out_splits, input_split_sizes=in_splits,
                                        group=self.group)
            out.copy_(o)
        else:
            self.dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits,
                                        group=self.group)
        return out

    def _sum(self, t, op=None):
        op = op or self.dist.ReduceOp.SUM
        if self._staged and t.is_cuda:
            h = t.cpu()
            self.dist.all_reduce(h, op=op, group=self.group)
            t.copy_(h)
        else:
            self.dist.all_reduce(t, op=op, group=self.group)
        return t

    def remap_columns(self, colidx: torch.Tensor) -> torch.Tensor:
        """Global column -> index in the padded all-gather buffer [world, max_rows]."""
        starts = torch.tensor(self.row_begins[:-1], dtype=torch.int64, device=colidx.device)
        c = colidx.to(torch.int64)
        owner = torch.bucketize(c, starts, right=True) - 1
        return (owner * self.max_rows + (c - starts[owner])).to(torch.int32)

    def gather(self, local: torch.Tensor, padded_send: torch.Tensor, full: torch.Tensor) -> torch.Tensor:
        padded_send[: local.numel()].copy_(local)
        self.dist.all_gather_into_tensor(torch.view_as_real(full).view(-1), torch.view_as_real(padded_send).view(-1),
                                         group=self.group)
        return full

    def allreduce(self, t: torch.Tensor) -> None:
        """Sum of a few complex scalars over the ranks (the VecDot / VecNorm all-reduce)."""
        if self.peer is not None and t.is_cuda and t.dtype == _C128 and t.numel() <= 64 and t.is_contiguous():
            self.peer.allreduce(t)
        else:
            self._sum(torch.view_as_real(t))

    def agree(self, flag: bool) -> bool:
        """True on every rank if `flag` is true on any (decisions taken from per-rank wall clocks)."""
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64,
                         device="cuda" if torch.cuda.is_available() else "cpu")
        self._sum(t, self.dist.ReduceOp.MAX)
        return bool(t.item() > 0)

    def build_halo(self, colidx: torch.Tensor):
        """Neighbour halo instead of the full all-gather (what PETSc's VecScatter does for MatMult).

        Returns (colidx_local, send_idx, send_splits, recv_splits): columns remapped into
        [own rows | received halo entries], the local row indices this rank must send (grouped by
        destination rank) and the per-rank counts of the exchange.  With the element-major
        internal numbering the halo is the thin interface between spatial slabs."""
        dev = colidx.device
        lo, hi = self.row_begins[self.rank], self.row_begins[self.rank + 1]
        c = colidx.to(torch.int64)
        outside = (c < lo) | (c >= hi)
        ext = torch.unique(c[outside])  # sorted global columns owned by other ranks
        starts = torch.tensor(self.row_begins[:-1], dtype=torch.int64, device=dev)
        owner = torch.bucketize(ext, starts, right=True) - 1
        recv_counts = torch.bincount(owner, minlength=self.world)
        send_counts = torch.empty_like(recv_counts)
        self._a2a(send_counts, recv_counts)
        recv_splits, send_splits = recv_counts.tolist(), send_counts.tolist()
        wanted = torch.empty((int(sum(send_splits)),), dtype=torch.int64, device=dev)
        self._a2a(wanted, ext, send_splits, recv_splits)
        send_idx = wanted - lo  # rows of mine the others asked for, grouped by destination
        n = hi - lo
        self._halo_ext, self._halo_lo, self._halo_hi = ext, lo, hi
        local = torch.where(outside, n + torch.searchsorted(ext, c), c - lo)
        return local.to(torch.int32), send_idx, send_splits, recv_splits

    def remap_halo(self, cols: torch.Tensor) -> torch.Tensor:
        """Same column map as build_halo for another index array (e.g. the entity column starts)."""
        c = cols.to(torch.int64)
        outside = (c < self._halo_lo) | (c >= self._halo_hi)
        n = self._halo_hi - self._halo_lo
        return torch.where(outside, n + torch.searchsorted(self._halo_ext, c), c - self._halo_lo).to(torch.int32)

    def exchange(self, x_local, send_idx, send_splits, recv_splits, sendbuf, xbuf):
        """xbuf = [x_local | halo]: pack, all_to_all over NCCL, unpack in place (transport "nccl")."""
        n = x_local.numel()
        xbuf[:n].copy_(x_local)
        torch.index_select(x_local, 0, send_idx, out=sendbuf)
        self.dist.all_to_all_single(torch.view_as_real(xbuf[n:]).view(-1), torch.view_as_real(sendbuf).view(-1),
                                    output_split_sizes=[2 * s for s in recv_splits],
                                    input_split_sizes=[2 * s for s in send_splits], group=self.group)
        return xbuf


def _exchange_multi(ctx, X, send_idx, send_splits, recv_splits, sendbuf, xbuf):
    """DistContext.exchange for an interleaved [n, k] block: xbuf = [X | halo rows]."""
    n, k = X.shape
    xbuf[:n].copy_(X)
    torch.index_select(X, 0, send_idx, out=sendbuf)
    ctx.dist.all_to_all_single(torch.view_as_real(xbuf[n:]).view(-1), torch.view_as_real(sendbuf).view(-1),
                               output_split_sizes=[2 * k * s for s in recv_splits],
                               input_split_sizes=[2 * k * s for s in send_splits], group=ctx.group)
    return xbuf


class Operator:
    """y = M^-1 A x on the owned rows, with the halo handled for the caller."""

    def __init__(self, A: CSRMatrix, pc: str = "none", ctx: DistContext = None, halo: str = "auto"):
        self.A, self.ctx = A, ctx
        self.n = A.rows
        dev = A.vals.device
        self.pc = pc
        self.inv_diag = None
        self.grad = None  # GradientSpace of the Hiptmair preconditioner
        if pc in ("jacobi", "hiptmair"):
            d = A.diagonal()
            self.inv_diag = torch.where(d == 0, torch.ones_like(d), 1.0 / d)  # PCJACOBI: zero diagonal -> 1
        elif pc != "none":
            raise ValueError("unsupported preconditioner %r (none, jacobi, hiptmair)" % pc)
        if pc == "hiptmair":
            if getattr(A, "asm_plan", None) is None:
                raise ValueError("-pc_type hiptmair needs the mesh behind the matrix: build the CSRMatrix with "
                                 "plan=AssemblyPlan(...)")
            if halo == "auto" and ctx is not None and ctx.world > 1:
                halo = "p2p"  # the gradient space is built in the [own | halo] numbering of the neighbour halo
        self.mode = "single"
        self.xchg = None  # peer-memory halo (transport "peer")
        if ctx is not None and ctx.world > 1:
            # x exchange before the SpMV: neighbour halo (packed all_to_all) when the off-block column
            # set is small -- the case with the element-major numbering -- else all-gather of x
            self.mode = halo
            if halo in ("auto", "p2p"):
                col_l, self.send_idx, self.send_splits, self.recv_splits = ctx.build_halo(A.colidx)
                next_ = int(sum(self.recv_splits))
                frac = torch.tensor([next_ / max(self.n, 1)], dtype=torch.float64, device=dev)
                ctx._sum(frac, ctx.dist.ReduceOp.MAX)
                self.mode = "p2p" if (halo == "p2p" or frac.item() < 0.5) else "allgather"
                if self.mode == "p2p":
                    self.halo_entries = next_
                    if ctx.peer is not None:
                        from .peer import PeerExchange

                        self.xchg = PeerExchange(ctx.peer, self.send_idx, self.send_splits, self.recv_splits)
                        self._halo_vectors = {}
                        self._split = None  # (ent_begin, ent_end, n_entities) of the halo-free rows, set below
                        # the vector for operands that do not live in peer-mapped memory (GMRES basis vectors, b):
                        # mapped now, so that the one-off IPC set-up is not paid inside the first solve
                        self.halo_vector(None, tag="scratch")
                    else:
                        self.sendbuf = torch.zeros((int(sum(self.send_splits)),), dtype=_C128, device=dev)
                        self.xbuf = torch.zeros((self.n + next_,), dtype=_C128, device=dev)
                    pref = getattr(A, "plan_ref", None)
                    cs = ctx.remap_halo(pref.column_starts()) if pref is not None else None
                    self.A_halo = CSRMatrix(A.rowptr, col_l, A.vals, self.n + next_, A.row_begin, plan=pref,
                                            colstart=cs, blocked=A.plan is not None)
            overlap = os.environ.get("PG_HALO_OVERLAP", "0")  # "1": when at least half the entities are interior; "force"
            if self.mode == "p2p" and self.xchg is not None and overlap in ("1", "force"):
                # interior-first MatMult: the entities whose columns are all owned are multiplied while the halo
                # push is in flight on a second stream; the rest after the flag wait
                sp = self.A_halo.halo_split(self.n)
                ok = torch.tensor([1.0 if (sp is not None and sp[1] - sp[0] >= (1 if overlap == "force" else sp[2] // 2))
                                   else 0.0],
                                  dtype=torch.float64, device=dev)
                ctx._sum(ok, ctx.dist.ReduceOp.MIN)  # the same launch sequence on every rank
                if ok.item() > 0:
                    self._split = sp
                    self._comm_stream = torch.cuda.Stream()
                    self._ev = [torch.cuda.Event(), torch.cuda.Event()]
            if self.mode == "allgather":
                self.colidx_local = ctx.remap_columns(A.colidx)
                self.send = torch.zeros((ctx.max_rows,), dtype=_C128, device=dev)
                self.full = torch.zeros((ctx.world * ctx.max_rows,), dtype=_C128, device=dev)
                cs = ctx.remap_columns(A.plan.column_starts()) if A.plan is not None else None
                self.A_halo = CSRMatrix(A.rowptr, self.colidx_local, A.vals, ctx.world * ctx.max_rows, A.row_begin,
                                        plan=A.plan, colstart=cs, blocked=A.plan is not None)
        if pc == "hiptmair":
            from .gradient import GradientSpace

            if self.mode == "allgather":
                raise ValueError("-pc_type hiptmair needs the neighbour halo (halo='p2p')")
            dist_run = self.mode == "p2p"
            self.grad = GradientSpace(A.asm_plan, dirichlet_rows=A.dirichlet_mask, ctx=ctx if dist_run else None,
                                      halo_ext=ctx._halo_ext if dist_run else None)
            self.grad.setup(self.A_halo if dist_run else A)
        self.spmv_calls = 0

    # -- peer transport: vectors that live in peer-mapped memory with their halo behind them ----------
    def halo_vector(self, k=None, tag="krylov"):
        """[n + halo] (k None) or [n + halo, k] complex vector in peer-mapped memory: the first n rows are
        this rank's entries, the tail is filled by the neighbours (pg_comm_push).  A Krylov driver keeps the
        vector it multiplies with A here (`[:n]` of it), so MatMult needs no copy of x.  Collective."""
        key = (k, tag)
        hv = self._halo_vectors.get(key)
        if hv is None:
            from .peer import SymmetricBuffer

            kk = 1 if k is None else int(k)
            buf = SymmetricBuffer(self.ctx.peer, (self.n + self.halo_entries) * kk * 16)
            t = buf.view(_C128, (self.n + self.halo_entries) * kk)
            t = t if k is None else t.view(self.n + self.halo_entries, kk)
            dst = self.xchg.target(buf, [sz * kk * 16 for sz in self.ctx.sizes], kk)
            hv = self._halo_vectors[key] = (t, dst, buf, kk)
            self._by_ptr = {v[0].data_ptr(): v for v in self._halo_vectors.values()}
        return hv[0]

    def _peer_product(self, x, k, mult, ranged=None):
        hv = self._by_ptr.get(x.data_ptr()) if self._halo_vectors else None
        if hv is None or hv[3] != k:
            full = self.halo_vector(None if x.dim() == 1 else k, tag="scratch")  # not resident: one copy of x
            full[: self.n].copy_(x)
            hv = self._by_ptr[full.data_ptr()]
        full = hv[0]
        if full.dim() != x.dim():  # an [n, 1] block and a vector share the layout
            full = full.reshape(-1) if x.dim() == 1 else full.reshape(-1, 1)
        if self._split is not None and ranged is not None:
            e0, e1, nb = self._split
            cur = torch.cuda.current_stream()
            self._ev[0].record(cur)                      # x is final
            self._comm_stream.wait_event(self._ev[0])
            with torch.cuda.stream(self._comm_stream):
                self.xchg.push(full, k, hv[1])           # overlaps the interior rows
                self._ev[1].record(self._comm_stream)
            ranged(full, (e0, e1))
            cur.wait_event(self._ev[1])
            self.xchg.wait()
            if e0 > 0:
                ranged(full, (0, e0))
            if e1 < nb:
                ranged(full, (e1, nb))
            self.xchg.ack()
            return
        self.xchg.push(full, k, hv[1])
        self.xchg.wait()
        mult(full)
        self.xchg.ack()

    def close(self):
        """Collective: release the peer-mapped vectors."""
        if self.xchg is not None:
            for _, _, buf, _ in self._halo_vectors.values():
                buf.close()
            self._halo_vectors, self._by_ptr = {}, {}
        if self.grad is not None:
            self.grad.close()

    def matvec(self, x: torch.Tensor, y: torch.Tensor, row_scale: torch.Tensor = None, dot: torch.Tensor = None):
        """y = A x on the owned rows.  With `dot` (one complex device scalar) the local part of x^T (A x) is left
        there by the same kernel when the matrix has a fused form; `self.fused_dot` says whether it was."""
        self._product(x, y, 1, row_scale, dot)
        return y

    def _product(self, x, y, k, row_scale, dot):
        self.fused_dot = False

        def mult(A, full):
            if dot is not None and row_scale is None:
                self.fused_dot = A.mult_fused_dot(full, y, k, dot)
            if not self.fused_dot:
                if k == 1 and full.dim() == 1:
                    A.mult(full, y, row_scale)
                else:
                    A.mult_multi(full, y, row_scale)

        if self.mode == "p2p" and self.xchg is not None:
            ranged = None
            if self._split is not None and dot is None and (k == 1 or k in (2, 4, 8)):
                def ranged(full, rng):
                    if k == 1 and full.dim() == 1:
                        self.A_halo.mult(full, y, row_scale, ent_range=rng)
                    else:
                        self.A_halo.mult_multi(full, y, row_scale, ent_range=rng)
            self._peer_product(x, k, lambda full: mult(self.A_halo, full), ranged)
        elif self.mode == "p2p":
            if x.dim() == 1:
                self.ctx.exchange(x, self.send_idx, self.send_splits, self.recv_splits, self.sendbuf, self.xbuf)
                mult(self.A_halo, self.xbuf)
            else:
                if getattr(self, "_mm", None) is None or self._mm[0] != k:
                    dev = x.device
                    self._mm = (k, torch.zeros((int(sum(self.send_splits)), k), dtype=_C128, device=dev),
                                torch.zeros((self.n + self.halo_entries, k), dtype=_C128, device=dev))
                _exchange_multi(self.ctx, x, self.send_idx, self.send_splits, self.recv_splits, self._mm[1], self._mm[2])
                mult(self.A_halo, self._mm[2])
        elif self.mode == "allgather":
            if x.dim() != 1:
                raise NotImplementedError("multi-right-hand-side SpMV needs the neighbour halo (halo='p2p')")
            self.ctx.gather(x, self.send, self.full)
            mult(self.A_halo, self.full)
        else:
            mult(self.A, x)
        self.spmv_calls += 1

    def matmat(self, X: torch.Tensor, Y: torch.Tensor, row_scale: torch.Tensor = None, dot: torch.Tensor = None):
        """Y = A X for k interleaved right-hand sides (X, Y: [n, k]); one pass over the matrix.  `dot` [k]: see
        matvec."""
        k = int(X.shape[1])
        if k == 1:  # an [n, 1] block is a vector: every halo mode applies
            self._product(X.reshape(-1), Y.reshape(-1), 1, row_scale, dot)
        else:
            self._product(X, Y, k, row_scale, dot)
        return Y

    def precond(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """y = M^-1 x for a vector [n] or an interleaved block [n, k] (x and y must not alias for hiptmair)."""
        if self.grad is not None:
            return self.grad.apply(x, self.inv_diag, y)
        if self.inv_diag is None:
            if y.data_ptr() != x.data_ptr():
                y.copy_(x)
        elif x.dim() == 2:
            check(lib().pg_zbscale_rows(self.n, int(x.shape[1]), ptr(self.inv_diag), ptr(x), ptr(y), stream_ptr()),
                  "pg_zbscale_rows")
        else:
            check(lib().pg_zpointwise_mult(self.n, ptr(x), ptr(self.inv_diag), ptr(y), stream_ptr()), "pg_zpointwise_mult")
        return y

    def apply(self, x, y, tmp=None):
        """y = M^-1 (A x); the Jacobi scaling rides in the SpMV epilogue, the gradient-space term of the
        Hiptmair preconditioner needs A x as a whole first (tmp)."""
        if self.grad is not None:
            if tmp is None:
                tmp = torch.empty_like(y)
            self.matvec(x, tmp)
            return self.grad.apply(tmp, self.inv_diag, y)
        return self.matvec(x, y, self.inv_diag)


class VecKernels:
    """Thin wrappers over the BLAS-1 entry points; results stay on the device."""

    def __init__(self, n, device, ctx: DistContext = None, kmax=32):
        self.n, self.ctx = int(n), ctx
        self.work = torch.empty((lib().pg_reduce_workspace_bytes(kmax) // 16,), dtype=_C128, device=device)
        self.dev = device

    def mdot(self, k, V, ldv, w, out):
        check(lib().pg_zmdotc(self.n, k, ptr(V), ldv, ptr(w), ptr(out), ptr(self.work), stream_ptr()), "pg_zmdotc")
        if self.ctx is not None and self.ctx.world > 1:
            self.ctx.allreduce(out[:k])
        return out

    def dot(self, x, y, out):
        """out[0] = sum conj(x) y."""
        check(lib().pg_zdotc(self.n, ptr(x), ptr(y), ptr(out), ptr(self.work), stream_ptr()), "pg_zdotc")
        if self.ctx is not None and self.ctx.world > 1:
            self.ctx.allreduce(out[:1])
        return out

    def dotu(self, x, y, out):
        """out[0] = sum x y (no conjugation)."""
        check(lib().pg_zdotu(self.n, ptr(x), ptr(y), ptr(out), ptr(self.work), stream_ptr()), "pg_zdotu")
        if self.ctx is not None and self.ctx.world > 1:
            self.ctx.allreduce(out[:1])
        return out

    def nrm2sq(self, x, out):
        check(lib().pg_dznrm2sq(self.n, ptr(x), ptr(out), ptr(self.work), stream_ptr()), "pg_dznrm2sq")
        if self.ctx is not None and self.ctx.world > 1:
            self.ctx.allreduce(out[:1])
        return out

    def maxpy(self, k, alpha, scale, V, ldv, w):
        check(lib().pg_zmaxpy(self.n, k, ptr(alpha), float(scale), ptr(V), ldv, ptr(w), stream_ptr()), "pg_zmaxpy")

    def maxpy_nrm2sq(self, k, alpha, scale, V, ldv, w, out):
        """w += scale * sum alpha_i V_i and out[0] = ||w||^2 of the result, one pass."""
        check(lib().pg_zmaxpy_nrm2sq(self.n, k, ptr(alpha), float(scale), ptr(V), ldv, ptr(w), ptr(out),
                                     ptr(self.work), stream_ptr()), "pg_zmaxpy_nrm2sq")
        if self.ctx is not None and self.ctx.world > 1:
            self.ctx.allreduce(out[:1])
        return out

    def copy_scaled(self, alpha, x, y, inv_real=False):
        check(lib().pg_zcopy_scaled(self.n, ptr(alpha), 1 if inv_real else 0, ptr(x), ptr(y), stream_ptr()),
              "pg_zcopy_scaled")

    def axpy(self, alpha, x, y):
        check(lib().pg_zaxpy(self.n, ptr(alpha), ptr(x), ptr(y), stream_ptr()), "pg_zaxpy")

    def aypx(self, beta, x, y):
        check(lib().pg_zaypx(self.n, ptr(beta), ptr(x), ptr(y), stream_ptr()), "pg_zaypx")

    def axpbypcz(self, a, x, b, y, c, z, w):
        check(lib().pg_zaxpbypcz(self.n, ptr(a), ptr(x), ptr(b), ptr(y), ptr(c), ptr(z), ptr(w), stream_ptr()),
              "pg_zaxpbypcz")

    def scal(self, alpha, x, inv_real=False):
        check(lib().pg_zscal(self.n, ptr(alpha), 1 if inv_real else 0, ptr(x), stream_ptr()), "pg_zscal")


class SolveResult:
    def __init__(self, x, iterations, residuals, converged, reason):
        self.x, self.iterations, self.residuals = x, iterations, residuals
        self.converged, self.reason = converged, reason


def gmres(op: Operator, b: torch.Tensor, rtol=1e-8, restart=30, maxit=10000, atol=1e-50, x0=None, monitor=None):
    """Left-preconditioned GMRES(restart), classical Gram-Schmidt (KSPGMRES defaults)."""
    n, dev = op.n, b.device
    vk = VecKernels(n, dev, op.ctx, kmax=restart + 2)
    V = torch.zeros((restart + 1, n), dtype=_C128, device=dev)
    w = torch.empty((n,), dtype=_C128, device=dev)
    tmp = torch.empty((n,), dtype=_C128, device=dev)
    x = torch.zeros((n,), dtype=_C128, device=dev) if x0 is None else x0.clone()
    hcol = torch.zeros((restart + 2,), dtype=_C128, device=dev)
    ycoef = torch.zeros((restart + 1,), dtype=_C128, device=dev)
    one = torch.ones((1,), dtype=_C128, device=dev)
    minus_one = -one
    scal = torch.zeros((2,), dtype=_C128, device=dev)

    op.precond(b, tmp)
    bnorm = math.sqrt(vk.nrm2sq(tmp, scal)[0].real.item())
    if bnorm == 0.0:
        return SolveResult(x, 0, [0.0], True, "zero rhs")
    tol = max(rtol * bnorm, atol)
    its, hist = 0, []
    x_is_zero = x0 is None
    while True:
        # r = M^-1 (b - A x)
        if x_is_zero:
            op.precond(b, V[0])
        else:
            op.matvec(x, tmp)
            vk.axpbypcz(one, b, minus_one, tmp, None, None, tmp)
            op.precond(tmp, V[0])
        beta = math.sqrt(vk.nrm2sq(V[0], scal)[0].real.item())
        if not hist:
            hist.append(beta)
            if monitor:
                monitor(0, beta)
        if beta <= tol:
            return SolveResult(x, its, hist, True, "rtol")
        if its >= maxit:
            return SolveResult(x, its, hist, False, "maxit")
        scal[1] = beta
        vk.scal(scal[1:], V[0], inv_real=True)
        H = np.zeros((restart + 1, restart), dtype=np.complex128)
        g = np.zeros(restart + 1, dtype=np.complex128)
        g[0] = beta
        cs = np.zeros(restart)
        sn = np.zeros(restart, dtype=np.complex128)
        k = 0
        while k < restart and its < maxit:
            op.apply(V[k], w, tmp)                     # w = M^-1 A v_k   (MatMult + PCApply fused)
            vk.mdot(k + 1, V, n, w, hcol)              # h_0..k = V^H w   (VecMDot)
            vk.maxpy_nrm2sq(k + 1, hcol, -1.0, V, n, w, hcol[k + 1:k + 2])  # w -= sum h_i v_i, ||w||^2
            hcol[k + 1] = torch.sqrt(hcol[k + 1].real)
            vk.copy_scaled(hcol[k + 1:k + 2], w, V[k + 1], inv_real=True)   # v_{k+1} = w / ||w||
            h = hcol[: k + 2].cpu().numpy()            # the one host sync of the iteration
            for i in range(k):                         # previous Givens rotations
                t = cs[i] * h[i] + sn[i] * h[i + 1]
                h[i + 1] = -np.conj(sn[i]) * h[i] + cs[i] * h[i + 1]
                h[i] = t
            a_, b_ = h[k], h[k + 1]
            den = math.sqrt(abs(a_) ** 2 + abs(b_) ** 2)
            if abs(a_) == 0.0:
                cs[k], sn[k] = 0.0, 1.0
            else:
                cs[k] = abs(a_) / den
                sn[k] = (a_ / abs(a_)) * np.conj(b_) / den
            h[k] = cs[k] * a_ + sn[k] * b_
            h[k + 1] = 0.0
            H[: k + 2, k] = h
            g[k + 1] = -np.conj(sn[k]) * g[k]
            g[k] = cs[k] * g[k]
            k += 1
            its += 1
            res = abs(g[k])
            hist.append(res)
            if monitor:
                monitor(its, res)
            if res <= tol or b_ == 0.0:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k else np.zeros(0)
        ycoef[:k] = torch.from_numpy(y).to(dev)
        vk.maxpy(k, ycoef, 1.0, V, n, x)               # x += V y
        x_is_zero = False
        if hist[-1] <= tol:
            return SolveResult(x, its, hist, True, "rtol")


def bicgstab(op: Operator, b: torch.Tensor, rtol=1e-8, maxit=10000, atol=1e-50, monitor=None):
    """Left-preconditioned BiCGStab (KSPBCGS)."""
    n, dev = op.n, b.device
    vk = VecKernels(n, dev, op.ctx, kmax=4)
    z = lambda: torch.zeros((n,), dtype=_C128, device=dev)  # noqa: E731
    x, r, rhat, p_, v, s, t, tmp = z(), z(), z(), z(), z(), z(), z(), z()
    sc = torch.zeros((8,), dtype=_C128, device=dev)
    one = torch.ones((1,), dtype=_C128, device=dev)
    op.precond(b, r)
    bnorm = math.sqrt(vk.nrm2sq(r, sc)[0].real.item())
    if bnorm == 0.0:
        return SolveResult(x, 0, [0.0], True, "zero rhs")
    tol = max(rtol * bnorm, atol)
    rhat.copy_(r)
    rho = alpha = omega = 1.0 + 0.0j
    hist = [bnorm]
    for it in range(1, maxit + 1):
        rho_new = complex(vk.dot(rhat, r, sc)[0].item())
        if rho_new == 0.0:
            return SolveResult(x, it - 1, hist, False, "breakdown rho")
        beta = (rho_new / rho) * (alpha / omega)
        # p = r + beta (p - omega v)
        sc[1], sc[2] = beta, -beta * omega
        vk.axpbypcz(one, r, sc[1:2], p_, sc[2:3], v, p_)
        op.apply(p_, v, tmp)
        den = complex(vk.dot(rhat, v, sc)[0].item())
        if den == 0.0:
            return SolveResult(x, it - 1, hist, False, "breakdown alpha")
        alpha = rho_new / den
        sc[1] = -alpha
        vk.axpbypcz(one, r, sc[1:2], v, None, None, s)     # s = r - alpha v
        op.apply(s, t, tmp)
        vk.dot(t, s, sc[3:4])
        vk.nrm2sq(t, sc[4:5])
        ts, tt = complex(sc[3].item()), sc[4].real.item()
        if tt == 0.0:
            return SolveResult(x, it - 1, hist, False, "breakdown omega")
        omega = ts / tt
        sc[1], sc[2] = alpha, omega
        vk.axpy(sc[1:2], p_, x)
        vk.axpy(sc[2:3], s, x)
        sc[1] = -omega
        vk.axpbypcz(one, s, sc[1:2], t, None, None, r)     # r = s - omega t
        rho = rho_new
        res = math.sqrt(vk.nrm2sq(r, sc)[0].real.item())
        hist.append(res)
        if monitor:
            monitor(it, res)
        if res <= tol:
            return SolveResult(x, it, hist, True, "rtol")
    return SolveResult(x, maxit, hist, False, "maxit")


def tfqmr(op: Operator, b: torch.Tensor, rtol=1e-8, maxit=10000, atol=1e-50, monitor=None):
    """Left-preconditioned transpose-free QMR (KSPTFQMR; Freund 1993, Saad Alg. 7.8): two operator
    applications per iteration, convergence on the quasi-residual bound tau*sqrt(m+1)."""
    n, dev = op.n, b.device
    vk = VecKernels(n, dev, op.ctx, kmax=4)
    z = lambda: torch.zeros((n,), dtype=_C128, device=dev)  # noqa: E731
    x, r0, w, u, v, d, Au, tmp = z(), z(), z(), z(), z(), z(), z(), z()
    sc = torch.zeros((8,), dtype=_C128, device=dev)
    one = torch.ones((1,), dtype=_C128, device=dev)
    op.precond(b, r0)
    tau = math.sqrt(vk.nrm2sq(r0, sc)[0].real.item())
    if tau == 0.0:
        return SolveResult(x, 0, [0.0], True, "zero rhs")
    tol = max(rtol * tau, atol)
    w.copy_(r0)
    u.copy_(r0)
    op.apply(u, v, tmp)
    Au.copy_(v)
    theta, eta, rho = 0.0, 0.0 + 0.0j, complex(tau * tau)
    hist = [tau]
    m = 0
    for it in range(1, maxit + 1):
        sigma = complex(vk.dot(r0, v, sc)[0].item())
        if sigma == 0.0:
            return SolveResult(x, it - 1, hist, False, "breakdown sigma")
        alpha = rho / sigma
        for half in (0, 1):
            if half == 1:  # u_{m+1} = u_m - alpha v_m
                sc[1] = -alpha
                vk.axpy(sc[1:2], v, u)
                op.apply(u, Au, tmp)
            sc[1] = -alpha
            vk.axpy(sc[1:2], Au, w)                          # w -= alpha B u
            sc[2] = theta * theta * eta / alpha
            vk.aypx(sc[2:3], u, d)                           # d = u + (theta^2 eta / alpha) d
            wn = math.sqrt(vk.nrm2sq(w, sc)[0].real.item())
            theta = wn / tau
            c = 1.0 / math.sqrt(1.0 + theta * theta)
            tau = tau * theta * c
            eta = c * c * alpha
            sc[3] = eta
            vk.axpy(sc[3:4], d, x)
            m += 1
            res = tau * math.sqrt(m + 1.0)
            if res <= tol:
                hist.append(res)
                if monitor:
                    monitor(it, res)
                return SolveResult(x, it, hist, True, "rtol")
        hist.append(res)
        if monitor:
            monitor(it, res)
        rho_new = complex(vk.dot(r0, w, sc)[0].item())
        if rho == 0.0 or rho_new == 0.0:
            return SolveResult(x, it, hist, False, "breakdown rho")
        beta = rho_new / rho
        rho = rho_new
        sc[1], sc[2] = beta, beta * beta
        vk.axpbypcz(sc[1:2], Au, sc[2:3], v, None, None, v)  # v = beta (B u_m + beta v)
        vk.aypx(sc[1:2], w, u)                               # u = w + beta u
        op.apply(u, Au, tmp)
        vk.axpy(one, Au, v)                                  # v += B u
    return SolveResult(x, maxit, hist, False, "maxit")


def cocg(op: Operator, b: torch.Tensor, rtol=1e-8, maxit=10000, atol=1e-50, monitor=None, check_every=10,
         max_seconds=None, norm_type="preconditioned"):
    """Conjugate-orthogonal CG for the complex SYMMETRIC system (KSPCG with -ksp_cg_type symmetric):
    one SpMV, two unconjugated dots and three vector updates per iteration, no restart.  Jacobi enters
    symmetrically through z = D^-1 r; convergence is tested on the preconditioned residual like PETSc.
    All coefficients (alpha = rho / p^T A p, beta = rho' / rho) are formed on the device, so the host
    synchronises only every `check_every` iterations to read the residual norm; the returned iteration
    count is therefore rounded up to that granularity.  This is cocg_multi with one right-hand side: per
    iteration the SpMV, one dot, the fused update/Jacobi/reduction pass and one AYPX."""
    mon = (lambda it, res: monitor(it, float(res[0]))) if monitor else None
    r = cocg_multi(op, b.reshape(-1, 1), rtol=rtol, maxit=maxit, atol=atol, monitor=mon, check_every=check_every,
                   max_seconds=max_seconds, norm_type=norm_type)
    return SolveResult(r.x.reshape(-1), r.iterations, [float(h[0]) for h in r.residuals], bool(r.converged[0]),
                       r.reason)


def cocr(op: Operator, b: torch.Tensor, rtol=1e-8, maxit=10000, atol=1e-50, monitor=None, check_every=10,
         max_seconds=None, norm_type="preconditioned"):
    """Conjugate-orthogonal conjugate residuals for the complex symmetric system (`-ksp_type cr`; PETSc's
    KSPCR is its Hermitian counterpart): see cocg_multi(method="cocr")."""
    mon = (lambda it, res: monitor(it, float(res[0]))) if monitor else None
    r = cocg_multi(op, b.reshape(-1, 1), rtol=rtol, maxit=maxit, atol=atol, monitor=mon, check_every=check_every,
                   max_seconds=max_seconds, method="cocr", norm_type=norm_type)
    return SolveResult(r.x.reshape(-1), r.iterations, [float(h[0]) for h in r.residuals], bool(r.converged[0]),
                       r.reason)


class MultiSolveResult:
    """Result of a lockstep solve of k right-hand sides: x is [n, k]."""

    def __init__(self, x, iterations, residuals, converged, reason):
        self.x, self.iterations, self.residuals = x, iterations, residuals
        self.converged, self.reason = converged, reason


def cocg_multi(op: Operator, B: torch.Tensor, rtol=1e-8, maxit=10000, atol=1e-50, monitor=None, check_every=10,
               max_seconds=None, method="cocg", norm_type="preconditioned"):
    """COCG (see cocg) on k right-hand sides in lockstep: B is [n, k], k in {1, 2, 4, 8}.  Per iteration
    one pass over the matrix for all k (pg_spmm), one fused update/Jacobi/reduction pass (pg_cocg_step),
    one dot and one AYPX; every right-hand side keeps its own alpha, beta and residual.  Iterates until
    every right-hand side meets its own tolerance (all-zero right-hand sides are born converged).

    method="cocr": conjugate-orthogonal conjugate residuals (Sogabe & Zhang 2007) on the preconditioned
    residual rt = M^-1 r: alpha = rt^T A rt / (A p)^T M^-1 (A p), x += alpha p, rt -= alpha M^-1 A p,
    beta = rt'^T A rt' / rt^T A rt, p = rt' + beta p, A p = A rt' + beta A p.  Still one SpMV per
    iteration, four more vector passes than COCG, and a smoother residual: 18-35 % fewer iterations on the
    reference's test mesh (p = 1, 2; measured with the same recurrences in numpy).

    norm_type (-ksp_norm_type): "preconditioned" (PETSc's default with left preconditioning) tests
    ||M^-1 r|| <= rtol ||M^-1 b||; "unpreconditioned" tests the TRUE residual ||b - A x|| <= rtol ||b||,
    recomputed with one extra pass over the matrix at the host checks once the preconditioned residual is
    within 100x of the tolerance (the Hiptmair preconditioner weights gradient components heavily, so the
    two norms differ by two orders of magnitude on the CSEM systems)."""
    import time as _time

    if norm_type not in ("preconditioned", "unpreconditioned"):
        raise ValueError("norm_type must be 'preconditioned' or 'unpreconditioned'")

    if method not in ("cocg", "cocr"):
        raise ValueError("method must be 'cocg' or 'cocr'")

    n, k = int(B.shape[0]), int(B.shape[1])
    if k not in (1, 2, 4, 8):
        raise ValueError("cocg_multi: k must be 1, 2, 4 or 8 (pad with zero right-hand sides)")
    dev = B.device
    L = lib()
    ctx = op.ctx if (op.ctx is not None and op.ctx.world > 1) else None
    work = torch.empty((L.pg_reduce_workspace_bytes(2 * k) // 16,), dtype=_C128, device=dev)
    Z_ = lambda: torch.zeros((n, k), dtype=_C128, device=dev)  # noqa: E731
    X, R, Z, P, Q = Z_(), B.contiguous().clone(), Z_(), Z_(), Z_()
    if op.xchg is not None:
        # peer transport: the vector that is multiplied with A every iteration (COCR: the preconditioned
        # residual, COCG: the direction) lives in peer-mapped memory in front of its halo
        hv = op.halo_vector(k)[:n]
        hv.zero_()
        if method == "cocr":
            Z = hv
        else:
            P = hv
    rho = [torch.zeros((k,), dtype=_C128, device=dev) for _ in range(2)]
    pq = torch.zeros((k,), dtype=_C128, device=dev)
    alpha2 = torch.zeros((2 * k,), dtype=_C128, device=dev)
    beta2 = torch.zeros((2 * k,), dtype=_C128, device=dev)
    out2 = torch.zeros((2 * k,), dtype=_C128, device=dev)
    st = stream_ptr

    def reduce_(t):
        if ctx is not None:
            ctx.allreduce(t)

    general = op.grad is not None  # preconditioner that is not a row scaling: explicit M^-1 applications
    dinv = None if general else op.inv_diag
    MQ = Z_() if general else None
    op.precond(R, Z)
    check(L.pg_zbnrm2sq(n, k, ptr(Z), ptr(out2), ptr(work), st()), "pg_zbnrm2sq")
    reduce_(out2[:k])
    bnorm = out2[:k].real.sqrt().cpu().numpy()
    tol = np.maximum(rtol * bnorm, atol)
    P.copy_(Z)
    check(L.pg_zbdotu(n, k, ptr(R), ptr(Z), ptr(rho[0]), ptr(work), st()), "pg_zbdotu")
    reduce_(rho[0])
    hist = [bnorm.copy()]
    if not (bnorm > 0).any():
        return MultiSolveResult(X, 0, hist, np.ones(k, dtype=bool), "zero rhs")
    unprec = norm_type == "unpreconditioned"
    true_hist = []
    if unprec:
        W = Z_()
        check(L.pg_zbnrm2sq(n, k, ptr(R), ptr(out2), ptr(work), st()), "pg_zbnrm2sq")  # R = B here
        reduce_(out2[:k])
        b2norm = out2[:k].real.sqrt().cpu().numpy()
        tol_true = np.maximum(rtol * b2norm, atol)
        Bc = B.contiguous()
        minus1 = torch.full((k,), -1.0, dtype=_C128, device=dev)

        def true_residual():
            op.matmat(X, W)                                                       # W = A X
            check(L.pg_zbaypx(n, k, ptr(minus1), ptr(Bc), ptr(W), st()), "pg_zbaypx")  # W = B - W
            check(L.pg_zbnrm2sq(n, k, ptr(W), ptr(alpha2), ptr(work), st()), "pg_zbnrm2sq")
            reduce_(alpha2[:k])
            return alpha2[:k].real.sqrt().cpu().numpy()
    it = 0
    done = np.zeros(k, dtype=bool)
    true_ratio, checks_since = None, 0
    t_start = _time.time()
    if method == "cocr":
        # Z = rt (preconditioned residual), Q = A p, AR = A rt; rho = rt^T A rt
        AR = Z_()
        op.matmat(Z, AR)
        Q.copy_(AR)
        check(L.pg_zbdotu(n, k, ptr(Z), ptr(AR), ptr(rho[0]), ptr(work), st()), "pg_zbdotu")
        reduce_(rho[0])

    # Fused dot products (pg_spmm_blocked_dot: r~^T A r~ in the MatMult epilogue; pg_cocr_direction_dot: the next
    # (A p)^T D^-1 (A p) in the direction pass) save two and one vector passes per iteration on paper.  Measured at
    # C3 on one B200 they LOSE: the block-level partial sum makes every SpMV block wait for its slowest warp
    # (COCR + Hiptmair 7.95 -> 8.49 ms per iteration, four sources in lockstep 17.3 -> 20.1 ms), the direction
    # variant is neutral.  Kept as tested options (tests/test_gpu_path.py), off by default.
    fuse = os.environ.get("PG_FUSED_SPMV_DOT", "0") == "1"
    fuse_dir = os.environ.get("PG_FUSED_DIRECTION_DOT", "0") == "1"
    pq_ready = [False]  # Jacobi: (A p)^T D^-1 (A p) of the next iteration comes out of the direction pass

    def iterations_cocr(count):
        cur = 0
        for _ in range(count):
            if general:   # MQ = M^-1 (A p), then the same recurrences with the row scaling dropped
                op.precond(Q, MQ)
                check(L.pg_zbdotu(n, k, ptr(Q), ptr(MQ), ptr(pq), ptr(work), st()), "pg_zbdotu")
                reduce_(pq)
            elif not pq_ready[0]:
                check(L.pg_zbdotu_w(n, k, ptr(Q), ptr(Q), ptr(dinv), ptr(pq), ptr(work), st()), "pg_zbdotu_w")
                reduce_(pq)
            check(L.pg_zbdiv(k, ptr(rho[cur]), ptr(pq), ptr(alpha2), st()), "pg_zbdiv")
            check(L.pg_cocr_update(n, k, ptr(alpha2), ptr(P), ptr(MQ if general else Q), ptr(dinv), ptr(X), ptr(Z),
                                   st()), "pg_cocr_update")
            # A r~ and r~^T A r~ in one pass over the matrix when the matrix has the fused kernel
            op.matmat(Z, AR, dot=rho[cur ^ 1] if fuse else None)
            if not op.fused_dot:
                check(L.pg_zbdotu(n, k, ptr(Z), ptr(AR), ptr(rho[cur ^ 1]), ptr(work), st()), "pg_zbdotu")
            reduce_(rho[cur ^ 1])
            check(L.pg_zbdiv(k, ptr(rho[cur ^ 1]), ptr(rho[cur]), ptr(beta2), st()), "pg_zbdiv")
            if general or not fuse_dir:
                check(L.pg_cocr_direction(n, k, ptr(beta2), ptr(Z), ptr(AR), ptr(P), ptr(Q), st()),
                      "pg_cocr_direction")
            else:
                check(L.pg_cocr_direction_dot(n, k, ptr(beta2), ptr(Z), ptr(AR), ptr(dinv), ptr(P), ptr(Q), ptr(pq),
                                              ptr(work), st()), "pg_cocr_direction_dot")
                reduce_(pq)
                pq_ready[0] = True
            cur ^= 1
        if cur:
            rho[0].copy_(rho[1])
        # the residual norm the host reads after the batch
        check(L.pg_zbnrm2sq(n, k, ptr(Z), ptr(out2[k:]), ptr(work), st()), "pg_zbnrm2sq")
        reduce_(out2[k:])

    def iterations_cocg(count):
        cur = 0  # count is even, or the last batch: rho[0] is the current rho on entry and on exit
        for _ in range(count):
            op.matmat(P, Q, dot=pq if fuse else None)
            if not op.fused_dot:
                check(L.pg_zbdotu(n, k, ptr(P), ptr(Q), ptr(pq), ptr(work), st()), "pg_zbdotu")
            reduce_(pq)
            check(L.pg_zbdiv(k, ptr(rho[cur]), ptr(pq), ptr(alpha2), st()), "pg_zbdiv")
            if general:   # x += alpha p, r -= alpha q, z = M^-1 r, rho' = r^T z, |z|^2
                check(L.pg_zbaxpy(n, k, ptr(alpha2), ptr(P), ptr(X), st()), "pg_zbaxpy")
                check(L.pg_zbaxpy(n, k, ptr(alpha2[k:]), ptr(Q), ptr(R), st()), "pg_zbaxpy")
                op.precond(R, Z)
                check(L.pg_zbdotu(n, k, ptr(R), ptr(Z), ptr(out2), ptr(work), st()), "pg_zbdotu")
                check(L.pg_zbnrm2sq(n, k, ptr(Z), ptr(out2[k:]), ptr(work), st()), "pg_zbnrm2sq")
            else:
                check(L.pg_cocg_step(n, k, ptr(alpha2), ptr(P), ptr(Q), ptr(dinv), ptr(X), ptr(R), ptr(Z),
                                     ptr(out2), ptr(work), st()), "pg_cocg_step")
            reduce_(out2)
            rho[cur ^ 1].copy_(out2[:k])
            check(L.pg_zbdiv(k, ptr(rho[cur ^ 1]), ptr(rho[cur]), ptr(beta2), st()), "pg_zbdiv")
            check(L.pg_zbaypx(n, k, ptr(beta2), ptr(Z), ptr(P), st()), "pg_zbaypx")
            cur ^= 1
        if cur:
            rho[0].copy_(rho[1])

    iterations = iterations_cocr if method == "cocr" else iterations_cocg

    # Single GPU: the `check_every` iterations between two host checks are one CUDA graph (no host work;
    # the ~10 launches per iteration otherwise dominate small systems).  Capture needs a non-default
    # stream, so the solve runs on a side stream; PG_CUDA_GRAPH=0 launches eagerly.
    import ctypes as _C

    gexec = _C.c_void_p()
    side = None
    first = False
    graphable = ctx is None or (ctx.peer is not None and op.mode == "p2p")  # every collective is a kernel
    if graphable and check_every > 0 and os.environ.get("PG_CUDA_GRAPH", "1") == "1" and maxit >= 2 * check_every:
        main_stream = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(main_stream)
        torch.cuda.set_stream(side)
    try:
        if side is not None:
            iterations(check_every)  # warm-up outside the capture (counts as iterations)
            it += check_every
            first = True
            if L.pg_graph_begin(st()) == 0:
                try:
                    iterations(check_every)
                finally:
                    if L.pg_graph_end(st(), _C.byref(gexec)) != 0:
                        gexec = _C.c_void_p()
        while it < maxit:
            if first:
                first = False  # the warm-up batch is checked before the first replay
            else:
                count = min(check_every, maxit - it)
                if gexec and count == check_every:
                    check(L.pg_graph_launch(gexec, st()), "pg_graph_launch")
                else:
                    iterations(count)
                it += count
            res2 = out2[k:].real.cpu().numpy()  # the host sync of this batch
            if ctx is not None and ctx.peer is not None:
                ctx.peer.status()  # a peer that stopped answering raises here instead of hanging the box
            if not np.isfinite(res2).all():
                return MultiSolveResult(X, it, hist, np.zeros(k, dtype=bool), "breakdown")
            res = np.sqrt(res2)
            hist.append(res)
            if monitor:
                monitor(it, res)
            done = res <= tol
            if unprec:
                # The true residual costs one more pass over the matrix, so it is not recomputed at every host check:
                # the ratio true / preconditioned norm of the last evaluation predicts it, and the next evaluation
                # waits until the prediction is within 2x of the tolerance (or 20 checks have gone by).  Every rank
                # sees the same all-reduced numbers, so the decision is the same everywhere.
                done = np.zeros(k, dtype=bool)
                if (res <= 100.0 * tol).all():
                    live = bnorm > 0
                    due = true_ratio is None or checks_since >= 20 or \
                        bool((true_ratio[live] * res[live] <= 2.0 * tol_true[live]).all())
                    if due:
                        tr = true_residual()
                        true_hist.append((it, tr))
                        done = tr <= tol_true
                        true_ratio = tr / np.maximum(res, 1e-300)
                        checks_since = 0
                    else:
                        checks_since += 1
            if done.all():
                out = MultiSolveResult(X, it, hist, done, "rtol")
                out.true_residuals = true_hist
                return out
            if max_seconds is not None and (_time.time() - t_start > max_seconds if ctx is None else
                                            ctx.agree(_time.time() - t_start > max_seconds)):
                out = MultiSolveResult(X, it, hist, done, "time limit")
                out.true_residuals = true_hist
                return out
        out = MultiSolveResult(X, maxit, hist, done if unprec else hist[-1] <= tol, "maxit")
        out.true_residuals = true_hist
        return out
    finally:
        if gexec:
            L.pg_graph_destroy(gexec)
        if side is not None:
            torch.cuda.set_stream(main_stream)
            main_stream.wait_stream(side)


def solve_multi(A: CSRMatrix, B: torch.Tensor, options=None, ctx: DistContext = None, monitor=None):
    """KSP front end for several right-hand sides sharing A: B is [n, nrhs] (any nrhs).  With
    -ksp_type cg -ksp_cg_type symmetric the right-hand sides advance in lockstep, up to 8 per pass over
    the matrix (cocg_multi); other solver types run one right-hand side after the other like the
    reference.  Returns (X [n, nrhs], list of per-batch results)."""
    o = dict(options or {})
    ksp = str(o.get("ksp_type", "gmres"))
    nrhs = int(B.shape[1])
    X = torch.empty((B.shape[0], nrhs), dtype=_C128, device=B.device)
    results = []
    if ksp not in ("cg", "cr"):
        for r in range(nrhs):
            res = solve(A, B[:, r].contiguous(), o, ctx=ctx, monitor=monitor)
            X[:, r] = res.x
            results.append(res)
        return X, results
    if ksp == "cg" and str(o.get("ksp_cg_type", "symmetric")) != "symmetric":
        raise ValueError("A is complex symmetric, not Hermitian: use -ksp_cg_type symmetric")
    pc = resolve_pc(o, have_mesh=getattr(A, "asm_plan", None) is not None)
    op = Operator(A, pc=pc, ctx=ctx, halo="p2p" if ctx is not None and ctx.world > 1 else "auto")
    rtol, maxit = float(o.get("ksp_rtol", 1e-5)), int(o.get("ksp_max_it", 10000))
    for r0 in range(0, nrhs, 8):
        kk = min(8, nrhs - r0)
        kpad = 1 if kk == 1 else 2 if kk == 2 else 4 if kk <= 4 else 8
        Bp = torch.zeros((B.shape[0], kpad), dtype=_C128, device=B.device)
        Bp[:, :kk] = B[:, r0:r0 + kk]
        res = cocg_multi(op, Bp, rtol=rtol, maxit=maxit, monitor=monitor, method="cocr" if ksp == "cr" else "cocg",
                         norm_type=str(o.get("ksp_norm_type", "preconditioned")))
        X[:, r0:r0 + kk] = res.x[:, :kk]
        results.append(res)
    return X, results


def parse_petsc_options(path_or_text):
    """Read the PETSc options file passed on the PETGEM command line
    (kernel.py:15, examples/case1/petsc.opts) -> dict of the options we honour."""
    import os

    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    opts = {}
    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        if not line.startswith("-"):
            continue
        parts = line.split()
        opts[parts[0][1:]] = parts[1] if len(parts) > 1 else True
    return opts


class UnsupportedSolverError(ValueError):
    """The options file asks for a solve this backend does not provide (direct factorisations)."""


# -pc_type values of the reference's shipped option files (examples/case*/petsc.opts, consumed by
# setFromOptions at solver.py:586-589) that have no counterpart here and what stands in for them
_PC_SUBSTITUTES = ("sor", "bjacobi", "asm", "gamg", "ilu", "icc", "eisenstat", "hypre", "ml")
_PC_DIRECT = ("lu", "cholesky", "svd")


def resolve_pc(options, notify=None, have_mesh=True):
    """-pc_type of the options file -> preconditioner of this backend.  Every substitution is announced on
    the master rank (`notify`, default Print.master); a direct factorisation (-pc_type lu/cholesky, usually
    with -ksp_type preonly: examples/case4/petsc.opts) is refused, there is no direct solver here."""
    if notify is None:
        from .common import Print
        notify = Print.master
    ksp = str(options.get("ksp_type", "gmres"))
    pc = str(options.get("pc_type", "jacobi"))
    if pc in _PC_DIRECT or ksp == "preonly":
        raise UnsupportedSolverError(
            "-ksp_type %s -pc_type %s asks for a direct factorisation (MUMPS/PETSc LU); the B200 backend has "
            "iterative solvers only: use -ksp_type cr|cg|gmres|bcgs|tfqmr with -pc_type jacobi" % (ksp, pc))
    if pc in _PC_SUBSTITUTES:
        # strongest preconditioner of this backend: Hiptmair's hybrid smoother when the mesh behind the
        # matrix is known (gradient.py), point Jacobi for a bare CSR matrix
        sub = "hiptmair" if have_mesh else "jacobi"
        notify("     -pc_type %s is not available on the B200 backend: using -pc_type %s instead "
               "(iteration counts will differ from the reference's)" % (pc, sub))
        pc = sub
    return pc


def solve(A: CSRMatrix, b: torch.Tensor, options=None, ctx: DistContext = None, monitor=None) -> SolveResult:
    """KSP front end: ksp_type gmres|bcgs|tfqmr|cg (symmetric)|cr, pc_type none|jacobi|hiptmair, ksp_rtol,
    ksp_gmres_restart, ksp_max_it.  Other -pc_type values of the reference's option files are replaced
    with a printed notice (resolve_pc); direct solves (-pc_type lu, -ksp_type preonly) raise."""
    o = dict(options or {})
    ksp = str(o.get("ksp_type", "gmres"))
    pc = resolve_pc(o, have_mesh=getattr(A, "asm_plan", None) is not None)
    rtol = float(o.get("ksp_rtol", 1e-5))  # PETSc default when the file does not set it
    maxit = int(o.get("ksp_max_it", 10000))
    op = Operator(A, pc=pc, ctx=ctx)
    norm_type = str(o.get("ksp_norm_type", "preconditioned"))
    if norm_type != "preconditioned" and ksp not in ("cg", "cr"):
        raise ValueError("-ksp_norm_type %s is available with -ksp_type cg|cr only (left preconditioning)" % norm_type)
    if ksp == "gmres":
        return gmres(op, b, rtol=rtol, restart=int(o.get("ksp_gmres_restart", 30)), maxit=maxit, monitor=monitor)
    if ksp in ("bcgs", "bicgstab"):
        return bicgstab(op, b, rtol=rtol, maxit=maxit, monitor=monitor)
    if ksp == "tfqmr":
        return tfqmr(op, b, rtol=rtol, maxit=maxit, monitor=monitor)
    if ksp == "cg":
        if str(o.get("ksp_cg_type", "symmetric")) != "symmetric":
            raise ValueError("A is complex symmetric, not Hermitian: use -ksp_cg_type symmetric")
        return cocg(op, b, rtol=rtol, maxit=maxit, monitor=monitor, norm_type=norm_type)
    if ksp == "cr":
        return cocr(op, b, rtol=rtol, maxit=maxit, monitor=monitor, norm_type=norm_type)
    raise ValueError("unsupported ksp_type %r (gmres, bcgs, tfqmr, cg, cr)" % ksp)

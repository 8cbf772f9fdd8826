"""Discrete gradient of the H1 space in the Nedelec dofs, and the Hiptmair preconditioner built on it.

The reference leaves preconditioning to PETSc (``-pc_type sor|asm|gamg`` in the shipped option files,
``examples/case*/petsc.opts``, consumed at ``petgem/solver.py:586-589``).  On the B200 path the
symmetric, fully parallel stand-in is the hybrid smoother of Hiptmair (1998):

    M^-1 = D^-1 + G diag(G^T A G)^-1 G^T

with ``G`` the discrete gradient: ``grad(psi_k) = sum_j G[j, k] N_j`` for the hierarchical H1 functions
``psi`` = vertex functions ``lambda_a`` and (p >= 2) quadratic edge functions ``lambda_a lambda_b``,
and ``N_j`` the global basis of ``hvfem.shape3DETet`` (hvfem.py:319-464).  The curl-curl part of A
annihilates ``range(G)``, which is why point Jacobi needs ~12 000 iterations at 2 Hz on the 5 M-tet
box; the gradient-space Jacobi term brings those modes back to the scale of the others.

Entries of G (all REAL, orientation handled by the global dof conventions):
  * Whitney dof of edge (a, b), a < b global node ids (first edge dof): -1 on node a, +1 on node b;
  * second edge dof E_1 = (lambda_b - lambda_a) w_ab: one entry on the edge's own quadratic function;
  * the first two dofs of a face (k = i + j = 1, families 0 and 1): entries on the three edge functions
    of the face, depending on the local face and its orientation code only;
  * every other dof: none (their gradients start at degree 3).
The coefficients are obtained once per process from the reference-element basis (``basis.py``) by a
small least-squares fit (``gradient_tables``), not hard-coded.

Setup (index manipulation) uses torch on the device; the numeric work (G^T r, G y, diag(G^T A G))
goes through the C ABI (pg_rcsr_apply, pg_galerkin_diagonal, pg_masked_reciprocal).
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from . import basis
from ._lib import check, lib, ptr, stream_ptr

_C128 = torch.complex128
# local face -> local edges (hvfem.py:177-184)
FACE_EDGES = np.array([[0, 1, 2], [0, 4, 3], [1, 5, 4], [2, 5, 3]], dtype=np.int64)


@functools.lru_cache(maxsize=None)
def gradient_tables():
    """(c_edge, F): coefficient of the edge function lambda_a lambda_b on the second dof of its edge, and
    F[local face, orientation, face dof 0..1, local edge of the face 0..2]: grad(lambda_a lambda_b) in the
    order-2 basis, fitted on the master tetrahedron (exact: the gradient lies in the space)."""
    rng = np.random.default_rng(0)
    pts = rng.dirichlet(np.ones(4), size=48)[:, 1:]
    Nx, _ = basis.evaluate_expanded(2, pts)
    lam = np.stack([1.0 - pts.sum(axis=1), pts[:, 0], pts[:, 1], pts[:, 2]])
    gl = basis.GRAD_LAMBDA
    R = np.stack([lam[a][:, None] * gl[b] + lam[b][:, None] * gl[a] for a, b in basis.LOCAL_EDGES])  # [6, npts, 3]
    R = R.reshape(6, -1).T
    F = np.zeros((4, 6, 2, 3))
    c_edge = None
    for o in range(6):
        J, S = basis.local_to_expanded(2, np.zeros(6, dtype=np.int64), np.full(4, o, dtype=np.int64))
        B = (Nx[J] * S[:, None, None]).reshape(J.size, -1).T
        C, *_ = np.linalg.lstsq(B, R, rcond=None)  # [20, 6]
        if np.abs(B @ C - R).max() > 1e-10:
            raise RuntimeError("gradient_tables: grad(lambda_a lambda_b) is not in the order-2 space")
        C[np.abs(C) < 1e-12] = 0.0
        ce = np.array([C[2 * e + 1, e] for e in range(6)])
        if np.abs(ce - ce[0]).max() > 1e-12 or (c_edge is not None and abs(ce[0] - c_edge) > 1e-12):
            raise RuntimeError("gradient_tables: inconsistent edge coefficient")
        c_edge = float(np.round(ce[0], 12))
        for f in range(4):
            for d in range(2):
                F[f, o, d] = C[12 + 2 * f + d, FACE_EDGES[f]]
        # nothing else may be non-zero: Whitney rows, other edges' E_1 rows, faces not containing the edge
        chk = C.copy()
        for e in range(6):
            chk[2 * e + 1, e] = 0.0
        for f in range(4):
            chk[12 + 2 * f:12 + 2 * f + 2, FACE_EDGES[f]] = 0.0
        if np.abs(chk).max() > 1e-12:
            raise RuntimeError("gradient_tables: unexpected coupling in the discrete gradient")
    return c_edge, np.round(F, 12)


def _csr_from_sorted(major, nrows, dev):
    counts = torch.bincount(major, minlength=nrows)
    rowptr = torch.zeros((nrows + 1,), dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(torch.int32)


class GradientSpace:
    """G restricted to the rows a process owns (plus, for the Galerkin diagonal, its halo rows)."""

    def __init__(self, plan, dirichlet_rows=None, ctx=None, halo_ext=None):
        """plan: AssemblyPlan of the matrix (its numbering and row block are used).
        dirichlet_rows: bool/uint8 tensor [N] in the numbering in use (rows fixed by zeroRowsColumns), or None.
        ctx, halo_ext: DistContext and the sorted global columns of the halo (multi-GPU)."""
        el, p = plan.elems, plan.p
        dev = el.device
        self.plan, self.ctx = plan, ctx
        self.p = p
        T, nE, nF = el.T, el.nEdges, el.nFaces
        nn = int(el.elemsN.max().item()) + 1
        nf = p * (p - 1)
        N = plan.N
        i64 = torch.int64
        perm = plan.dof_permutation().to(i64) if plan.order_host is not None else None

        def internal(ref):
            return ref if perm is None else perm[ref]

        elemsE, elemsF = el.elemsE.to(i64), el.elemsF.to(i64)
        # global edge -> (min, max) node pair, scattered from the per-element rows (edgesNodes.dat)
        EN = torch.zeros((nE, 2), dtype=i64, device=dev)
        EN[elemsE.reshape(-1)] = el.edgesNodes.to(i64).reshape(-1, 2)
        e_ids = torch.arange(nE, dtype=i64, device=dev)
        rows = [internal(e_ids * p).repeat_interleave(2)]
        cols = [EN.reshape(-1)]
        vals = [torch.tensor([-1.0, 1.0], dtype=torch.float64, device=dev).repeat(nE)]
        self.n_h1 = nn
        if p >= 2:
            c_edge, F = gradient_tables()
            self.n_h1 = nn + nE
            rows.append(internal(e_ids * p + 1))
            cols.append(nn + e_ids)
            vals.append(torch.full((nE,), c_edge, dtype=torch.float64, device=dev))
            # faces: coefficients from the first incident element (lowest index) of every face
            t_ids = torch.arange(T, dtype=i64, device=dev).repeat_interleave(4)
            first = torch.full((nF,), T, dtype=i64, device=dev)
            first.scatter_reduce_(0, elemsF.reshape(-1), t_ids, reduce="amin")
            f_ids = torch.arange(nF, dtype=i64, device=dev)
            lf = (elemsF[first] == f_ids[:, None]).to(torch.int8).argmax(dim=1)
            _, code = el.geometry()  # all elements: a halo face may hang on an element outside the owned range
            o = (code.to(i64)[first] >> (6 + 3 * lf)) & 7
            Ft = torch.as_tensor(F, device=dev)                      # [4, 6, 2, 3]
            fe = torch.as_tensor(FACE_EDGES, device=dev)[lf]         # [nF, 3] local edges
            ge = torch.gather(elemsE[first], 1, fe)                  # [nF, 3] global edges
            coef = Ft[lf, o]                                         # [nF, 2, 3]
            for d in range(2):
                rows.append(internal(nE * p + f_ids * nf + d).repeat_interleave(3))
                cols.append((nn + ge).reshape(-1))
                vals.append(coef[:, d, :].reshape(-1))
            del code
        rows, cols, vals = torch.cat(rows), torch.cat(cols), torch.cat(vals)
        keep = vals != 0
        # Dirichlet: fixed rows carry no gradient, and every H1 function that reaches a fixed row is fixed
        self.h1_fixed = torch.zeros((self.n_h1,), dtype=torch.bool, device=dev)
        if dirichlet_rows is not None:
            fixed_row = dirichlet_rows.to(torch.bool)[rows]
            self.h1_fixed[cols[fixed_row & keep]] = True
            keep &= ~fixed_row
            keep &= ~self.h1_fixed[cols]
        rows, cols, vals = rows[keep], cols[keep], vals[keep]

        # local numbering: owned rows first, then the halo columns of A
        lo, n_own = plan.row_begin, plan.local_rows
        self.n_own = n_own
        loc = rows - lo
        owned = (loc >= 0) & (loc < n_own)
        if halo_ext is not None and halo_ext.numel() > 0:
            pos = torch.searchsorted(halo_ext, rows).clamp_(max=halo_ext.numel() - 1)
            in_halo = (~owned) & (halo_ext[pos] == rows)
            loc = torch.where(owned, loc, torch.where(in_halo, n_own + pos, torch.full_like(loc, -1)))
        else:
            loc = torch.where(owned, loc, torch.full_like(loc, -1))
        # H1 functions this process needs: those reached by an owned row
        h1_ids = torch.unique(cols[owned])  # sorted global H1 ids
        self.h1_ids = h1_ids
        nh = int(h1_ids.numel())
        self.nh = nh
        if nh:
            cpos = torch.searchsorted(h1_ids, cols).clamp_(max=nh - 1)
            ok = (loc >= 0) & (h1_ids[cpos] == cols)
        else:
            cpos, ok = cols, torch.zeros_like(owned)
        loc, cpos, vals, owned = loc[ok], cpos[ok], vals[ok], owned[ok]
        n_ext = n_own + (0 if halo_ext is None else int(halo_ext.numel()))
        # G over the owned rows: CSR by row
        key = loc[owned] * max(nh, 1) + cpos[owned]
        order = torch.argsort(key)
        self.g_rowptr = _csr_from_sorted(loc[owned][order], n_own, dev)
        self.g_col = cpos[owned][order].to(torch.int32)
        self.g_val = vals[owned][order].contiguous()
        # G^T over the owned rows (apply) and over owned + halo rows (Galerkin diagonal): CSR by H1 function
        keyT = cpos * max(n_ext, 1) + loc
        orderT = torch.argsort(keyT)
        ext_rowptr = _csr_from_sorted(cpos[orderT], nh, dev)
        ext_col = loc[orderT].to(torch.int32)
        ext_val = vals[orderT].contiguous()
        if halo_ext is None or halo_ext.numel() == 0:
            self.gt_rowptr, self.gt_col, self.gt_val = ext_rowptr, ext_col, ext_val
        else:
            o2 = orderT[owned[orderT]]
            self.gt_rowptr = _csr_from_sorted(cpos[o2], nh, dev)
            self.gt_col = loc[o2].to(torch.int32)
            self.gt_val = vals[o2].contiguous()
        self._ext = (ext_rowptr, ext_col, ext_val)
        self.dg_inv = None
        self._shared = None
        if ctx is not None and ctx.world > 1:
            self._build_interface()

    # ---- multi-GPU: H1 functions reached from the rows of several ranks -------------------------
    def _build_interface(self):
        ctx, dev = self.ctx, self.h1_ids.device
        sizes = torch.zeros((ctx.world,), dtype=torch.int64, device=dev)
        sizes[ctx.rank] = self.nh
        ctx._sum(sizes)
        smax = int(sizes.max().item())
        mine = torch.full((smax,), -1, dtype=torch.int64, device=dev)
        mine[: self.nh] = self.h1_ids
        if ctx._staged:  # gloo: host tensors
            allh = torch.empty((ctx.world * smax,), dtype=torch.int64)
            ctx.dist.all_gather_into_tensor(allh, mine.cpu(), group=ctx.group)
            allids = allh.to(dev)
        else:
            allids = torch.empty((ctx.world * smax,), dtype=torch.int64, device=dev)
            ctx.dist.all_gather_into_tensor(allids, mine, group=ctx.group)
        self._shared = []   # per other rank: positions (in my list) of the H1 functions we both reach
        for r in range(ctx.world):
            if r == ctx.rank:
                self._shared.append(None)
                continue
            theirs = allids[r * smax: r * smax + int(sizes[r].item())]
            if self.nh == 0 or theirs.numel() == 0:
                self._shared.append(torch.zeros((0,), dtype=torch.int64, device=dev))
                continue
            pos = torch.searchsorted(theirs, self.h1_ids).clamp_(max=theirs.numel() - 1)
            self._shared.append(torch.nonzero(theirs[pos] == self.h1_ids).reshape(-1))
        self._splits = [0 if s is None else int(s.numel()) for s in self._shared]
        self._send_idx = torch.cat([s for s in self._shared if s is not None]) if ctx.world > 1 else None
        self._xchg = None
        if ctx.peer is not None:
            # peer transport: the partial sums are pushed into the neighbours' receive areas (laid out by
            # source rank) and added by one gather kernel: row i of `acc` lists the received copies of
            # function i in rank order -> deterministic, and the dg_inv scaling rides along
            from .peer import PeerExchange

            self._xchg = PeerExchange(ctx.peer, self._send_idx, self._splits, self._splits)
            total = int(sum(self._splits))
            rows = self._send_idx  # received copy j (source-rank major, same order as my sends) adds to row rows[j]
            order = torch.argsort(rows, stable=True)
            self._acc = (_csr_from_sorted(rows[order], self.nh, dev), order.to(torch.int32).contiguous(),
                         torch.ones((max(total, 1),), dtype=torch.float64, device=dev))
            self._recv = {}

    def _recv_area(self, k):
        if k not in self._recv:
            from .peer import SymmetricBuffer

            total = int(sum(self._splits))
            buf = SymmetricBuffer(self.ctx.peer, max(total, 1) * k * 16)
            self._recv[k] = (buf.view(_C128, max(total, 1) * k), self._xchg.target(buf, [0] * self.ctx.world, k), buf)
        return self._recv[k]

    def close(self):
        """Collective: release the peer-mapped receive areas."""
        if getattr(self, "_xchg", None) is not None:
            for _, _, buf in self._recv.values():
                buf.close()
            self._recv = {}

    def sum_over_ranks(self, part: torch.Tensor, scale: torch.Tensor = None) -> torch.Tensor:
        """part [nh, k] (or [nh]): partial sums of this rank -> totals (times `scale` [nh] if given).  The
        functions shared with other ranks travel packed: pushed over peer memory and added in rank order by
        one kernel (transport "peer"), or one all_to_all and an add per neighbour (transport "nccl")."""
        if self._shared is None:
            if scale is not None:
                check(lib().pg_zbscale_rows(self.nh, part.reshape(self.nh, -1).shape[1], ptr(scale), ptr(part),
                                            ptr(part), stream_ptr()), "pg_zbscale_rows")
            return part
        ctx = self.ctx
        flat = part.reshape(self.nh, -1)
        k = flat.shape[1]
        if self._xchg is not None:
            recv, dst, _ = self._recv_area(k)
            self._xchg.push(flat, k, dst)
            self._xchg.wait()
            rp, ci, ones = self._acc
            if scale is None:
                scale = self._ones()
            check(lib().pg_rcsr_apply(self.nh, ptr(rp), ptr(ci), ptr(ones), k, ptr(recv), ptr(scale), ptr(scale),
                                      ptr(flat), ptr(flat), stream_ptr()), "pg_rcsr_apply")
            self._xchg.ack()
            return part
        send = flat[self._send_idx].contiguous()
        recv = torch.empty_like(send)
        ctx._a2a(torch.view_as_real(recv).view(-1), torch.view_as_real(send).view(-1),
                 [2 * k * s for s in self._splits], [2 * k * s for s in self._splits])
        off = 0
        for r, idx in enumerate(self._shared):
            if idx is None or idx.numel() == 0:
                continue
            flat.index_add_(0, idx, recv[off: off + idx.numel()])
            off += idx.numel()
        if scale is not None:
            check(lib().pg_zbscale_rows(self.nh, k, ptr(scale), ptr(flat), ptr(flat), stream_ptr()), "pg_zbscale_rows")
        return part

    def _ones(self):
        if getattr(self, "_one", None) is None:
            self._one = torch.ones((max(self.nh, 1),), dtype=_C128, device=self.g_val.device)
        return self._one

    # ---- numeric setup -----------------------------------------------------------------------------
    def setup(self, A_local):
        """dg_inv = 1 / diag(G^T A G) from the matrix the Krylov operator uses (columns in the local
        [own | halo] numbering), 0 for fixed or untouched functions."""
        dev = self.g_val.device
        d = torch.zeros((max(self.nh, 1),), dtype=_C128, device=dev)[: self.nh]
        rp, ci, vv = self._ext
        check(lib().pg_galerkin_diagonal(self.nh, ptr(rp), ptr(ci), ptr(vv), self.n_own, ptr(A_local.rowptr),
                                         ptr(A_local.colidx), ptr(A_local.vals), ptr(d), stream_ptr()),
              "pg_galerkin_diagonal")
        d = self.sum_over_ranks(d)
        check(lib().pg_masked_reciprocal(self.nh, None, ptr(d), stream_ptr()), "pg_masked_reciprocal")
        self.dg_inv = d
        self._ext = None
        self._buf = {}
        return self

    def _work(self, k, dev):
        if k not in self._buf:
            self._buf[k] = torch.zeros((max(self.nh, 1), k), dtype=_C128, device=dev)[: self.nh]
        return self._buf[k]

    def apply(self, R: torch.Tensor, dinv: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
        """Z = dinv .* R + G (dg_inv .* (G^T R)) for an interleaved block R [n, k] (or a vector [n])."""
        k = 1 if R.dim() == 1 else int(R.shape[1])
        y = self._work(k, R.device)
        L = lib()
        if self._shared is None:
            check(L.pg_rcsr_apply(self.nh, ptr(self.gt_rowptr), ptr(self.gt_col), ptr(self.gt_val), k, ptr(R),
                                  ptr(self.dg_inv), None, None, ptr(y), stream_ptr()), "pg_rcsr_apply")
        else:
            check(L.pg_rcsr_apply(self.nh, ptr(self.gt_rowptr), ptr(self.gt_col), ptr(self.gt_val), k, ptr(R),
                                  None, None, None, ptr(y), stream_ptr()), "pg_rcsr_apply")
            self.sum_over_ranks(y, self.dg_inv)
        check(L.pg_rcsr_apply(self.n_own, ptr(self.g_rowptr), ptr(self.g_col), ptr(self.g_val), k, ptr(y), None,
                              ptr(dinv), ptr(R), ptr(Z), stream_ptr()), "pg_rcsr_apply")
        return Z

    def to_scipy(self):
        """G over the owned rows as scipy CSR [n_own, nh] (tests)."""
        import scipy.sparse as sp

        return sp.csr_matrix((self.g_val.cpu().numpy(), self.g_col.cpu().numpy(), self.g_rowptr.cpu().numpy()),
                             shape=(self.n_own, self.nh))

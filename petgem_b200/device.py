"""Device-side objects of the hot path: element data, assembly plan, CSR matrix.

torch is used for device memory, streams and (multi-GPU) torch.distributed only;
every computation goes through the C ABI in include/petgem_b200.h.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import basis
from ._lib import PetgemB200Error, check, lib, ptr, require_cuda, stream_ptr

MU0 = 4.0 * np.pi * 1e-7  # solver.py:175


def _dev(device=None):
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


@functools.lru_cache(maxsize=None)
def _table_host(p: int) -> np.ndarray:
    SM, SK = basis.element_tables(p)
    # [nexp, nexp, 12]: SK[0..5] then SM[0..5] for each expanded pair (J, K)
    tab = np.concatenate([np.moveaxis(SK, 0, -1), np.moveaxis(SM, 0, -1)], axis=-1)
    return np.ascontiguousarray(tab)


_TABLE_CACHE = {}


def element_table(p: int, device=None) -> torch.Tensor:
    """Reference-element contraction table of order p on the device, built once per order and device by
    pg_tables_init (basis evaluation + exact quadrature + contraction, all in the library)."""
    dev = _dev(device)
    key = (p, str(dev))
    if key not in _TABLE_CACHE:
        nexp = lib().pg_nexp(p)
        tab = torch.empty((nexp, nexp, 12), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib().pg_tables_init(p, ptr(tab), stream_ptr()), "pg_tables_init")
        _TABLE_CACHE[key] = tab
    return _TABLE_CACHE[key]


def locate_points(elems: "ElementData", points) -> torch.Tensor:
    """Containing element of each point (lowest index; -1: outside the mesh) -> int32 [npts] on the device
    (postprocessing.py:532-539 / preprocessing.py:414-420)."""
    pts = torch.as_tensor(np.ascontiguousarray(np.atleast_2d(points), dtype=np.float64)).to(elems.device)
    out = torch.empty((pts.shape[0],), dtype=torch.int32, device=elems.device)
    check(lib().pg_locate_points(elems.T, ptr(elems.nodes), pts.shape[0], ptr(pts), 1.0e-12, ptr(out), stream_ptr()),
          "pg_locate_points")
    return out


def interpolate_fields(elems: "ElementData", p: int, code: torch.Tensor, x: torch.Tensor, points, pt_elem, omega: float,
                       mu: float = MU0, perm: torch.Tensor = None) -> torch.Tensor:
    """fieldInterpolator (postprocessing.py:479-616) on the device -> [npts, 6] complex128 (E, H)."""
    pts = torch.as_tensor(np.ascontiguousarray(np.atleast_2d(points), dtype=np.float64)).to(elems.device)
    out = torch.empty((pts.shape[0], 6), dtype=torch.complex128, device=elems.device)
    check(lib().pg_interpolate_fields(pts.shape[0], ptr(pts), ptr(pt_elem), p, ptr(elems.nodes), ptr(code),
                                      ptr(elems.elemsE), ptr(elems.elemsF), elems.nEdges, elems.nFaces, ptr(perm),
                                      ptr(x), float(omega), float(mu), ptr(out), stream_ptr()), "pg_interpolate_fields")
    return out


def csem_rhs(elems: "ElementData", p: int, code: torch.Tensor, source_elem: int, position, moment, omega: float,
             b: torch.Tensor, mu: float = MU0, perm: torch.Tensor = None, row_begin: int = 0) -> torch.Tensor:
    """b += i omega mu (moment . N_j(position)) on the dofs of the source element (solver.py:247-316)."""
    pos = np.ascontiguousarray(position, dtype=np.float64)
    mom = np.ascontiguousarray(moment, dtype=np.float64)
    check(lib().pg_csem_rhs(p, int(source_elem), ptr(pos), ptr(mom), ptr(elems.nodes), ptr(code), ptr(elems.elemsE),
                            ptr(elems.elemsF), elems.nEdges, elems.nFaces, ptr(perm), int(row_begin), b.numel(),
                            float(omega), float(mu), ptr(b), stream_ptr()), "pg_csem_rhs")
    return b


class ElementData:
    """Per-element inputs of the element loop (rows of the scratch files read at
    solver.py:193-211), resident in HBM as flat SoA-of-rows arrays."""

    def __init__(self, nodes, elemsN, elemsE, edgesNodes, facesEdges, elemsF, sigma, nEdges, nFaces, device=None):
        dev = _dev(device)
        T = int(np.asarray(elemsN).shape[0])

        def f64(a, cols):
            t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(T, cols))
            return t.to(dev, non_blocking=False)

        def i32(a, cols):
            t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32).reshape(T, cols))
            return t.to(dev, non_blocking=False)

        self.T = T
        self.nodes = f64(nodes, 12)
        self.elemsN = i32(elemsN, 4)
        self.elemsE = i32(elemsE, 6)
        self.edgesNodes = i32(edgesNodes, 12)
        self.facesEdges = i32(facesEdges, 12)
        self.elemsF = i32(elemsF, 4)
        self.sigma = f64(sigma, 2)
        self.nEdges, self.nFaces = int(nEdges), int(nFaces)
        self.device = dev

    @classmethod
    def from_mesh(cls, nodes_xyz, elemsN, elemsE, edgesNodes, elemsF, facesE, sigma, device=None):
        """Build the per-element rows from global mesh tables (preprocessing.py:92-216)."""
        elemsN = np.asarray(elemsN)
        return cls(
            np.asarray(nodes_xyz)[elemsN].reshape(elemsN.shape[0], 12),
            elemsN,
            elemsE,
            np.asarray(edgesNodes)[np.asarray(elemsE)].reshape(elemsN.shape[0], 12),
            np.asarray(facesE)[np.asarray(elemsF)].reshape(elemsN.shape[0], 12),
            elemsF,
            sigma,
            np.asarray(edgesNodes).shape[0],
            np.asarray(facesE).shape[0],
            device=device,
        )

    def geometry(self, elem_range=None, out=None):
        """computeJacobian + computeElementOrientation -> (geo [T,12], code [T]).  With elem_range
        (t0, t1) only those elements are evaluated (a rank needs just the elements touching its rows);
        the arrays keep their global shape so that element ids index them directly."""
        if out is None:
            geo = torch.empty((self.T, 12), dtype=torch.float64, device=self.device)
            code = torch.empty((self.T,), dtype=torch.int32, device=self.device)
        else:
            geo, code = out
        t0, t1 = (0, self.T) if elem_range is None else elem_range
        if t1 > t0:
            check(
                lib().pg_element_geometry(
                    t1 - t0, ptr(self.nodes[t0:]), ptr(self.elemsN[t0:]), ptr(self.elemsE[t0:]),
                    ptr(self.edgesNodes[t0:]), ptr(self.facesEdges[t0:]), ptr(self.sigma[t0:]), ptr(geo[t0:]),
                    ptr(code[t0:]), stream_ptr(),
                ),
                "pg_element_geometry",
            )
        return geo, code

    def dofs(self, p: int) -> torch.Tensor:
        """computeConnectivityDOFS (hvfem.py:15-98) -> [T, n] int32."""
        n = basis.ndof_element(p)
        out = torch.empty((self.T, n), dtype=torch.int32, device=self.device)
        check(
            lib().pg_connectivity_dofs(self.T, p, ptr(self.elemsE), ptr(self.elemsF), self.nEdges, self.nFaces,
                                       ptr(out), stream_ptr()),
            "pg_connectivity_dofs",
        )
        return out


def element_matrices(p: int, geo: torch.Tensor, code: torch.Tensor):
    """Batched computeElementalMatrices (hvfem.py:223-316) -> (Me, Ke) [T,n,n] float64."""
    T, n = geo.shape[0], basis.ndof_element(p)
    Me = torch.empty((T, n, n), dtype=torch.float64, device=geo.device)
    Ke = torch.empty_like(Me)
    check(
        lib().pg_element_matrices(T, p, ptr(geo), ptr(code), ptr(element_table(p, geo.device)), ptr(Me), ptr(Ke),
                                  stream_ptr()),
        "pg_element_matrices",
    )
    return Me, Ke


def element_systems(p: int, geo: torch.Tensor, code: torch.Tensor, omega: float, mu: float = MU0):
    """Batched Ae = K - i*omega*mu*M (solver.py:223) -> [T,n,n] complex128."""
    T, n = geo.shape[0], basis.ndof_element(p)
    Ae = torch.empty((T, n, n), dtype=torch.complex128, device=geo.device)
    check(
        lib().pg_element_systems(T, p, ptr(geo), ptr(code), ptr(element_table(p, geo.device)), -omega * mu, ptr(Ae),
                                 stream_ptr()),
        "pg_element_systems",
    )
    return Ae


def morton_element_rank(nodes12: np.ndarray) -> np.ndarray:
    """Position of every element along a Morton (Z-order) curve through the element centroids.
    nodes12: [T,12] per-element vertex coordinates (the nodes.dat rows) -> int32 [T]."""
    X = np.asarray(nodes12, dtype=np.float64).reshape(-1, 4, 3).mean(axis=1)
    lo, hi = X.min(axis=0), X.max(axis=0)
    q = np.minimum(((X - lo) / np.maximum(hi - lo, 1e-300) * 1024.0).astype(np.int64), 1023)  # 10 bits per axis

    def spread(v):  # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v

    key = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    order = np.argsort(key, kind="stable")
    rank = np.empty(order.size, dtype=np.int32)
    rank[order] = np.arange(order.size, dtype=np.int32)
    return rank


class AssemblyPlan:
    """Symbolic phase (pattern, incidence lists, slot positions) for one mesh and order."""

    def __init__(self, elems: ElementData, p: int, order="reference", row_range=None, elem_rank=None):
        """order: 'reference' (PETGEM numbering), 'locality' (element-major: entities by first incident
        element; with elem_rank [T] = position of each element in a space-filling traversal, by first
        incident element along that curve -- see morton_element_rank), or an explicit entity order."""
        self.elems, self.p = elems, p
        self.n = basis.ndof_element(p)
        nEnt = elems.nEdges + (elems.nFaces if p >= 2 else 0) + (elems.T if p >= 3 else 0)
        self.nEnt = nEnt
        if isinstance(order, str):
            if order == "reference":
                order_host = None
            elif order == "locality":
                order_host = np.empty(nEnt, dtype=np.int32)
                rank_dev = None
                if elem_rank is not None:
                    rank_dev = torch.as_tensor(np.ascontiguousarray(elem_rank, dtype=np.int32)).to(elems.device)
                check(
                    lib().pg_plan_ranked_order(elems.T, p, ptr(elems.elemsE), ptr(elems.elemsF), elems.nEdges,
                                               elems.nFaces, ptr(rank_dev), ptr(order_host), stream_ptr()),
                    "pg_plan_ranked_order",
                )
            else:
                raise ValueError("order must be 'reference', 'locality' or an int32 array")
        else:
            order_host = np.ascontiguousarray(order, dtype=np.int32)
            if order_host.shape != (nEnt,):
                raise ValueError("entity order must have %d entries" % nEnt)
        self.order_host = order_host
        rb, re_ = (0, -1) if row_range is None else row_range
        handle = C.c_void_p()
        check(
            lib().pg_plan_create(elems.T, p, ptr(elems.elemsE), ptr(elems.elemsF), elems.nEdges, elems.nFaces,
                                 ptr(order_host), rb, re_, C.byref(handle), stream_ptr()),
            "pg_plan_create",
        )
        self._h = handle
        L = lib()
        self.N = L.pg_plan_num_dofs(handle)
        self.local_rows = L.pg_plan_local_rows(handle)
        self.row_begin = L.pg_plan_row_begin(handle)
        self.nnz = L.pg_plan_nnz(handle)
        self.contributions = L.pg_plan_contributions(handle)
        self.max_row_length = L.pg_plan_max_row_length(handle)
        t0, t1 = C.c_int64(), C.c_int64()
        check(L.pg_plan_element_range(handle, C.byref(t0), C.byref(t1)), "pg_plan_element_range")
        self.element_range = (int(t0.value), int(t1.value))
        self._csr = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().pg_plan_destroy(h)
            except Exception:
                pass

    def csr(self):
        """(rowptr int64 [rows+1], colidx int32 [nnz]) of the owned rows."""
        if self._csr is None:
            dev = self.elems.device
            rowptr = torch.empty((self.local_rows + 1,), dtype=torch.int64, device=dev)
            colidx = torch.empty((max(self.nnz, 1),), dtype=torch.int32, device=dev)[: self.nnz]
            check(lib().pg_plan_csr(self._h, ptr(rowptr), ptr(colidx), stream_ptr()), "pg_plan_csr")
            self._csr = (rowptr, colidx)
        return self._csr

    def column_starts(self) -> torch.Tensor:
        """First column (numbering in use) of every column entity of every entity's list."""
        n = lib().pg_plan_num_column_entities(self._h)
        out = torch.empty((max(n, 1),), dtype=torch.int32, device=self.elems.device)[:n]
        check(lib().pg_plan_column_starts(self._h, ptr(out), stream_ptr()), "pg_plan_column_starts")
        return out

    def dof_permutation(self) -> torch.Tensor:
        """perm[ref_dof] = index in the numbering in use."""
        perm = torch.empty((self.N,), dtype=torch.int32, device=self.elems.device)
        check(lib().pg_plan_dof_permutation(self._h, ptr(perm), stream_ptr()), "pg_plan_dof_permutation")
        return perm

    def entity_aligned_row(self, row: int) -> int:
        return int(lib().pg_plan_entity_aligned_row(self._h, int(row)))

    def set_dirichlet(self, bd_entity):
        """bd_entity: [nEnt] uint8 (device tensor or host array) or None."""
        if bd_entity is None:
            check(lib().pg_plan_set_dirichlet(self._h, None, None, None, stream_ptr()), "pg_plan_set_dirichlet")
            self.bd_entity = None
            return
        t = torch.as_tensor(bd_entity, dtype=torch.uint8).to(self.elems.device).contiguous()
        if t.numel() != self.nEnt:
            raise ValueError("bd_entity must have %d entries" % self.nEnt)
        self.bd_entity = t
        check(
            lib().pg_plan_set_dirichlet(self._h, ptr(self.elems.elemsE), ptr(self.elems.elemsF), ptr(t), stream_ptr()),
            "pg_plan_set_dirichlet",
        )

    def dirichlet_row_mask(self):
        """uint8 [N] in the numbering in use: rows of the Dirichlet entities given to set_dirichlet
        (all dofs of boundary edges and faces, mesh.py:280-321); None without Dirichlet entities."""
        bd = getattr(self, "bd_entity", None)
        if bd is None:
            return None
        el, p = self.elems, self.p
        nE, nF, nf = el.nEdges, el.nFaces, p * (p - 1)
        ref = torch.zeros((self.N,), dtype=torch.uint8, device=el.device)
        ref[: nE * p] = bd[:nE].repeat_interleave(p)
        if p >= 2:
            ref[nE * p: nE * p + nF * nf] = bd[nE: nE + nF].repeat_interleave(nf)
        if self.order_host is None:
            return ref
        out = torch.empty_like(ref)
        out[self.dof_permutation().to(torch.int64)] = ref
        return out

    def assemble(self, geo, code, omega, mu=MU0, apply_dirichlet=False, diag=1.0, out=None):
        """Numeric phase -> vals [nnz] complex128."""
        dev = self.elems.device
        self.dirichlet_applied = bool(apply_dirichlet)
        vals = out if out is not None else torch.empty((max(self.nnz, 1),), dtype=torch.complex128, device=dev)[: self.nnz]
        check(
            lib().pg_assemble(self._h, ptr(geo), ptr(code), ptr(element_table(self.p, dev)), -omega * mu,
                              1 if apply_dirichlet else 0, float(diag), ptr(vals), stream_ptr()),
            "pg_assemble",
        )
        return vals


class CSRMatrix:
    """Complex128 CSR block of owned rows [row_begin, row_begin+rows) x N columns."""

    def __init__(self, rowptr, colidx, vals, N, row_begin=0, plan=None, colstart=None, blocked=None):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.N, self.row_begin = int(N), int(row_begin)
        self.rows = int(rowptr.numel() - 1)
        self.nnz = int(vals.numel())
        # with the plan that produced vals (p = 2) MatMult uses the entity-blocked kernel: the plan's
        # per-entity column lists replace colidx (17 B per nonzero instead of 20, half the x gathers).
        # Measured on B200 at C3 (tools/spmv_bench.py) with the in-order grid: 5.19 ms against 5.96 ms for
        # the CSR kernel.  blocked=False or PG_SPMV_BLOCKED=0 selects the CSR kernel.
        import os
        if blocked is None:
            blocked = os.environ.get("PG_SPMV_BLOCKED", "1") == "1"
        self.plan_ref = plan if (plan is not None and plan.p == 2) else None  # entity blocks of this matrix
        self.asm_plan = plan  # mesh + numbering behind this matrix (gradient-space preconditioner)
        # rows fixed by MatZeroRowsColumns, uint8 [N] in the numbering in use (None: no Dirichlet rows)
        self.dirichlet_mask = plan.dirichlet_row_mask() if (plan is not None and getattr(plan, "dirichlet_applied",
                                                                                         False)) else None
        self.plan = self.plan_ref if blocked else None
        self.colstart = colstart  # None = the plan's own global column starts

    def halo_split(self, n_own: int):
        """(ent_begin, ent_end, n_entities): the entities (plan processing order) in [ent_begin, ent_end) touch no
        column >= n_own, i.e. no halo entry of an [own | halo] vector; None without entity blocks."""
        if self.plan is None or self.colstart is None:
            return None
        import ctypes as C
        e0, e1 = C.c_int64(0), C.c_int64(0)
        check(lib().pg_plan_halo_split(self.plan._h, ptr(self.colstart), int(n_own), C.byref(e0), C.byref(e1),
                                       stream_ptr()), "pg_plan_halo_split")
        return int(e0.value), int(e1.value), self.rows // 2  # p = 2: two rows per owned entity

    def mult(self, x: torch.Tensor, y: torch.Tensor = None, row_scale: torch.Tensor = None, ent_range=None) -> torch.Tensor:
        """y = A x  (MatMult); x has N entries, y the owned rows.  With row_scale, y = row_scale .* (A x)
        (the Jacobi preconditioner applied in the SpMV epilogue).  ent_range = (a, b): only the rows of the
        entities [a, b) of the plan's processing order (entity-blocked matrices)."""
        if y is None:
            y = torch.empty((self.rows,), dtype=torch.complex128, device=x.device)
        if self.plan is not None:
            if ent_range is not None:
                check(lib().pg_spmv_blocked_range(self.plan._h, int(ent_range[0]), int(ent_range[1]), ptr(self.colstart),
                                                  ptr(self.vals), ptr(x), ptr(row_scale), ptr(y), stream_ptr()),
                      "pg_spmv_blocked_range")
                return y
            check(
                lib().pg_spmv_blocked(self.plan._h, ptr(self.colstart), ptr(self.vals), ptr(x), ptr(row_scale),
                                      ptr(y), stream_ptr()),
                "pg_spmv_blocked",
            )
            return y
        if ent_range is not None:
            raise PetgemB200Error("mult: ent_range needs the entity-blocked form")
        check(
            lib().pg_spmv_scaled(self.rows, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(x),
                                 ptr(row_scale), ptr(y), stream_ptr()),
            "pg_spmv",
        )
        return y

    def mult_fused_dot(self, X: torch.Tensor, Y: torch.Tensor, k: int, out: torch.Tensor) -> bool:
        """Y = A X (k = 1: vectors, else [n, k] blocks) and out[r] = sum_rows X[row, r] Y[row, r] (unconjugated) in
        ONE pass (pg_spmm_blocked_dot): the x^T A x of COCG / COCR without re-reading x and A x.  Returns False
        when this matrix has no fused kernel for the shape (no entity blocks, k = 2, unaligned views): the
        caller then multiplies and takes the dot product separately."""
        if self.plan is None or k not in (1, 4, 8) or (X.data_ptr() | self.vals.data_ptr()) & 31:
            return False
        if not hasattr(self, "_dot_work"):
            self._dot_work = {}
        w = self._dot_work.get(k)
        if w is None:
            nbytes = lib().pg_spmv_dot_workspace_bytes(self.plan._h, k)
            w = self._dot_work[k] = torch.empty((nbytes // 16,), dtype=torch.complex128, device=self.vals.device)
        check(
            lib().pg_spmm_blocked_dot(self.plan._h, ptr(self.colstart), ptr(self.vals), k, ptr(X), None, ptr(Y), ptr(out),
                                      ptr(w), stream_ptr()),
            "pg_spmm_blocked_dot",
        )
        return True

    def mult_multi(self, X: torch.Tensor, Y: torch.Tensor = None, row_scale: torch.Tensor = None, ent_range=None) -> torch.Tensor:
        """Y = A X for k interleaved right-hand sides: X is [N, k], Y [rows, k] (C-contiguous), k in
        {1, 2, 4, 8}.  The matrix is streamed once for all k (several sources / MT polarizations)."""
        k = int(X.shape[1])
        if Y is None:
            Y = torch.empty((self.rows, k), dtype=torch.complex128, device=X.device)
        if not (X.is_contiguous() and Y.is_contiguous()):
            raise PetgemB200Error("mult_multi: X and Y must be C-contiguous [n, k] blocks")
        if ent_range is not None:
            if self.plan_ref is None or k not in (2, 4, 8):
                raise PetgemB200Error("mult_multi: ent_range needs the entity-blocked form and k in (2, 4, 8)")
            check(lib().pg_spmm_blocked_range(self.plan_ref._h, int(ent_range[0]), int(ent_range[1]), ptr(self.colstart),
                                              ptr(self.vals), k, ptr(X), ptr(row_scale), ptr(Y), stream_ptr()),
                  "pg_spmm_blocked_range")
            return Y
        if self.plan_ref is not None and k in (2, 4, 8):
            # p = 2: the plan's 2x2 entity blocks halve the gathers (measured at C3, k = 4: 10.6 ms against
            # 18.0 ms for the CSR form and 4 x 5.9 ms for four single passes)
            check(
                lib().pg_spmm_blocked(self.plan_ref._h, ptr(self.colstart), ptr(self.vals), k, ptr(X), ptr(row_scale),
                                      ptr(Y), stream_ptr()),
                "pg_spmm_blocked",
            )
            return Y
        check(
            lib().pg_spmm(self.rows, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), k, ptr(X), ptr(row_scale),
                          ptr(Y), stream_ptr()),
            "pg_spmm",
        )
        return Y

    def diagonal(self) -> torch.Tensor:
        d = torch.empty((self.rows,), dtype=torch.complex128, device=self.vals.device)
        check(
            lib().pg_csr_diagonal(self.rows, self.row_begin, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(d),
                                  stream_ptr()),
            "pg_csr_diagonal",
        )
        return d

    def zeroRowsColumns(self, bd_rows, diag: float = 1.0):
        """A.zeroRowsColumns(rows) (solver.py:562): rows given as global dof ids."""
        mask = torch.zeros((self.N,), dtype=torch.uint8, device=self.vals.device)
        idx = torch.as_tensor(np.asarray(bd_rows, dtype=np.int64), device=self.vals.device)
        mask[idx] = 1
        self.dirichlet_mask = mask if self.dirichlet_mask is None else (self.dirichlet_mask | mask)
        check(
            lib().pg_zero_rows_columns(self.rows, self.row_begin, ptr(self.rowptr), ptr(self.colidx), ptr(mask),
                                       float(diag), ptr(self.vals), stream_ptr()),
            "pg_zero_rows_columns",
        )

    def spmv_bytes(self) -> int:
        """Algorithmic bytes of one SpMV: 20 nnz + 40 rows."""
        return 20 * self.nnz + 40 * self.rows

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix(
            (self.vals.cpu().numpy(), self.colidx.cpu().numpy(), self.rowptr.cpu().numpy()), shape=(self.rows, self.N)
        )


__all__ = [
    "ElementData", "AssemblyPlan", "CSRMatrix", "element_table", "element_matrices", "element_systems", "MU0",
    "locate_points", "interpolate_fields", "csem_rhs",
    "PetgemB200Error",
]

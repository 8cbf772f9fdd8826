"""Host logic of the gradient-space preconditioner (petgem_b200/gradient.py) on the CPU: the discrete
gradient G built from the mesh tables must lie in the null space of the curl-curl matrix assembled by the
oracle (K G = 0), for the reference numbering and for a permuted one, with Dirichlet rows and with row
blocks + halo (two ranks over gloo).  The CUDA kernels that apply G are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeElems:
    """CPU stand-in for device.ElementData: the attributes GradientSpace reads."""

    def __init__(self, tab):
        from petgem_b200 import hvfem

        elemsN, elemsE, elemsF = tab["elemsN"], tab["elemsE"], tab["elemsF"]
        self.T = elemsN.shape[0]
        self.nEdges, self.nFaces = tab["edgesNodes"].shape[0], tab["facesE"].shape[0]
        self.device = torch.device("cpu")
        self.elemsN = torch.as_tensor(elemsN.astype(np.int32))
        self.elemsE = torch.as_tensor(elemsE.astype(np.int32))
        self.elemsF = torch.as_tensor(elemsF.astype(np.int32))
        self.edgesNodes = torch.as_tensor(tab["edgesNodes"][elemsE].reshape(self.T, 12).astype(np.int32))
        eo, fo = hvfem.computeElementOrientation_batch(elemsE, elemsN, tab["edgesNodes"][elemsE], tab["facesE"][elemsF])
        self._code = torch.as_tensor(hvfem.pack_orientation(eo, fo).astype(np.int32))

    def geometry(self):
        return None, self._code


class _FakePlan:
    def __init__(self, elems, p, N, perm=None, row_range=None):
        self.elems, self.p, self.N = elems, p, N
        self.order_host = None if perm is None else perm
        self._perm = perm
        self.row_begin, row_end = (0, N) if row_range is None else row_range
        self.local_rows = row_end - self.row_begin

    def dof_permutation(self):
        return torch.as_tensor(self._perm.astype(np.int32))


def _curl_curl(oracle, tab, p):
    """K (omega = 0) of the small mesh with the oracle, reference numbering."""
    T = tab["elemsN"].shape[0]
    n = p * (p + 2) * (p + 3) // 2
    sigma = np.ones((T, 2))
    Ae = np.zeros((T, n, n), dtype=np.complex128)
    for t in range(T):
        Ae[t] = oracle.element_system(tab["nodes"][tab["elemsN"][t]], tab["elemsN"][t], tab["elemsE"][t],
                                      tab["edgesNodes"][tab["elemsE"][t]], tab["facesE"][tab["elemsF"][t]], sigma[t], p,
                                      0.0, 1.0)
    dofs, *_, N = oracle.compute_connectivity_dofs(tab["elemsE"], tab["elemsF"], p)
    rp, ci, v = oracle.assemble_global(Ae, dofs, N)
    return oracle.to_scipy(rp, ci, v).real.tocsr(), N


def _mesh(m=2):
    from petgem_b200 import synthetic

    nodes, elemsN = synthetic.kuhn_box(m, length=100.0 * m)
    return synthetic.mesh_tables(nodes, elemsN)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_gradient_is_in_the_null_space_of_curl_curl(oracle, p):
    from petgem_b200.gradient import GradientSpace

    tab = _mesh(2)
    K, N = _curl_curl(oracle, tab, p)
    el = _FakeElems(tab)
    gs = GradientSpace(_FakePlan(el, p, N))
    G = gs.to_scipy()
    nn = tab["nodes"].shape[0]
    assert G.shape == (N, nn + (el.nEdges if p >= 2 else 0))
    assert abs(K @ G).max() <= 1e-12 * abs(K).max()
    # full column rank up to the constants: rank(G) = (#H1 functions) - 1
    assert np.linalg.matrix_rank(G.toarray()) == G.shape[1] - 1
    # permuted numbering: the rows move with the dofs
    perm = np.random.default_rng(5).permutation(N)
    Gp = GradientSpace(_FakePlan(el, p, N, perm=perm)).to_scipy()
    assert abs(Gp[perm] - G).max() == 0.0


def test_dirichlet_rows_and_columns_are_dropped(oracle):
    from petgem_b200.gradient import GradientSpace

    p = 2
    tab = _mesh(2)
    el = _FakeElems(tab)
    nE, nF = el.nEdges, el.nFaces
    N = p * nE + p * (p - 1) * nF
    fixed = np.zeros(N, dtype=bool)
    fixed[(tab["bEdges"][:, None] * p + np.arange(p)).ravel()] = True
    fixed[(nE * p + tab["bFaces"][:, None] * 2 + np.arange(2)).ravel()] = True
    gs = GradientSpace(_FakePlan(el, p, N), dirichlet_rows=torch.as_tensor(fixed))
    G = sp.csr_matrix((gs.g_val.numpy(), gs.h1_ids.numpy()[gs.g_col.numpy()], gs.g_rowptr.numpy()),
                      shape=(N, gs.n_h1))
    assert abs(G[np.nonzero(fixed)[0]]).sum() == 0.0
    nn = tab["nodes"].shape[0]
    bnodes = np.unique(tab["edgesNodes"][tab["bEdges"]])
    h1_fixed = np.zeros(gs.n_h1, dtype=bool)
    h1_fixed[bnodes] = True
    h1_fixed[nn + tab["bEdges"]] = True
    assert np.array_equal(gs.h1_fixed.numpy(), h1_fixed)
    assert abs(G[:, np.nonzero(h1_fixed)[0]]).sum() == 0.0
    # interior functions keep their full support
    full = GradientSpace(_FakePlan(el, p, N)).to_scipy()
    free = np.nonzero(~h1_fixed)[0]
    assert abs(G[:, free] - sp.diags((~fixed).astype(float)) @ full[:, free]).max() == 0.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import petgem_oracle as oracle

        from petgem_b200.gradient import GradientSpace
        from petgem_b200.krylov import DistContext

        p = 2
        tab = _mesh(2)
        K, N = _curl_curl(oracle, tab, p)
        rng = np.random.default_rng(11)
        A = (K + sp.diags(rng.uniform(1.0, 2.0, size=N))).astype(np.complex128).tocsr()  # any matrix with K's pattern
        el = _FakeElems(tab)
        cut = 2 * (N // 4)
        cuts = [0, cut, N]
        ctx = DistContext(cuts[:-1], N)
        lo, hi = cuts[rank], cuts[rank + 1]
        Aloc = A[lo:hi]
        ctx.build_halo(torch.from_numpy(Aloc.indices.astype(np.int32)))
        gs = GradientSpace(_FakePlan(el, p, N, row_range=(lo, hi)), ctx=ctx, halo_ext=ctx._halo_ext)
        Gfull = GradientSpace(_FakePlan(el, p, N)).to_scipy()
        # owned rows of G in the local H1 numbering
        Gl = sp.csr_matrix((gs.g_val.numpy(), gs.h1_ids.numpy()[gs.g_col.numpy()], gs.g_rowptr.numpy()),
                           shape=(hi - lo, gs.n_h1))
        assert abs(Gl - Gfull[lo:hi]).max() == 0.0
        # partial Galerkin diagonal (numpy restatement of pg_galerkin_diagonal) summed over the ranks
        ext_rp, ext_ci, ext_v = (t.numpy() for t in gs._ext)
        halo = ctx._halo_ext.numpy()
        glob = np.concatenate([np.arange(lo, hi), halo])
        d = np.zeros(gs.nh, dtype=np.complex128)
        for k in range(gs.nh):
            idx, g = glob[ext_ci[ext_rp[k]:ext_rp[k + 1]]], ext_v[ext_rp[k]:ext_rp[k + 1]]
            own = (idx >= lo) & (idx < hi)
            d[k] = g[own] @ (A[idx[own]][:, idx] @ g)
        tot = gs.sum_over_ranks(torch.from_numpy(d)).numpy()
        ref = (Gfull.T @ A @ Gfull).diagonal()[gs.h1_ids.numpy()]
        assert np.abs(tot - ref).max() <= 1e-12 * np.abs(ref).max()
        # G^T r: partial products summed over the ranks equal the global product on the functions reached
        r = rng.normal(size=N) + 1j * rng.normal(size=N)
        part = torch.from_numpy(np.asarray(Gl.T @ r[lo:hi])[gs.h1_ids.numpy()].reshape(-1, 1).copy())
        tot = gs.sum_over_ranks(part).numpy()[:, 0]
        assert np.abs(tot - (Gfull.T @ r)[gs.h1_ids.numpy()]).max() <= 1e-12 * np.abs(r).max() * 20
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_row_blocks_and_interface_sums_world2():
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}

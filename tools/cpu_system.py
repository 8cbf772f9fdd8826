"""Host-side prototyping helper (NOT on the product path): builds the global system of a small
mesh with numpy from the reference-element tables, for experiments with preconditioners and to
time direct solves used by the parity tests.  Same arithmetic as the kernels:
Ae = sum_c gK[c] SK[c][J,K] s s - i omega mu sum_c gM[c] SM[c][J,K] s s.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from petgem_b200 import basis, hvfem  # noqa: E402

OMEGA, MU = 2 * np.pi * 2.0, 4e-7 * np.pi


def build_system(tab, sigma, p, omega=OMEGA, mu=MU, dirichlet=True, chunk=4096):
    nodes, elemsN = tab["nodes"], tab["elemsN"]
    elemsE, elemsF = tab["elemsE"], tab["elemsF"]
    T = elemsN.shape[0]
    nE, nF = tab["edgesNodes"].shape[0], tab["facesE"].shape[0]
    eo, fo = hvfem.computeElementOrientation_batch(elemsE, elemsN, tab["edgesNodes"][elemsE], tab["facesE"][elemsF])
    X = nodes[elemsN]
    J = X[:, 1:] - X[:, :1]
    geo = hvfem.geometric_factors(J, sigma)
    SM, SK = basis.element_tables(p)
    Jx, S = basis.local_to_expanded(p, eo, fo)
    dofs = hvfem.dofs_of_elements(elemsE, elemsF, np.arange(T), nE, nF, p)
    n = dofs.shape[1]
    N = p * nE + p * (p - 1) * nF + (p * (p - 1) * (p - 2) // 2) * T
    rows, cols, vals = [], [], []
    for t0 in range(0, T, chunk):
        sl = slice(t0, min(T, t0 + chunk))
        jx, s, g = Jx[sl], S[sl], geo[sl]
        K = np.zeros((jx.shape[0], n, n))
        M = np.zeros_like(K)
        for c in range(6):
            K += g[:, c, None, None] * SK[c][jx[:, :, None], jx[:, None, :]]
            M += g[:, 6 + c, None, None] * SM[c][jx[:, :, None], jx[:, None, :]]
        ss = s[:, :, None] * s[:, None, :]
        Ae = (K - 1j * omega * mu * M) * ss
        d = dofs[sl]
        rows.append(np.repeat(d, n, axis=1).ravel())
        cols.append(np.tile(d, (1, n)).ravel())
        vals.append(Ae.ravel())
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()
    bd = None
    if dirichlet:
        bdm = np.zeros(N, dtype=bool)
        be = np.asarray(tab["bEdges"])
        bdm[(be[:, None] * p + np.arange(p)).ravel()] = True
        if p >= 2:
            bf = np.asarray(tab["bFaces"])
            nf = p * (p - 1)
            bdm[(nE * p + bf[:, None] * nf + np.arange(nf)).ravel()] = True
        keep = sp.diags((~bdm).astype(float))
        A = (keep @ A @ keep + sp.diags(bdm.astype(complex))).tocsr()
        bd = np.nonzero(bdm)[0]
    return A, dofs, bd, dict(eo=eo, fo=fo, N=N, nE=nE, nF=nF)


def csem_rhs(tab, dofs, p, N, src, omega=OMEGA, mu=MU, bd=None):
    nodes, elemsN = tab["nodes"], tab["elemsN"]
    cen = nodes[elemsN].mean(axis=1)
    te = int(np.argmin(((cen - src) ** 2).sum(axis=1)))
    Xe = nodes[elemsN[te]]
    J, Ji = hvfem.computeJacobian(Xe)
    eo, fo = hvfem.computeElementOrientation(tab["elemsE"][te], elemsN[te], tab["edgesNodes"][tab["elemsE"][te]],
                                             tab["facesE"][tab["elemsF"][te]])
    b_, _ = hvfem.computeBasisFunctions(eo, fo, J, Ji, p, np.array([0.25, 0.25, 0.25]))
    b = np.zeros(N, dtype=complex)
    b[dofs[te]] = 1j * omega * mu * (np.array([1.0, 0.0, 0.0]) @ b_[:, :, 0])
    if bd is not None:
        b[bd] = 0
    return b


def test_mesh_case(p):
    z = dict(np.load(os.path.join(ROOT, "tests", "golden", "test_mesh_topology.npz")))
    sig = np.array([1.0, 0.01, 1.0, 3.3333])[z["tags"] - 1]
    A, dofs, bd, info = build_system(z, np.stack([sig, sig], 1), p)
    b = csem_rhs(z, dofs, p, info["N"], np.array([1750.0, 1750.0, -975.0]), bd=bd)
    return z, A, b, dofs, bd, info


def kuhn_case(m, p, length=3500.0):
    from petgem_b200 import synthetic

    nodes, elemsN = synthetic.kuhn_box(m, length=length)
    tab = synthetic.mesh_tables(nodes, elemsN)
    sigma = synthetic.layered_sigma(nodes, elemsN)
    A, dofs, bd, info = build_system(tab, sigma, p)
    b = csem_rhs(tab, dofs, p, info["N"], np.array([0.5, 0.5, -0.28]) * length, bd=bd)
    return tab, A, b, dofs, bd, info


if __name__ == "__main__":
    import time

    import scipy.sparse.linalg as spla

    p = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    t0 = time.time()
    z, A, b, dofs, bd, info = test_mesh_case(p)
    print("p=%d N=%d nnz=%d build %.1fs" % (p, A.shape[0], A.nnz, time.time() - t0))
    t0 = time.time()
    lu = spla.splu(A.tocsc())
    x = lu.solve(b)
    print("splu %.1fs  |r|/|b| = %.2e" % (time.time() - t0, np.linalg.norm(A @ x - b) / np.linalg.norm(b)))

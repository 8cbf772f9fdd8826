"""Preconditioner prototypes on the CPU (NOT product code): iteration counts of COCR on a small
synthetic box for Jacobi, entity-block Jacobi and the gradient-space (auxiliary nodal) correction.
usage: python tools/pc_proto.py m p
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tools.cpu_system import OMEGA, MU, kuhn_case, test_mesh_case  # noqa: E402
from petgem_b200 import basis  # noqa: E402


def cocr(A, b, M, rtol=1e-8, maxit=20000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    bn = np.linalg.norm(z)
    p = z.copy()
    Az = A @ z
    Ap = Az.copy()
    rho = z @ Az
    for it in range(1, maxit + 1):
        MAp = M(Ap)
        alpha = rho / (Ap @ MAp)
        x += alpha * p
        z -= alpha * MAp
        if np.linalg.norm(z) <= rtol * bn:
            return x, it
        Az = A @ z
        rho_new = z @ Az
        beta = rho_new / rho
        rho = rho_new
        p = z + beta * p
        Ap = Az + beta * Ap
    return x, maxit


def block_jacobi(A, info, p):
    """inverse of the entity diagonal blocks (edges p, faces p(p-1), interiors) as a sparse matrix."""
    N, nE, nF = info["N"], info["nE"], info["nF"]
    sizes = [(0, nE, p)]
    if p >= 2:
        sizes.append((nE * p, nF, p * (p - 1)))
    if p >= 3:
        T = (N - nE * p - nF * p * (p - 1)) // (p * (p - 1) * (p - 2) // 2)
        sizes.append((nE * p + nF * p * (p - 1), T, p * (p - 1) * (p - 2) // 2))
    blocks = []
    A = A.tocsr()
    for off, cnt, r in sizes:
        idx = off + np.arange(cnt * r).reshape(cnt, r)
        B = np.zeros((cnt, r, r), dtype=complex)
        for i in range(r):
            for j in range(r):
                B[:, i, j] = np.asarray(A[idx[:, i], idx[:, j]]).ravel()
        Bi = np.linalg.inv(B)
        rows = np.repeat(idx, r, axis=1).ravel()
        cols = np.tile(idx, (1, r)).ravel()
        blocks.append(sp.coo_matrix((Bi.ravel(), (rows, cols)), shape=(N, N)))
    return sum(blocks).tocsr()


def h1_grads(p, pts):
    """gradients of the hierarchical H1 functions of order <= p on the master tet: vertices, edge bubbles
    (p>=2) -> [nfun, npts, 3] (only what p<=2 needs)."""
    lam = np.stack([1 - pts.sum(1), pts[:, 0], pts[:, 1], pts[:, 2]])
    gl = basis.GRAD_LAMBDA
    out = [np.broadcast_to(gl[i], (pts.shape[0], 3)) for i in range(4)]
    if p >= 2:
        for a, b in basis.LOCAL_EDGES:
            out.append(lam[a][:, None] * gl[b] + lam[b][:, None] * gl[a])
    return np.stack(out)


def discrete_gradient(tab, dofs, info, p, bd):
    """G [N, Nh1]: grad psi_k = sum_j G[j,k] N_j, built element by element (set, not add)."""
    eo, fo = info["eo"], info["fo"]
    T = dofs.shape[0]
    nn = tab["nodes"].shape[0]
    nE = info["nE"]
    rng = np.random.default_rng(0)
    pts = rng.dirichlet(np.ones(4), size=40)[:, 1:]
    Nx, _ = basis.evaluate_expanded(p, pts)
    Gh = h1_grads(p, pts)  # [nh, npts, 3]
    nh = Gh.shape[0]
    codes = np.concatenate([eo, fo], axis=1)
    uniq, inv = np.unique(codes, axis=0, return_inverse=True)
    Cs = np.zeros((uniq.shape[0], dofs.shape[1], nh))
    for ci, cd in enumerate(uniq):
        J, S = basis.local_to_expanded(p, cd[:6], cd[6:])
        B = (Nx[J] * S[:, None, None]).reshape(J.size, -1).T  # [npts*3, n]
        R = Gh.reshape(nh, -1).T
        C, res, *_ = np.linalg.lstsq(B, R, rcond=None)
        assert np.abs(B @ C - R).max() < 1e-10
        Cs[ci] = C
    h1 = [tab["elemsN"]]
    if p >= 2:
        h1.append(nn + tab["elemsE"])
    h1 = np.concatenate(h1, axis=1)
    Nh = nn + (nE if p >= 2 else 0)
    C_all = Cs[inv.ravel()]  # [T, n, nh]
    rows = np.repeat(dofs, nh, axis=1).ravel()
    cols = np.tile(h1, (1, dofs.shape[1])).ravel()
    v = C_all.ravel()
    nz = np.abs(v) > 1e-12
    key = rows[nz].astype(np.int64) * Nh + cols[nz]
    uk, first = np.unique(key, return_index=True)
    G = sp.coo_matrix((v[nz][first], (uk // Nh, uk % Nh)), shape=(info["N"], Nh)).tocsr()
    # Dirichlet: rows of boundary dofs removed, and H1 dofs on the boundary fixed
    bdm = np.zeros(info["N"], dtype=bool)
    bdm[bd] = True
    touched = np.asarray(abs(G[bd]).sum(axis=0)).ravel() > 0
    G = sp.diags((~bdm).astype(float)) @ G @ sp.diags((~touched).astype(float))
    return G.tocsr(), ~touched


def chebyshev(Ag, dinv, lmax, lmin, degree):
    """fixed-degree Chebyshev iteration for Ag y = r (Jacobi scaled), linear in r."""
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)

    def apply(r):
        y = np.zeros_like(r)
        res = r.copy()
        sigma = theta / delta
        rho_old = 1.0 / sigma
        d = dinv * res / theta
        for k in range(degree):
            y = y + d
            res = r - Ag @ y
            rho = 1.0 / (2 * sigma - rho_old)
            d = rho * rho_old * d + 2 * rho / delta * (dinv * res)
            rho_old = rho
        return y
    return apply


def main():
    m, p = int(sys.argv[1]), int(sys.argv[2])
    t0 = time.time()
    if m == 0:
        tab, A, b, dofs, bd, info = test_mesh_case(p)
    else:
        h = float(sys.argv[3]) if len(sys.argv) > 3 else 3500.0 / m
        tab, A, b, dofs, bd, info = kuhn_case(m, p, length=h * m)
    N = A.shape[0]
    print("m=%d p=%d N=%d nnz=%d (%.1fs)" % (m, p, N, A.nnz, time.time() - t0), flush=True)
    dinv = 1.0 / A.diagonal()
    res = {}
    t0 = time.time()
    _, res["jacobi"] = cocr(A, b, lambda r: dinv * r)
    print("jacobi", res["jacobi"], "%.1fs" % (time.time() - t0), flush=True)
    Binv = block_jacobi(A, info, p) if p >= 2 else sp.diags(dinv)
    if p >= 2:
        _, res["bjacobi"] = cocr(A, b, lambda r: Binv @ r)
        print("block jacobi", res["bjacobi"], flush=True)
    for pg in range(1, min(p, 2) + 1):
        G, free = discrete_gradient(tab, dofs, info, pg if p == pg else p, bd) if pg == p else (None, None)
        if G is None:
            # lower-order gradient space inside the order-p space: take the vertex columns only
            Gf, freef = discrete_gradient(tab, dofs, info, p, bd)
            nn = tab["nodes"].shape[0]
            G, free = Gf[:, :nn], freef[:nn]
        K = (A + A.conj()) / 2  # curl-curl part is real, mass part imaginary
        print("   pg=%d  |K G| / |K| = %.2e" % (pg, abs(K.real @ G).max() / abs(K.real).max()), flush=True)
        Ag = (G.T @ A @ G).tocsr()
        fr = np.nonzero(free)[0]
        Agf = Ag[fr][:, fr].tocsc()
        lu = spla.splu(Agf)

        def M_exact(r, Binv=Binv, G=G, lu=lu, fr=fr):
            y = np.zeros(G.shape[1], dtype=complex)
            y[fr] = lu.solve((G.T @ r)[fr])
            return Binv @ r + G @ y
        _, its = cocr(A, b, M_exact)
        print("block jacobi + grad(P%d) exact nodal solve:" % pg, its, flush=True)
        if pg == 2:
            # hybrid: Jacobi on the whole P2 gradient space + exact / Chebyshev solve on its vertex block
            nn = tab["nodes"].shape[0]
            dg2 = np.zeros(G.shape[1], dtype=complex)
            dg2[fr] = 1.0 / Agf.diagonal()
            frv = fr[fr < nn]
            Avv = Ag[frv][:, frv].tocsc()
            luv = spla.splu(Avv)

            def M_hyb(r, Binv=Binv, G=G):
                g = G.T @ r
                y = dg2 * g
                y[frv] += luv.solve(g[frv])
                return Binv @ r + G @ y
            _, its = cocr(A, b, M_hyb)
            print("   hybrid: jacobi on grad(P2) + exact vertex-block solve:", its, flush=True)

            def M_hyb2(r, Binv=Binv, G=G):
                g = G.T @ r
                y = dg2 * g
                return Binv @ r + G @ y
            _, its = cocr(A, b, M_hyb2)
            print("   jacobi on grad(P2) only:", its, flush=True)
        dg = 1.0 / Agf.diagonal()
        # eigenvalue bound of the Jacobi-scaled nodal operator (power iteration)
        v = np.random.default_rng(1).normal(size=fr.size) + 0j
        for _ in range(30):
            v = dg * (Agf @ v)
            lmax = np.linalg.norm(v)
            v /= lmax
        for deg, ratio in ((10, 30.0), (20, 100.0), (40, 400.0)):
            ch = chebyshev(Agf, dg, 1.1 * lmax, 1.1 * lmax / ratio, deg)

            def M_ch(r, ch=ch, G=G, fr=fr, Binv=Binv):
                y = np.zeros(G.shape[1], dtype=complex)
                y[fr] = ch((G.T @ r)[fr])
                return Binv @ r + G @ y
            _, its = cocr(A, b, M_ch)
            print("   chebyshev deg %d (kappa %g): %d" % (deg, ratio, its), flush=True)


if __name__ == "__main__":
    main()

"""CPU oracle for the PETGEM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy restatement of the reference algorithm for the path
    element matrices -> global complex CSR assembly -> Dirichlet -> SpMV / Krylov
Every function cites the reference ``file:line`` (into /root/reference) it follows.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
(``petgem_b200``) never does and has no CPU fallback.

Pinning (see tests/test_oracle_golden.py, fixtures under tests/golden/ made by
``oracle/make_golden.py`` from the UNMODIFIED reference imported in the build
container):
  * shape functions / curls, element matrices Me/Ke (p=1..6, all orientation
    codes, both detJ signs, VTI sigma), orientation codes, DOF numbering, mesh
    topology and boundary dofs are pinned against the reference's own outputs;
    the reference's only hot-path known answers (tests/test_mesh.py:31-36) are
    asserted as well.
  * The PETSc half (MatSetValues/ADD_VALUES, MatZeroRowsColumns, KSP) lives in
    petsc4py, a third-party dependency that is absent here and UNPINNED in the
    reference (requirements.txt:1, "PETSc 3.7 ... 3.17 tested" DESCRIPTION.rst:33).
    Its documented semantics are restated with scipy.sparse; the reference ships
    no expected output for it (tests/test_petsc.py asserts nothing):
    PARITY UNPINNED for those three steps.

Quadrature: the reference looks up tabulated rules "of order 2p"
(hvfem.py:251, :1055-1610).  They are exact for degree 2p (checked to 1e-13 in
make_golden.py) and all integrands have degree <= 2p, so this oracle integrates
with a Gauss-Jacobi conical product rule from scipy.special instead of copying
1250 lines of tables; agreement with the reference's Me/Ke is what the golden
test measures (<= 1e-12 of max|.|).
"""
import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

# ---------------------------------------------------------------------------
# mesh topology (mesh.py, vectors.py)
# ---------------------------------------------------------------------------
EDGE_NODES = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]  # mesh.py:27-32
FACE_NODES = [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)]  # mesh.py:70-73
FACE_EDGES = [(0, 1, 2), (0, 4, 3), (1, 5, 4), (2, 5, 3)]  # mesh.py:116-123


def _unique_sorted_rows(rows):
    """vectors.py:15-64: ids = lexicographic rank of the sorted node tuple."""
    srt = np.sort(rows, axis=1)
    uniq, inverse = np.unique(srt, axis=0, return_inverse=True)
    return uniq, inverse.reshape(-1)


def compute_edges(elemsN):
    """mesh.py:18-53 -> (elemsE [T,6], edgesNodes [E,2] sorted pairs)."""
    T = elemsN.shape[0]
    pairs = np.stack([elemsN[:, list(e)] for e in EDGE_NODES], axis=1).reshape(T * 6, 2)
    edgesNodes, inv = _unique_sorted_rows(pairs)
    return inv.reshape(T, 6).astype(np.int64), edgesNodes.astype(np.int64)


def compute_faces(elemsN):
    """mesh.py:56-89 -> (elemsF [T,4], facesN [F,3]).

    facesN rows keep the node order of the first occurrence (mesh.py:82 with
    vectors.py:28: ``out = matrix[J,:]``), as the reference does.
    """
    T = elemsN.shape[0]
    tri = np.stack([elemsN[:, list(f)] for f in FACE_NODES], axis=1).reshape(T * 4, 3)
    srt = np.sort(tri, axis=1)
    _, first, inv = np.unique(srt, axis=0, return_index=True, return_inverse=True)
    return inv.reshape(T, 4).astype(np.int64), tri[first].astype(np.int64)


def compute_faces_edges(elemsF, elemsE, nFaces):
    """mesh.py:92-125: edges of face f in the local-face order of the lowest-index
    element containing f (vectors.py:121-127 stores elements in ascending order)."""
    T = elemsF.shape[0]
    facesE = np.zeros((nFaces, 3), dtype=np.int64)
    seen = np.zeros(nFaces, dtype=bool)
    for t in range(T):
        for k in range(4):
            f = elemsF[t, k]
            if not seen[f]:
                seen[f] = True
                facesE[f] = elemsE[t, list(FACE_EDGES[k])]
    return facesE


def compute_boundary_faces(elemsF, facesN):
    """mesh.py:149-221: faces referenced by exactly one element."""
    nFaces = elemsF.max() + 1
    count = np.bincount(elemsF.reshape(-1), minlength=nFaces)
    bFaces = np.nonzero(count == 1)[0]
    return facesN[bFaces].T.copy(), bFaces


def compute_boundary_edges(edgesNodes, bfacesN):
    """mesh.py:224-277: edges of the boundary faces, ascending edge id."""
    pairs = np.concatenate([bfacesN[[0, 1]].T, bfacesN[[1, 2]].T, bfacesN[[2, 0]].T], axis=0)
    pairs = np.unique(np.sort(pairs, axis=1), axis=0)
    nn = int(edgesNodes.max()) + 1
    key_all = edgesNodes[:, 0] * nn + edgesNodes[:, 1]
    key_b = pairs[:, 0] * nn + pairs[:, 1]
    return np.nonzero(np.isin(key_all, key_b))[0]


def compute_boundaries(dof_edges, dof_faces, bEdges, bFaces):
    """mesh.py:280-321: boundary dofs = dofs of boundary edges, then of boundary faces."""
    parts = [dof_edges[bEdges].reshape(-1)]
    if dof_faces.size:
        parts.append(dof_faces[bFaces].reshape(-1))
    return np.concatenate(parts).astype(np.int64)


# ---------------------------------------------------------------------------
# DOF numbering (hvfem.py:15-98)
# ---------------------------------------------------------------------------
def compute_connectivity_dofs(elemsE, elemsF, p):
    T = elemsE.shape[0]
    nE, nF = elemsE.max() + 1, elemsF.max() + 1
    ne, nf, nv = p, p * (p - 1), p * (p - 1) * (p - 2) // 2
    dof_edges = np.arange(nE * ne, dtype=np.int64).reshape(nE, ne)
    dof_faces = nE * ne + np.arange(nF * nf, dtype=np.int64).reshape(nF, nf)
    dof_volume = nE * ne + nF * nf + np.arange(T * nv, dtype=np.int64).reshape(T, nv)
    dofs = np.concatenate(
        [dof_edges[elemsE].reshape(T, 6 * ne), dof_faces[elemsF].reshape(T, 4 * nf), dof_volume], axis=1
    )
    return dofs, dof_edges, dof_faces, dof_volume, int(dofs.max()) + 1


# ---------------------------------------------------------------------------
# geometry and orientation (hvfem.py:101-220)
# ---------------------------------------------------------------------------
def compute_jacobian(coordEle):
    """hvfem.py:101-119: rows of J are edge vectors x_i - x_0."""
    J = coordEle[1:4] - coordEle[0]
    return J, np.linalg.inv(J)


def compute_element_orientation(edgesEle, nodesEle, edgesNodesEle, globalEdgesInFace):
    """hvfem.py:122-220 for one element."""
    eo = np.zeros(6, dtype=np.int64)
    for i, (a, b) in enumerate(EDGE_NODES):
        if nodesEle[a] == edgesNodesEle[i, 1] and nodesEle[b] == edgesNodesEle[i, 0]:
            eo[i] = 1
    code = {12: 0, 31: 1, 23: 2, 32: 3, 13: 4, 21: 5}
    fo = np.zeros(4, dtype=np.int64)
    for i, le in enumerate(FACE_EDGES):
        k1 = k2 = 0
        for k in range(3):
            if edgesEle[le[0]] == globalEdgesInFace[i, k]:
                k1 = k + 1
            if edgesEle[le[1]] == globalEdgesInFace[i, k]:
                k2 = k + 1
        fo[i] = code.get(10 * k1 + k2, 0)
    return eo, fo


# ---------------------------------------------------------------------------
# shape functions (hvfem.py:319-1052), one point, one orientation
# ---------------------------------------------------------------------------
def poly_legendre(x, t, n):
    """hvfem.py:629-664 (P_0..P_{n-1})."""
    P = np.zeros(max(n, 2))
    P[0] = 1.0
    P[1] = 2.0 * x - t
    for i in range(1, n - 1):
        P[i + 1] = ((2 * i + 1) * P[1] * P[i] - i * t * t * P[i - 1]) / (i + 1)
    return P


def poly_jacobi(x, t, n, minalpha):
    """hvfem.py:733-788: rows alpha = minalpha + 2 i."""
    P = np.zeros((n + 1, n + 1))
    P[:, 0] = 1.0
    y = 2 * x - t
    if n >= 1:
        for i in range(n):
            P[i, 1] = y + (minalpha + 2 * i) * x
    for i in range(n - 1):
        al = minalpha + 2 * i
        for j in range(2, n - i + 1):
            ai = 2 * j * (j + al) * (2 * j + al - 2)
            bi = 2 * j + al - 1
            ci = (2 * j + al) * (2 * j + al - 2)
            di = 2 * (j + al - 1) * (j - 1) * (2 * j + al)
            P[i, j] = (bi * (ci * y + al * al * t) * P[i, j - 1] - di * t * t * P[i, j - 2]) / ai
    return P


def poly_ijacobi(x, t, n, minalpha):
    """hvfem.py:667-730 -> (L, P, R) each [n, n]; column j-1 <-> order j."""
    pt = poly_jacobi(x, t, n, minalpha)
    L = np.zeros((n, n))
    R = np.zeros((n, n))
    L[:, 0] = x
    for i in range(n - 1):
        al = minalpha + 2 * i
        for j in range(2, n - i + 1):
            tia = 2 * j + al
            L[i, j - 1] = (
                (j + al) / ((tia - 1) * tia) * pt[i, j]
                + al / ((tia - 2) * tia) * t * pt[i, j - 1]
                - (j - 1) / ((tia - 2) * (tia - 1)) * t * t * pt[i, j - 2]
            )
            R[i, j - 1] = -(j - 1) * (pt[i, j - 1] + t * pt[i, j - 2]) / (tia - 2)
    return L, pt[:n, :n], R


def hom_ijacobi(S, DS, n, minalpha, idec):
    """hvfem.py:583-626.  DS columns are gradients of S[0], S[1]."""
    if idec:
        L, P, _ = poly_ijacobi(S[1], 1.0, n, minalpha)
        R = np.zeros_like(P)
    else:
        L, P, R = poly_ijacobi(S[1], S[0] + S[1], n, minalpha)
    D = np.zeros((3, n, n))
    for i in range(n):
        for j in range(n - i):
            D[:, i, j] = P[i, j] * DS[:, 1] + R[i, j] * (DS[:, 0] + DS[:, 1])
    return L, D


def anc_ee(S, DS, n):
    """hvfem.py:467-513 (general branch)."""
    P = poly_legendre(S[1], S[0] + S[1], n)
    w = S[0] * DS[:, 1] - S[1] * DS[:, 0]
    cw = np.cross(DS[:, 0], DS[:, 1])
    EE = np.zeros((3, n))
    CE = np.zeros((3, n))
    for i in range(n):
        EE[:, i] = P[i] * w
        CE[:, i] = (i + 2) * P[i] * cw
    return EE, CE


def anc_etri(S, DS, n):
    """hvfem.py:516-566: triangle ancillary functions of order n (Idec False)."""
    ET = np.zeros((3, max(n - 1, 0), max(n - 1, 0)))
    CT = np.zeros_like(ET)
    if n < 2:
        return ET, CT
    EE, CE = anc_ee(S[:2], DS[:, :2], n - 1)
    sL = np.array([S[0] + S[1], S[2]])
    DsL = np.stack([DS[:, 0] + DS[:, 1], DS[:, 2]], axis=1)
    L, DL = hom_ijacobi(sL, DsL, n - 1, 1, False)
    for ij in range(1, n):
        for i in range(ij):
            j = ij - i
            ET[:, i, j - 1] = EE[:, i] * L[i, j - 1]
            CT[:, i, j - 1] = L[i, j - 1] * CE[:, i] + np.cross(DL[:, i, j - 1], EE[:, i])
    return ET, CT


ORIENT_TRI = [(0, 1, 2), (1, 2, 0), (2, 0, 1), (0, 2, 1), (1, 0, 2), (2, 1, 0)]  # hvfem.py:843-867


def shape3d_etet(X, p, eo, fo):
    """hvfem.py:319-464: (ShapE, CurlE) [3, n] at master point X."""
    n = p * (p + 2) * (p + 3) // 2
    lam = np.array([1.0 - X[0] - X[1] - X[2], X[0], X[1], X[2]])  # hvfem.py:1039-1042
    dlam = np.zeros((3, 4))
    dlam[:, 0] = -1.0
    dlam[0, 1] = dlam[1, 2] = dlam[2, 3] = 1.0
    N = np.zeros((3, n))
    C = np.zeros((3, n))
    m = 0
    for e, (a, b) in enumerate(EDGE_NODES):  # hvfem.py:370-385 + OrientE :791-822
        pair = (b, a) if eo[e] == 1 else (a, b)
        EE, CE = anc_ee(lam[list(pair)], dlam[:, list(pair)], p)
        N[:, m:m + p], C[:, m:m + p] = EE, CE
        m += p
    if p >= 2:
        for f, verts in enumerate(FACE_NODES):  # hvfem.py:391-413 + OrientTri :825-878
            tri = [verts[k] for k in ORIENT_TRI[fo[f]]]
            base = m
            for fam in range(2):
                abc = tri[fam:] + tri[:fam]
                ET, CT = anc_etri(lam[abc], dlam[:, abc], p)
                slot = base + fam
                for k in range(1, p):
                    for r in range(k):
                        N[:, slot], C[:, slot] = ET[:, r, k - r - 1], CT[:, r, k - r - 1]
                        slot += 2
            m += p * (p - 1)
    if p >= 3:  # hvfem.py:417-453
        base = m
        for fam in range(3):
            abcd = [(v + fam) % 4 for v in range(4)]
            abc, d = abcd[:3], abcd[3]
            ET, CT = anc_etri(lam[abc], dlam[:, abc], p - 1)
            S2 = np.array([1.0 - lam[d], lam[d]])
            DS2 = np.stack([-dlam[:, d], dlam[:, d]], axis=1)
            L, DL = hom_ijacobi(S2, DS2, p - 2, 2, True)
            slot = base + fam
            for j in range(2, p):
                for k in range(1, j):
                    for r in range(k):
                        pp, q = k - r, j - k
                        N[:, slot] = ET[:, r, pp - 1] * L[k - 1, q - 1]
                        C[:, slot] = L[k - 1, q - 1] * CT[:, r, pp - 1] + np.cross(DL[:, k - 1, q - 1], ET[:, r, pp - 1])
                        slot += 3
        m += p * (p - 1) * (p - 2) // 2
    assert m == n
    return N, C


# ---------------------------------------------------------------------------
# element matrices (hvfem.py:223-316)
# ---------------------------------------------------------------------------
_QUAD = {}


def tet_rule(degree):
    """Conical Gauss-Jacobi rule exact for `degree` on the master tetrahedron
    (stands in for compute3DGaussPoints, hvfem.py:1055-1610; see module header)."""
    if degree not in _QUAD:
        n = degree // 2 + 1
        rules = []
        for al in (2, 1, 0):
            x, w = roots_jacobi(n, al, 0)
            rules.append(((x + 1) / 2, w / 2 ** (al + 1)))
        pts, wts = [], []
        for u, wu in zip(*rules[0]):
            for v, wv in zip(*rules[1]):
                for s, ws in zip(*rules[2]):
                    pts.append((u, v * (1 - u), s * (1 - u) * (1 - v)))
                    wts.append(wu * wv * ws)
        _QUAD[degree] = (np.array(pts), np.array(wts))
    return _QUAD[degree]


_CLASS_CACHE = {}


def _class_basis(p, eo, fo):
    key = (p, tuple(int(v) for v in eo), tuple(int(v) for v in fo))
    if key not in _CLASS_CACHE:
        pts, wts = tet_rule(2 * p)
        NN = np.stack([np.stack(shape3d_etet(X, p, eo, fo)) for X in pts])  # [g, 2, 3, n]
        _CLASS_CACHE[key] = (NN[:, 0], NN[:, 1], wts)
    return _CLASS_CACHE[key]


def compute_elemental_matrices(eo, fo, J, Jinv, p, sigmaEle):
    """hvfem.py:223-316: same quadrature loop, written as two contractions."""
    Nref, Cref, W = _class_basis(p, eo, fo)
    det = np.linalg.det(J)  # signed, hvfem.py:265
    e_r = np.diag([sigmaEle[0], sigmaEle[0], sigmaEle[1]])
    Nreal = np.einsum("ab,gbj->gaj", Jinv, Nref)  # hvfem.py:292
    Creal = np.einsum("gaj,ab->gjb", Cref, J) / det  # hvfem.py:304
    Me = np.einsum("g,gaj,ab,gbk->jk", W, Nreal, e_r, Nreal) * det  # :297, :313
    Ke = np.einsum("g,gja,gka->jk", W, Creal, Creal) * det  # :310, :314
    return Me, Ke


def element_system(coordEle, nodesEle, edgesEle, edgesNodesEle, edgesFace, sigmaEle, p, omega, mu):
    """solver.py:214-224: Ae = K - i omega mu M for one element."""
    J, Jinv = compute_jacobian(coordEle)
    eo, fo = compute_element_orientation(edgesEle, nodesEle, edgesNodesEle, edgesFace)
    Me, Ke = compute_elemental_matrices(eo, fo, J, Jinv, p, sigmaEle)
    return Ke - 1j * omega * mu * Me


# ---------------------------------------------------------------------------
# global assembly, Dirichlet, solve (solver.py:188-235, :552-590; PETSc semantics)
# ---------------------------------------------------------------------------
def assemble_global(Ae_all, dofs, N):
    """MatSetValues(ADD_VALUES) for every element then MatAssembly:
    duplicates summed in element order, explicit zeros kept, columns sorted."""
    T, n = dofs.shape
    rows = np.repeat(dofs, n, axis=1).reshape(-1)
    cols = np.tile(dofs, (1, n)).reshape(-1)
    key = rows.astype(np.int64) * N + cols
    order = np.argsort(key, kind="stable")  # stable: ascending element order per entry
    ks = key[order]
    start = np.concatenate([[True], ks[1:] != ks[:-1]])
    seg = np.cumsum(start) - 1
    vals = np.zeros(int(seg[-1]) + 1, dtype=np.complex128)
    np.add.at(vals, seg, Ae_all.reshape(-1)[order])
    ukey = ks[start]
    r, c = ukey // N, ukey % N
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, c.astype(np.int32), vals


def csr_pattern(dofs, N):
    """Pattern only (rowptr int64, colidx int32) of the union of element cliques."""
    T, n = dofs.shape
    key = (np.repeat(dofs, n, axis=1).astype(np.int64) * N + np.tile(dofs, (1, n))).reshape(-1)
    ukey = np.unique(key)
    r = ukey // N
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr), (ukey % N).astype(np.int32)


def zero_rows_columns(rowptr, colidx, vals, bd, diag=1.0):
    """MatZeroRowsColumns (solver.py:562): zero rows and columns of bd, put diag on the
    diagonal, keep the pattern."""
    N = rowptr.size - 1
    mask = np.zeros(N, dtype=bool)
    mask[bd] = True
    rows = np.repeat(np.arange(N), np.diff(rowptr))
    out = vals.copy()
    out[mask[rows] | mask[colidx]] = 0.0
    out[(rows == colidx) & mask[rows]] = diag
    return out


def to_scipy(rowptr, colidx, vals):
    N = rowptr.size - 1
    return sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))


def spmv(rowptr, colidx, vals, x):
    """MatMult: y = A x."""
    return to_scipy(rowptr, colidx, vals) @ x


def gmres(matvec, b, rtol=1e-8, restart=30, maxit=10000, pc=None, x0=None):
    """PETSc KSPGMRES defaults: left preconditioning, restart 30,
    classical Gram-Schmidt without refinement, zero initial guess, convergence
    on the preconditioned residual norm relative to ||M^-1 b||."""
    apply_pc = (lambda v: v) if pc is None else pc
    n = b.size
    x = np.zeros(n, dtype=np.complex128) if x0 is None else x0.copy()
    bnorm = np.linalg.norm(apply_pc(b))
    if bnorm == 0.0:
        return x, 0, [0.0]
    its = 0
    hist = []
    while True:
        r = apply_pc(b - matvec(x))
        beta = np.linalg.norm(r)
        if not hist:
            hist.append(beta)
        if beta <= rtol * bnorm or its >= maxit:
            return x, its, hist
        V = np.zeros((restart + 1, n), dtype=np.complex128)
        H = np.zeros((restart + 1, restart), dtype=np.complex128)
        V[0] = r / beta
        g = np.zeros(restart + 1, dtype=np.complex128)
        g[0] = beta
        cs = np.zeros(restart, dtype=np.complex128)
        sn = np.zeros(restart, dtype=np.complex128)
        k = 0
        while k < restart and its < maxit:
            w = apply_pc(matvec(V[k]))
            h = V[: k + 1].conj() @ w  # VecMDot
            w = w - h @ V[: k + 1]  # VecMAXPY
            H[: k + 1, k] = h
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] != 0:
                V[k + 1] = w / H[k + 1, k]
            for i in range(k):  # apply previous rotations
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -np.conj(sn[i]) * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            a, bb = H[k, k], H[k + 1, k]
            den = np.sqrt(abs(a) ** 2 + abs(bb) ** 2)
            cs[k] = abs(a) / den if a != 0 else 0.0
            sn[k] = (a / abs(a)) * np.conj(bb) / den if a != 0 else 1.0
            H[k, k] = cs[k] * a + sn[k] * bb
            H[k + 1, k] = 0.0
            g[k + 1] = -np.conj(sn[k]) * g[k]
            g[k] = cs[k] * g[k]
            k += 1
            its += 1
            hist.append(abs(g[k]))
            if abs(g[k]) <= rtol * bnorm:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        x = x + y @ V[:k]


def bicgstab(matvec, b, rtol=1e-8, maxit=10000, pc=None):
    """KSPBCGS with left preconditioning (unconjugated shadow residual as PETSc)."""
    apply_pc = (lambda v: v) if pc is None else pc
    x = np.zeros_like(b, dtype=np.complex128)
    r = apply_pc(b - matvec(x))
    bnorm = np.linalg.norm(apply_pc(b))
    rhat = r.copy()
    rho = alpha = omega = 1.0
    v = np.zeros_like(r)
    p_ = np.zeros_like(r)
    hist = [np.linalg.norm(r)]
    for it in range(1, maxit + 1):
        rho_new = np.vdot(rhat, r)
        beta = (rho_new / rho) * (alpha / omega)
        p_ = r + beta * (p_ - omega * v)
        v = apply_pc(matvec(p_))
        alpha = rho_new / np.vdot(rhat, v)
        s = r - alpha * v
        t = apply_pc(matvec(s))
        omega = np.vdot(t, s) / np.vdot(t, t)
        x = x + alpha * p_ + omega * s
        r = s - omega * t
        rho = rho_new
        hist.append(np.linalg.norm(r))
        if hist[-1] <= rtol * bnorm:
            return x, it, hist
    return x, maxit, hist


def cocg(matvec, b, rtol=1e-8, maxit=10000, dinv=None):
    """Conjugate-orthogonal CG for a complex SYMMETRIC matrix (van der Vorst & Melissen 1990; what
    `-ksp_type cg -ksp_cg_type symmetric` runs in PETSc): CG with the unconjugated bilinear form x^T y,
    symmetric Jacobi through z = dinv * r, convergence on ||z|| / ||dinv * b||."""
    d = np.ones_like(b) if dinv is None else dinv
    x = np.zeros_like(b, dtype=np.complex128)
    r = b.astype(np.complex128).copy()
    z = d * r
    p_ = z.copy()
    rho = r @ z
    bnorm = np.linalg.norm(z)
    hist = [bnorm]
    for it in range(1, maxit + 1):
        q = matvec(p_)
        alpha = rho / (p_ @ q)
        x = x + alpha * p_
        r = r - alpha * q
        z = d * r
        rho_new = r @ z
        p_ = z + (rho_new / rho) * p_
        rho = rho_new
        hist.append(np.linalg.norm(z))
        if hist[-1] <= rtol * bnorm:
            return x, it, hist
    return x, maxit, hist


def cocr(matvec, b, rtol=1e-8, maxit=10000, dinv=None):
    """Conjugate-orthogonal conjugate residuals for a complex symmetric matrix (Sogabe & Zhang 2007,
    preconditioned form): recurrences on rt = dinv * r with the unconjugated bilinear form; one matvec per
    iteration (A p is updated by recurrence); convergence on ||rt|| / ||dinv * b||."""
    d = np.ones_like(b) if dinv is None else dinv
    x = np.zeros_like(b, dtype=np.complex128)
    rt = d * b.astype(np.complex128)
    p_ = rt.copy()
    art = matvec(rt)
    ap = art.copy()
    rho = rt @ art
    bnorm = np.linalg.norm(rt)
    hist = [bnorm]
    for it in range(1, maxit + 1):
        map_ = d * ap
        alpha = rho / (ap @ map_)
        x = x + alpha * p_
        rt = rt - alpha * map_
        art = matvec(rt)
        rho_new = rt @ art
        beta = rho_new / rho
        rho = rho_new
        p_ = rt + beta * p_
        ap = art + beta * ap
        hist.append(np.linalg.norm(rt))
        if hist[-1] <= rtol * bnorm:
            return x, it, hist
    return x, maxit, hist


# ---------------------------------------------------------------------------
# CSEM right-hand side and receiver interpolation
# ---------------------------------------------------------------------------
def xyz_to_xietazeta(coordEle, point):
    """hvfem.py:2347-2490 (affine inverse map; same result as the expanded formulas)."""
    J = coordEle[1:4] - coordEle[0]
    return np.linalg.solve(J.T, point - coordEle[0])


def compute_basis_functions(eo, fo, J, Jinv, p, X):
    """hvfem.py:2493-2550 at one master point."""
    N, C = shape3d_etet(X, p, eo, fo)
    return Jinv @ N, J.T @ C / np.linalg.det(J.T)


def source_rotation(azimuth, dip):
    """hvfem.py:2303-2344."""
    a, b = np.deg2rad(azimuth), np.deg2rad(dip)
    M1 = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    M2 = np.array([[np.cos(b), 0, -np.sin(b)], [0, 1.0, 0], [np.sin(b), 0, np.cos(b)]])
    return M1 @ M2 @ np.array([1.0, 0, 0])


def csem_rhs(N, coordEle, nodesEle, edgesEle, edgesNodesEle, edgesFace, dofsEle, p, position,
             azimuth, dip, current, length, omega, mu):
    """solver.py:247-316."""
    rot = source_rotation(azimuth, dip)
    field = current * length * rot
    J, Jinv = compute_jacobian(coordEle)
    eo, fo = compute_element_orientation(edgesEle, nodesEle, edgesNodesEle, edgesFace)
    X = xyz_to_xietazeta(coordEle, position)
    basis, _ = compute_basis_functions(eo, fo, J, Jinv, p, X)
    b = np.zeros(N, dtype=np.complex128)
    np.add.at(b, dofsEle, 1j * omega * mu * (field @ basis))
    return b


def locate_points(nodes, elemsN, points, tol=1e-12):
    """Brute-force containing element (stands in for Delaunay.find_simplex with the
    mesh connectivity, postprocessing.py:532-539): lowest-index element whose
    barycentric coordinates are all >= -tol."""
    out = np.full(points.shape[0], -1, dtype=np.int64)
    X0 = nodes[elemsN[:, 0]]
    Jt = np.stack([nodes[elemsN[:, k]] - X0 for k in (1, 2, 3)], axis=2)  # columns
    Jti = np.linalg.inv(Jt)
    for i, pt in enumerate(points):
        loc = np.einsum("tab,tb->ta", Jti, pt[None, :] - X0)
        l0 = 1.0 - loc.sum(axis=1)
        ok = (loc >= -tol).all(axis=1) & (l0 >= -tol)
        idx = np.nonzero(ok)[0]
        if idx.size:
            out[i] = idx[0]
    return out


def field_interpolator(x, nodes, elemsN, elemsE, edgesNodes, elemsF, facesE, dofs, points, p, omega, mu,
                       elements=None):
    """postprocessing.py:479-616: E and H at receiver points -> [npts, 6] complex."""
    if elements is None:
        elements = locate_points(nodes, elemsN, points)
    out = np.zeros((points.shape[0], 6), dtype=np.complex128)
    for i, (pt, t) in enumerate(zip(points, elements)):
        coordEle = nodes[elemsN[t]]
        J, Jinv = compute_jacobian(coordEle)
        eo, fo = compute_element_orientation(elemsE[t], elemsN[t], edgesNodes[elemsE[t]], facesE[elemsF[t]])
        X = xyz_to_xietazeta(coordEle, pt)
        basis, curl = compute_basis_functions(eo, fo, J, Jinv, p, X)
        xe = x[dofs[t]]
        out[i, :3] = basis @ xe
        out[i, 3:] = (curl @ xe) / (1j * omega * mu)
    return out

"""Reference-element H(curl) tables for the B200 element kernels.

The reference re-evaluates the whole hierarchical basis for every element at
every Gauss point (``petgem/hvfem.py:270-281`` calling ``shape3DETet``
``hvfem.py:319-464``), although the basis on the master tetrahedron depends only
on the polynomial order and on the orientation of each edge/face, never on the
geometry.  Here everything that is geometry-independent is evaluated ONCE, per
order p and per *entity variant*, vectorised over evaluation points, and folded
into small dense tensors that the CUDA kernels contract with per-element 3x3
geometric factors:

    Me[j,k] = sum_c gM[c] * SM[c, J(j), J(k)] * s(j) s(k)
    Ke[j,k] = sum_c gK[c] * SK[c, J(j), J(k)] * s(j) s(k)

c runs over the 6 independent entries (00,11,22,01,02,12) of the symmetric
geometric factors gM = detJ * J^-T diag(sh,sh,sv) J^-1 and gK = J J^T / detJ
(same arithmetic as ``hvfem.py:292-314``), J(j) is the *expanded* index of local
dof j (see :func:`expanded_layout`) and s(j) = +-1 is the edge-flip sign.

Expanded (orientation-resolved) function set, in this order:
  * 6 edges x p functions, evaluated for edge orientation 0; orientation 1
    (``OrientE``, ``hvfem.py:791-822``) only flips the sign of the functions of
    even polynomial degree, i.e. s = (-1)^(i+1) for the i-th edge function;
  * 4 faces x 6 orientations (``OrientTri``, ``hvfem.py:825-878``) x p(p-1)
    functions (two interleaved families, ``hvfem.py:402-413``);
  * p(p-1)(p-2)/2 interior (bubble) functions, three interleaved families
    (``hvfem.py:429-453``), orientation-free.

The integrals are evaluated with our own conical Gauss-Jacobi rule of degree
2p+1 (positive weights).  The reference uses tabulated rules "of order 2p"
(``hvfem.py:251``, ``:1055-1610``) which integrate degree-2p polynomials
exactly, and every integrand here has degree <= 2p, so both give the exact
integral up to rounding; no quadrature table is copied.
"""
from __future__ import annotations

import functools

import numpy as np

# local topology of the master tetrahedron (hvfem.py:150-161, :177-184, :976-1011)
LOCAL_EDGES = np.array([[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]], dtype=np.int64)
LOCAL_FACES = np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3], [0, 2, 3]], dtype=np.int64)
FACE_PERMS = np.array(
    [[0, 1, 2], [1, 2, 0], [2, 0, 1], [0, 2, 1], [1, 0, 2], [2, 1, 0]], dtype=np.int64
)
# (a,b) pairs of the packed symmetric 3x3 index c
SYM_PAIRS = ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))

# constant gradients of the affine coordinates (rows = lambda_0..3), hvfem.py:1039-1050
GRAD_LAMBDA = np.array(
    [[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
)


def ndof_element(p: int) -> int:
    return p * (p + 2) * (p + 3) // 2


def ndof_edge(p: int) -> int:
    return p


def ndof_face(p: int) -> int:
    return p * (p - 1)


def ndof_volume(p: int) -> int:
    return p * (p - 1) * (p - 2) // 2


def expanded_layout(p: int) -> dict:
    """Offsets of the expanded function set for order p."""
    ne, nf, nv = ndof_edge(p), ndof_face(p), ndof_volume(p)
    return {
        "p": p,
        "n": ndof_element(p),
        "ne": ne,
        "nf": nf,
        "nv": nv,
        "edge_off": 0,
        "face_off": 6 * ne,
        "vol_off": 6 * ne + 24 * nf,
        "nexp": 6 * ne + 24 * nf + nv,
    }


# ----------------------------------------------------------------------------
# scaled orthogonal polynomials, vectorised over points
# ----------------------------------------------------------------------------
def _scaled_legendre(x, t, nmax):
    """P_0..P_nmax of the shifted scaled Legendre family at (x; t)."""
    out = np.empty((nmax + 1,) + x.shape)
    out[0] = 1.0
    if nmax >= 1:
        y = 2.0 * x - t
        out[1] = y
        tt = t * t
        for i in range(1, nmax):
            out[i + 1] = ((2 * i + 1) * y * out[i] - i * tt * out[i - 1]) / (i + 1)
    return out


def _scaled_jacobi(x, t, nmax, alpha):
    """P^alpha_0..P^alpha_nmax of the shifted scaled Jacobi family at (x; t)."""
    out = np.empty((nmax + 1,) + x.shape)
    out[0] = 1.0
    if nmax >= 1:
        y = 2.0 * x - t
        out[1] = y + alpha * x
        tt = t * t
        aa = float(alpha * alpha)
        for j in range(2, nmax + 1):
            a = 2.0 * j * (j + alpha) * (2 * j + alpha - 2)
            b = 2.0 * j + alpha - 1
            c = (2.0 * j + alpha) * (2 * j + alpha - 2)
            d = 2.0 * (j + alpha - 1) * (j - 1) * (2 * j + alpha)
            out[j] = (b * (c * y + aa * t) * out[j - 1] - d * tt * out[j - 2]) / a
    return out


def _integrated_jacobi(x, t, nmax, alpha):
    """L^alpha_1..L^alpha_nmax with dL/dx and dL/dt, index 0 <-> order 1."""
    P = _scaled_jacobi(x, t, nmax, alpha)
    L = np.empty((nmax,) + x.shape)
    dLdx = np.empty_like(L)
    dLdt = np.zeros_like(L)
    L[0] = x
    dLdx[0] = P[0]
    tt = t * t
    for j in range(2, nmax + 1):
        t0 = 2.0 * j + alpha
        a = (j + alpha) / ((t0 - 1) * t0)
        b = alpha / ((t0 - 2) * t0)
        c = (j - 1) / ((t0 - 2) * (t0 - 1))
        L[j - 1] = a * P[j] + b * t * P[j - 1] - c * tt * P[j - 2]
        dLdx[j - 1] = P[j - 1]
        dLdt[j - 1] = -(j - 1) * (P[j - 1] + t * P[j - 2]) / (t0 - 2)
    return L, dLdx, dLdt


# ----------------------------------------------------------------------------
# ancillary function families (values [nfun, npts, 3], curls likewise)
# ----------------------------------------------------------------------------
def _edge_family(s0, s1, g0, g1, nfun):
    """E_i = P_i(s1; s0+s1) (s0 grad s1 - s1 grad s0), curl E_i = (i+2) P_i g0 x g1."""
    P = _scaled_legendre(s1, s0 + s1, max(nfun - 1, 0))
    whitney = s0[:, None] * g1[None, :] - s1[:, None] * g0[None, :]
    cw = np.cross(g0, g1)
    val = P[:nfun, :, None] * whitney[None, :, :]
    mult = np.arange(2, nfun + 2, dtype=float)
    curl = (mult[:, None] * P[:nfun])[:, :, None] * cw[None, None, :]
    return val, curl


def _triangle_family(s, g, order):
    """Triangle ancillary functions of the given order for the triple s=(s0,s1,s2).

    Returns dict (i, j) -> (value [npts,3], curl [npts,3]) for i>=0, j>=1, i+j<=order-1.
    """
    out = {}
    if order < 2:
        return out
    E, cE = _edge_family(s[0], s[1], g[0], g[1], order - 1)
    t = s[0] + s[1] + s[2]
    gsum = g[0] + g[1] + g[2]
    for i in range(order - 1):
        jmax = order - 1 - i
        L, dLdx, dLdt = _integrated_jacobi(s[2], t, jmax, 2 * i + 1)
        for j in range(1, jmax + 1):
            gradL = dLdx[j - 1][:, None] * g[2][None, :] + dLdt[j - 1][:, None] * gsum[None, :]
            val = E[i] * L[j - 1][:, None]
            curl = L[j - 1][:, None] * cE[i] + np.cross(gradL, E[i])
            out[(i, j)] = (val, curl)
    return out


def _face_functions(lam, verts, perm, p):
    """Oriented face functions in storage order (hvfem.py:402-413)."""
    nf = ndof_face(p)
    npts = lam.shape[1]
    val = np.zeros((nf, npts, 3))
    curl = np.zeros((nf, npts, 3))
    if nf == 0:
        return val, curl
    tri = [verts[perm[0]], verts[perm[1]], verts[perm[2]]]
    for fam in range(2):
        abc = [tri[(0 + fam) % 3], tri[(1 + fam) % 3], tri[(2 + fam) % 3]]
        fun = _triangle_family([lam[v] for v in abc], [GRAD_LAMBDA[v] for v in abc], p)
        slot = fam
        for k in range(1, p):
            for i in range(0, k):
                v, c = fun[(i, k - i)]
                val[slot], curl[slot] = v, c
                slot += 2
    return val, curl


def _bubble_functions(lam, p):
    """Interior functions in storage order (hvfem.py:429-453)."""
    nv = ndof_volume(p)
    npts = lam.shape[1]
    val = np.zeros((nv, npts, 3))
    curl = np.zeros((nv, npts, 3))
    if nv == 0:
        return val, curl
    for fam in range(3):
        a, b, c, d = [(v + fam) % 4 for v in range(4)]
        tri = _triangle_family([lam[a], lam[b], lam[c]], [GRAD_LAMBDA[v] for v in (a, b, c)], p - 1)
        one = np.ones_like(lam[d])
        slot = fam
        cache = {}
        for j in range(2, p):
            for k in range(1, j):
                if k not in cache:
                    cache[k] = _integrated_jacobi(lam[d], one, p - 2, 2 * k)
                L, dLdx, _ = cache[k]
                q = j - k
                for r in range(0, k):
                    tv, tc = tri[(r, k - r)]
                    gradL = dLdx[q - 1][:, None] * GRAD_LAMBDA[d][None, :]
                    val[slot] = tv * L[q - 1][:, None]
                    curl[slot] = L[q - 1][:, None] * tc + np.cross(gradL, tv)
                    slot += 3
    return val, curl


def evaluate_expanded(p: int, pts: np.ndarray):
    """All expanded functions and curls at master-element points.

    :param pts: [npts, 3] (xi, eta, zeta)
    :return: (N, C) each [nexp, npts, 3]
    """
    pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
    lay = expanded_layout(p)
    lam = np.stack([1.0 - pts[:, 0] - pts[:, 1] - pts[:, 2], pts[:, 0], pts[:, 1], pts[:, 2]])
    N = np.zeros((lay["nexp"], pts.shape[0], 3))
    C = np.zeros_like(N)
    for e, (a, b) in enumerate(LOCAL_EDGES):
        v, c = _edge_family(lam[a], lam[b], GRAD_LAMBDA[a], GRAD_LAMBDA[b], p)
        N[e * p:(e + 1) * p], C[e * p:(e + 1) * p] = v, c
    nf = lay["nf"]
    for f in range(4):
        for o in range(6):
            v, c = _face_functions(lam, LOCAL_FACES[f], FACE_PERMS[o], p)
            off = lay["face_off"] + (f * 6 + o) * nf
            N[off:off + nf], C[off:off + nf] = v, c
    v, c = _bubble_functions(lam, p)
    N[lay["vol_off"]:], C[lay["vol_off"]:] = v, c
    return N, C


def local_to_expanded(p: int, edge_or, face_or):
    """Expanded index J(j) and sign s(j) of every local dof for given orientations.

    ``edge_or`` [..., 6] in {0,1}; ``face_or`` [..., 4] in 0..5 (codes of
    ``computeElementOrientation``, hvfem.py:122-220).  Vectorised over leading dims.
    """
    lay = expanded_layout(p)
    edge_or = np.asarray(edge_or, dtype=np.int64)
    face_or = np.asarray(face_or, dtype=np.int64)
    lead = edge_or.shape[:-1]
    J = np.empty(lead + (lay["n"],), dtype=np.int64)
    S = np.ones(lead + (lay["n"],), dtype=np.float64)
    i = np.arange(p)
    flip = np.where(i % 2 == 0, -1.0, 1.0)  # (-1)^(i+1)
    for e in range(6):
        J[..., e * p:(e + 1) * p] = e * p + i
        S[..., e * p:(e + 1) * p] = np.where(edge_or[..., e, None] == 1, flip, 1.0)
    nf = lay["nf"]
    kf = np.arange(nf)
    for f in range(4):
        J[..., 6 * p + f * nf:6 * p + (f + 1) * nf] = (
            lay["face_off"] + (f * 6 + face_or[..., f, None]) * nf + kf
        )
    nv = lay["nv"]
    J[..., 6 * p + 4 * nf:] = lay["vol_off"] + np.arange(nv)
    return J, S


# ----------------------------------------------------------------------------
# quadrature (own rule; see module docstring)
# ----------------------------------------------------------------------------
def _gauss_jacobi_01(n: int, alpha: int):
    """n-point Gauss rule on [0,1] for the weight (1-u)^alpha (Golub-Welsch)."""
    a, b = float(alpha), 0.0
    k = np.arange(n, dtype=float)
    # three-term recurrence of Jacobi(alpha, 0) on [-1, 1]
    diag = np.empty(n)
    diag[0] = (b - a) / (a + b + 2.0)
    kk = k[1:]
    diag[1:] = (b * b - a * a) / ((2 * kk + a + b) * (2 * kk + a + b + 2.0))
    off = (
        2.0 / (2 * kk + a + b)
        * np.sqrt(kk * (kk + a) * (kk + b) * (kk + a + b) / ((2 * kk + a + b - 1.0) * (2 * kk + a + b + 1.0)))
    )
    T = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
    x, V = np.linalg.eigh(T)
    mu0 = 2.0 ** (a + b + 1) / (a + b + 1)  # integral of (1-x)^a on [-1,1] (b=0)
    w = mu0 * V[0] ** 2
    # map to [0,1]: x=2u-1, (1-x)^a = 2^a (1-u)^a, dx = 2 du
    return (x + 1.0) / 2.0, w / 2.0 ** (a + 1)


@functools.lru_cache(maxsize=None)
def tet_quadrature(degree: int):
    """Positive-weight conical-product rule on the master tetrahedron exact for `degree`."""
    n = degree // 2 + 1
    u, wu = _gauss_jacobi_01(n, 2)
    v, wv = _gauss_jacobi_01(n, 1)
    w, ww = _gauss_jacobi_01(n, 0)
    U, V, W = np.meshgrid(u, v, w, indexing="ij")
    wt = wu[:, None, None] * wv[None, :, None] * ww[None, None, :]
    x = U
    y = V * (1.0 - U)
    z = W * (1.0 - U) * (1.0 - V)
    pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    return pts, wt.ravel()


# ----------------------------------------------------------------------------
# contraction tables
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def element_tables(p: int):
    """Packed-symmetric reference tensors SM, SK: [6, nexp, nexp] float64.

    S[c] = T^{ab} + T^{ba} (a<b) or T^{aa}, with T^{ab}_{JK} = int N_J^a N_K^b over the
    master tetrahedron (curls for SK).
    """
    pts, wts = tet_quadrature(2 * p + 1)
    N, C = evaluate_expanded(p, pts)
    out = []
    for F in (N, C):
        Fw = F * wts[None, :, None]
        T = np.einsum("jga,kgb->abjk", Fw, F, optimize=True)
        S = np.empty((6,) + T.shape[2:])
        for c, (a, b) in enumerate(SYM_PAIRS):
            S[c] = T[a, a] if a == b else T[a, b] + T[b, a]
        out.append(np.ascontiguousarray(S))
    return out[0], out[1]


# denominators that make the reference tensors integral (exact rationals: integrals of
# integer-coefficient polynomials over the master tetrahedron); used by the p <= 2 kernel,
# which keeps SK*DK and SM*DM as 16-bit integers in shared memory (pg_assemble.cu)
INTEGER_SCALES = {1: (6, 120), 2: (120, 5040), 3: (5040, 362880), 4: (362880, 39916800), 5: (39916800, 6227020800)}


def integer_tables(p: int):
    """(SM*DM, SK*DK) rounded to integers, with the worst rounding distance (must be ~1e-12)."""
    DK, DM = INTEGER_SCALES[p]
    SM, SK = element_tables(p)
    nM, nK = np.rint(SM * DM), np.rint(SK * DK)
    err = max(np.abs(SM * DM - nM).max(), np.abs(SK * DK - nK).max())
    return nM.astype(np.int64), nK.astype(np.int64), float(err)

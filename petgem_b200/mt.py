"""Magnetotelluric right-hand side (host side of the hot path's caller).

Mirrors what ``petgem/solver.py:318-512`` does for ``mode: mt``: a 1-D finite-element solve of the
layered-earth problem along z gives the excitation on the four lateral sides and the top of the box
(``petgem/mt1d.py:21-76``), and every boundary face contributes the surface integral of the
tangential basis functions against that field (Neumann condition), integrated with a symmetric
triangle rule of degree 2p.  Vectorised numpy, evaluated on every rank (b is replicated); the
two polarizations then share the matrix and are solved in lockstep (``krylov.solve_multi``).
"""
from __future__ import annotations

import numpy as np

from . import basis
from .quadrature2d import triangle_quadrature

# local faces by local nodes (hvfem.py:2665-2683) and outward reference normals (hvfem.py:2627-2645)
FACE_NODES = np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3], [0, 2, 3]], dtype=np.int64)
REF_NORMALS = np.array([[0.0, 0.0, -1.0], [0.0, -1.0, 0.0], [1.0, 1.0, 1.0], [-1.0, 0.0, 0.0]])
REF_NORMALS[2] /= np.sqrt(3.0)
REF_VERTICES = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def _interp_first_segments(x, u, xp):
    """``mt1d.linearInterp1D`` (mt1d.py:96-127) for ascending nodes x: linear interpolation inside,
    linear EXTRAPOLATION with the first segment below x[0], and ZERO above x[-1] (the reference's
    sweep never assigns those points)."""
    x, u, xp = np.asarray(x, dtype=np.float64), np.asarray(u), np.asarray(xp, dtype=np.float64)
    out = np.zeros(xp.shape, dtype=u.dtype)
    if x.size < 2:
        return out
    seg = np.clip(np.searchsorted(x, xp, side="left"), 1, x.size - 1)  # segment [seg-1, seg], x[seg-1] < xp <= x[seg]
    h = x[seg] - x[seg - 1]
    t = (xp - x[seg - 1]) / h
    val = (1.0 - t) * u[seg - 1] + t * u[seg]
    inside = xp <= x[-1]
    out[inside] = val[inside]
    return out


def eval_MT1D(za, zb, za_D, zb_D, sigma0, x0, omega, mu, n_nodes_ref, degree=1, interpolate_at=None):
    """1-D MT problem on [za, zb] with Dirichlet values (za_D, zb_D), linear elements
    (``mt1d.eval_MT1D``, mt1d.py:21-76; only degree 1 is implemented there too):
    (K + i omega mu M(sigma)) u = 0.  sigma is sampled at the mesh nodes from (x0, sigma0) with the
    reference's interpolation rule; the tridiagonal system is solved directly."""
    from scipy.linalg import solve_banded

    if degree != 1:
        raise NotImplementedError("eval_MT1D: degree 1 only (as in the reference)")
    n_nodes = int(np.ceil(n_nodes_ref - 1)) + 1
    x = np.linspace(za, zb, n_nodes)
    x0 = np.asarray(x0, dtype=np.float64)
    x0u, pos = np.unique(x0, return_index=True)
    sigma = _interp_first_segments(x0u, np.asarray(sigma0, dtype=np.float64)[pos], x)
    # element matrices in closed form (3-point Gauss is exact for them): J = (x[e+1] - x[e]) / 2, signed
    J = 0.5 * np.diff(x)
    s0, s1 = sigma[:-1], sigma[1:]
    m00, m01, m11 = J * (s0 / 2 + s1 / 6), J * (s0 + s1) / 6, J * (s0 / 6 + s1 / 2)
    k = 0.5 / J
    c = 1j * omega * mu
    diag = np.zeros(n_nodes, dtype=np.complex128)
    diag[:-1] += k + c * m00
    diag[1:] += k + c * m11
    off = -k + c * m01  # couples node e and e+1
    f = np.zeros(n_nodes, dtype=np.complex128)
    # essential conditions (mt1d.applyEssentialBC): move the known values to the right-hand side
    f[1] -= off[0] * za_D
    f[-2] -= off[-1] * zb_D
    ab = np.zeros((3, n_nodes), dtype=np.complex128)
    ab[0, 1:] = off
    ab[1] = diag
    ab[2, :-1] = off
    ab[0, 1] = 0.0
    ab[2, 0] = 0.0
    ab[0, -1] = 0.0
    ab[2, -2] = 0.0
    ab[1, 0] = ab[1, -1] = 1.0
    f[0], f[-1] = za_D, zb_D
    u = solve_banded((1, 1), ab, f)
    if interpolate_at is None:
        return u
    pts = np.asarray(interpolate_at, dtype=np.float64)
    order = np.argsort(x)
    return _interp_first_segments(x[order], u[order], pts.reshape(-1)).reshape(pts.shape)


def neumann_excitation(face_flag, polarization, ud):
    """``hvfem.getNeumannBCface`` (hvfem.py:2556-2624), vectorised over faces: (ex, ey, ez) from the
    1-D field ud [nb, ng] by the side of the box the face lies on (0 bottom, 1 left, 2 front,
    3 right, 4 back, 5 top) and the polarization (1 = x, 2 = y)."""
    face_flag = np.asarray(face_flag)
    z = np.zeros_like(ud)
    ex, ey, ez = z.copy(), z.copy(), z.copy()
    f = face_flag[:, None]
    if polarization == 1:
        ex = np.where(f == 0, ud, np.where(f == 5, -ud, ex))
        ez = np.where(f == 1, -ud, np.where(f == 3, ud, ez))
    elif polarization == 2:
        ey = np.where(f == 0, ud, np.where(f == 5, -ud, ey))
        ez = np.where(f == 2, -ud, np.where(f == 4, ud, ez))
    else:
        raise ValueError("polarization mode not supported (1 = x, 2 = y)")
    return ex, ey, ez


def mt_rhs(boundary_rows, z_max, z_min, Nord, omega, mu, polarizations, total_dofs, n_nodes_1d=int(1e6)):
    """Right-hand sides b[i] of the MT problem, one per polarization (solver.py:318-512).

    boundary_rows: [nb, 53 + n] rows of ``boundaryElements.dat`` (preprocessing.py:326-367): nodes 0:4,
    coordinates 4:16, faces 16:20, edges of the faces 20:32, edges 32:38, nodes of the edges 38:50,
    plane flag 50, global face id 51, sigma 52, dofs 53:.
    """
    from . import hvfem

    rows = np.asarray(boundary_rows, dtype=np.float64)
    nb = rows.shape[0]
    p = int(Nord)
    n = basis.ndof_element(p)
    pts2, wts = triangle_quadrature(2 * p)
    ng = wts.size
    nodes_ele = rows[:, 0:4].astype(np.int64)
    coord = rows[:, 4:16].reshape(nb, 4, 3)
    faces_ele = rows[:, 16:20].astype(np.int64)
    edges_face = rows[:, 20:32].astype(np.int64).reshape(nb, 4, 3)
    edges_ele = rows[:, 32:38].astype(np.int64)
    edges_nodes = rows[:, 38:50].astype(np.int64).reshape(nb, 6, 2)
    face_type = rows[:, 50].astype(np.int64)
    face_global = rows[:, 51].astype(np.int64)
    sigma_face = rows[:, 52]
    dofs = rows[:, 53:53 + n].astype(np.int64)

    face_local = np.argmax(faces_ele == face_global[:, None], axis=1)
    # quadrature points of each local face in the master tetrahedron (hvfem.py:2686-2724)
    ref_pts = np.zeros((4, ng, 3))
    for f in range(4):
        o, a, b = REF_VERTICES[FACE_NODES[f]]
        ref_pts[f] = o + pts2[:, :1] * (a - o) + pts2[:, 1:2] * (b - o)
    rp = ref_pts[face_local]  # [nb, ng, 3]
    lam0 = 1.0 - rp[:, :, 0] - rp[:, :, 1] - rp[:, :, 2]
    z_pts = (lam0 * coord[:, 0, 2][:, None] + rp[:, :, 0] * coord[:, 1, 2][:, None]
             + rp[:, :, 1] * coord[:, 2, 2][:, None] + rp[:, :, 2] * coord[:, 3, 2][:, None])
    # a point of a top face may round a few ulp above z_max; the reference's interpolation then returns 0
    # instead of u(z_max) = 1 (see oracle/make_golden_mt.py): clip, the field is continuous there
    z_pts = np.minimum(z_pts, z_max)
    # conductivity profile for the 1-D problem: the faces of the right side (flag 3), at their centroids
    right = face_type == 3
    fn = FACE_NODES[face_local[right]]
    cz = np.take_along_axis(coord[right][:, :, 2], fn, axis=1).sum(axis=1) / 3.0
    u = eval_MT1D(z_max, z_min, 1.0, 0.0, sigma_face[right], cz, omega, mu, n_nodes_1d, 1, z_pts)

    # geometry and orientation of the boundary elements
    jac = coord[:, 1:4, :] - coord[:, 0:1, :]
    inv = np.linalg.inv(jac)
    normal = np.einsum("eab,eb->ea", inv, REF_NORMALS[face_local])
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    fnodes = FACE_NODES[face_local]
    v0 = np.take_along_axis(coord, fnodes[:, 0][:, None, None].repeat(3, axis=2), axis=1)[:, 0]
    v1 = np.take_along_axis(coord, fnodes[:, 1][:, None, None].repeat(3, axis=2), axis=1)[:, 0]
    v2 = np.take_along_axis(coord, fnodes[:, 2][:, None, None].repeat(3, axis=2), axis=1)[:, 0]
    det2d = np.linalg.norm(np.cross(v1 - v0, v2 - v0), axis=1)
    eo, fo = hvfem.computeElementOrientation_batch(edges_ele, nodes_ele, edges_nodes, edges_face)
    Jx, S = basis.local_to_expanded(p, eo, fo)  # [nb, n]
    # expanded functions at the quadrature points of the four local faces
    Nexp = np.stack([basis.evaluate_expanded(p, ref_pts[f])[0] for f in range(4)])  # [4, nexp, ng, 3]
    const = 1j * omega * mu
    modes = [{"x": 1, "y": 2}.get(pol, pol) for pol in polarizations]
    out = [np.zeros(int(total_dofs), dtype=np.complex128) for _ in modes]
    chunk = max(64, (1 << 23) // (n * ng * 3))  # boundary faces per pass: [faces, n, ng, 3] temporaries <= 64 MB
    for c0 in range(0, nb, chunk):
        sl = slice(c0, min(c0 + chunk, nb))
        Nref = Nexp[face_local[sl, None], Jx[sl]] * S[sl, :, None, None]              # [nc, n, ng, 3]
        Nreal = np.einsum("eab,ejgb->ejga", inv[sl], Nref)
        nrm = normal[sl]
        tang = Nreal - np.einsum("ejga,ea->ejg", Nreal, nrm)[..., None] * nrm[:, None, None, :]
        for b, mode in zip(out, modes):
            ex, ey, ez = neumann_excitation(face_type[sl], mode, u[sl])
            exc = np.stack([ex, ey, ez], axis=-1)  # [nc, ng, 3]
            contrib = np.einsum("ejga,ega,g->ej", tang, exc, wts) * det2d[sl, None] * const
            np.add.at(b, dofs[sl].reshape(-1), contrib.reshape(-1))
    return out

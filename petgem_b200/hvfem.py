"""Drop-in mirror of the ``petgem.hvfem`` functions on the hot path.

Same names, argument meaning and results as the reference module (file:line cited
per function) so that ``solver.py`` / ``postprocessing.py``-style callers keep
working; the per-element arithmetic runs on the B200 through the C ABI, batched.
Single-element calls are accepted (they become a batch of one); the batched entry
points (``*_batch``) are what the Solver uses.
"""
from __future__ import annotations

import numpy as np

from . import basis
from .common import Print


def computeConnectivityDOFS(elemsE, elemsF, Nord):
    """hvfem.py:15-98, vectorised; numbering bit-identical (edges, faces, interiors)."""
    elemsE, elemsF = np.asarray(elemsE, dtype=np.int64), np.asarray(elemsF, dtype=np.int64)
    T = elemsE.shape[0]
    nE, nF = int(elemsE.max()) + 1, int(elemsF.max()) + 1
    ne, nf, nv = basis.ndof_edge(Nord), basis.ndof_face(Nord), basis.ndof_volume(Nord)
    dof_edges = np.arange(nE * ne, dtype=np.int64).reshape(nE, ne)
    dof_faces = nE * ne + np.arange(nF * nf, dtype=np.int64).reshape(nF, nf)
    dof_volume = nE * ne + nF * nf + np.arange(T * nv, dtype=np.int64).reshape(T, nv)
    dof_connectivity = np.concatenate(
        [dof_edges[elemsE].reshape(T, 6 * ne), dof_faces[elemsF].reshape(T, 4 * nf), dof_volume], axis=1
    )
    return dof_connectivity, dof_edges, dof_faces, dof_volume, int(dof_connectivity.max()) + 1


def dofs_of_elements(elemsE_rows, elemsF_rows, elem_ids, nEdges, nFaces, Nord):
    """Rows of the dof table (hvfem.py:73-93) for selected elements, from the closed form of
    the numbering: edge e -> e*p+i, face f -> nE*p + f*p(p-1)+k, interior of element t after those."""
    eE = np.asarray(elemsE_rows, dtype=np.int64).reshape(-1, 6)
    eF = np.asarray(elemsF_rows, dtype=np.int64).reshape(-1, 4)
    t = np.asarray(elem_ids, dtype=np.int64).reshape(-1)
    ne, nf, nv = basis.ndof_edge(Nord), basis.ndof_face(Nord), basis.ndof_volume(Nord)
    parts = [(eE[:, :, None] * ne + np.arange(ne)).reshape(eE.shape[0], 6 * ne),
             (nEdges * ne + eF[:, :, None] * nf + np.arange(nf)).reshape(eF.shape[0], 4 * nf),
             nEdges * ne + nFaces * nf + t[:, None] * nv + np.arange(nv)[None, :]]
    return np.concatenate(parts, axis=1)


def computeJacobian(eleNodes):
    """hvfem.py:101-119 (host, one element: used by RHS / receiver code)."""
    eleNodes = np.asarray(eleNodes, dtype=np.float64)
    jacobian = eleNodes[1:4] - eleNodes[0]
    return jacobian, np.linalg.inv(jacobian)


_FACE_CODE = {12: 0, 31: 1, 23: 2, 32: 3, 13: 4, 21: 5}
_EDGE_LOCAL = ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))
_FACE_LOCAL_EDGES = ((0, 1, 2), (0, 4, 3), (1, 5, 4), (2, 5, 3))


def computeElementOrientation(edgesEle, nodesEle, edgesNodesEle, globalEdgesInFace):
    """hvfem.py:122-220 (host, one element).  The batched version is the CUDA kernel
    behind ``ElementData.geometry``."""
    edgesEle, nodesEle = np.asarray(edgesEle), np.asarray(nodesEle)
    edgesNodesEle, globalEdgesInFace = np.asarray(edgesNodesEle), np.asarray(globalEdgesInFace)
    eo = np.zeros(6, dtype=np.int64)
    for i, (a, b) in enumerate(_EDGE_LOCAL):
        eo[i] = int(nodesEle[a] == edgesNodesEle[i, 1] and nodesEle[b] == edgesNodesEle[i, 0])
    fo = np.zeros(4, dtype=np.int64)
    for i, le in enumerate(_FACE_LOCAL_EDGES):
        k1 = k2 = 0
        for k in range(3):
            if edgesEle[le[0]] == globalEdgesInFace[i, k]:
                k1 = k + 1
            if edgesEle[le[1]] == globalEdgesInFace[i, k]:
                k2 = k + 1
        fo[i] = _FACE_CODE.get(10 * k1 + k2, 0)
    return eo, fo


def computeElementOrientation_batch(edgesEle, nodesEle, edgesNodesEle, globalEdgesInFace):
    """computeElementOrientation for many elements at once (host numpy): edgesEle [m,6], nodesEle [m,4],
    edgesNodesEle [m,6,2], globalEdgesInFace [m,4,3] -> (edge_orientation [m,6], face_orientation [m,4])."""
    edgesEle, nodesEle = np.asarray(edgesEle), np.asarray(nodesEle)
    edgesNodesEle, gef = np.asarray(edgesNodesEle), np.asarray(globalEdgesInFace)
    a = np.array([e[0] for e in _EDGE_LOCAL])
    b = np.array([e[1] for e in _EDGE_LOCAL])
    eo = ((nodesEle[:, a] == edgesNodesEle[:, :, 1]) & (nodesEle[:, b] == edgesNodesEle[:, :, 0])).astype(np.int64)
    lut = np.zeros(34, dtype=np.int64)
    for code, val in _FACE_CODE.items():
        lut[code] = val
    fo = np.zeros((edgesEle.shape[0], 4), dtype=np.int64)
    for i, le in enumerate(_FACE_LOCAL_EDGES):
        # position (1-based, 0 = absent; the last match wins like the reference's loop) of the face's first
        # two local edges in the global face's edge triple
        m1 = edgesEle[:, le[0], None] == gef[:, i, :]
        m2 = edgesEle[:, le[1], None] == gef[:, i, :]
        k1 = np.where(m1.any(axis=1), 3 - np.argmax(m1[:, ::-1], axis=1), 0)
        k2 = np.where(m2.any(axis=1), 3 - np.argmax(m2[:, ::-1], axis=1), 0)
        fo[:, i] = lut[10 * k1 + k2]
    return eo, fo


def pack_orientation(edge_orientation, face_orientation):
    """(eo [...,6], fo [...,4]) -> uint32 code used by the kernels (include/petgem_b200.h)."""
    eo = np.asarray(edge_orientation, dtype=np.uint32)
    fo = np.asarray(face_orientation, dtype=np.uint32)
    code = np.zeros(eo.shape[:-1], dtype=np.uint32)
    for i in range(6):
        code |= (eo[..., i] & 1) << i
    for f in range(4):
        code |= (fo[..., f] & 7) << (6 + 3 * f)
    return code


def unpack_orientation(code):
    code = np.asarray(code).astype(np.uint32)
    eo = np.stack([(code >> i) & 1 for i in range(6)], axis=-1).astype(np.int64)
    fo = np.stack([(code >> (6 + 3 * f)) & 7 for f in range(4)], axis=-1).astype(np.int64)
    return eo, fo


def geometric_factors(jacobian, sigmaEle):
    """Packed symmetric factors (gK[6], gM[6]) of one or many elements from J and sigma.
    Used only to feed the kernels when the caller already holds J (API compatibility with
    computeElementalMatrices); the Solver computes them on the device from coordinates."""
    J = np.asarray(jacobian, dtype=np.float64)
    sig = np.asarray(sigmaEle, dtype=np.float64)
    det = np.linalg.det(J)
    Ji = np.linalg.inv(J)
    S = np.zeros(J.shape)
    S[..., 0, 0], S[..., 1, 1], S[..., 2, 2] = sig[..., 0], sig[..., 0], sig[..., 1]
    GK = J @ np.swapaxes(J, -1, -2) / det[..., None, None]
    GM = np.swapaxes(Ji, -1, -2) @ S @ Ji * det[..., None, None]
    idx = basis.SYM_PAIRS
    return np.stack([GK[..., a, b] for a, b in idx] + [GM[..., a, b] for a, b in idx], axis=-1)


def computeElementalMatrices(edge_orientation, face_orientation, jacobian, invjacob, Nord, sigmaEle):
    """hvfem.py:223-316 for one element -> (Me, Ke), evaluated on the GPU."""
    import torch

    from . import device as dv

    if Nord < 1 or Nord > 6:
        Print.master("        Nedelec order %s not supported (1..6)" % Nord)
        exit(-1)
    geo = torch.from_numpy(geometric_factors(jacobian, sigmaEle).reshape(1, 12)).to(dv._dev())
    code = torch.from_numpy(pack_orientation(edge_orientation, face_orientation).reshape(1).astype(np.int32)).to(geo.device)
    Me, Ke = dv.element_matrices(Nord, geo, code)
    return Me[0].cpu().numpy(), Ke[0].cpu().numpy()


def computeElementalMatrices_batch(elems, Nord):
    """All elements at once: ElementData -> (Me, Ke) device tensors [T,n,n]."""
    from . import device as dv

    geo, code = elems.geometry()
    return dv.element_matrices(Nord, geo, code)


# ---------------------------------------------------------------------------
# point evaluation of the basis (RHS and receivers; host, tiny)
# ---------------------------------------------------------------------------
def tetrahedronXYZToXiEtaZeta(eleNodes, points):
    """hvfem.py:2347-2490: affine inverse map (same result as the expanded formulas)."""
    eleNodes = np.asarray(eleNodes, dtype=np.float64)
    J = eleNodes[1:4] - eleNodes[0]
    pts = np.asarray(points, dtype=np.float64)
    if pts.ndim == 1:
        return np.linalg.solve(J.T, pts - eleNodes[0])
    return np.linalg.solve(J.T, (pts - eleNodes[0]).T).T


def shape3DETet(X, Nord, NoriE, NoriF):
    """hvfem.py:319-464 -> (NrdofE, ShapE [3,n], CurlE [3,n]) at master point X; Nord may be
    the 11-vector the reference passes (uniform order assumed, as the reference always does)."""
    p = int(np.max(Nord))
    N, C = basis.evaluate_expanded(p, np.asarray(X, dtype=np.float64).reshape(1, 3))
    J, S = basis.local_to_expanded(p, NoriE, NoriF)
    return J.size, (N[J, 0, :] * S[:, None]).T.copy(), (C[J, 0, :] * S[:, None]).T.copy()


def computeBasisFunctions(edge_orientation, face_orientation, jacobian, invjacob, Nord, points):
    """hvfem.py:2493-2550 -> (basis, curl_basis) [3, n, npoints] in the real element."""
    pts = np.asarray(points, dtype=np.float64)
    if pts.ndim == 1:
        pts = pts.reshape(1, 3)
    N, C = basis.evaluate_expanded(Nord, pts)
    J, S = basis.local_to_expanded(Nord, edge_orientation, face_orientation)
    Nref = N[J] * S[:, None, None]  # [n, npts, 3]
    Cref = C[J] * S[:, None, None]
    jac = np.asarray(jacobian, dtype=np.float64)
    basis_real = np.einsum("ab,jgb->ajg", np.asarray(invjacob, dtype=np.float64), Nref)
    curl_real = np.einsum("ba,jgb->ajg", jac, Cref) / np.linalg.det(jac)
    return basis_real, curl_real


def computeSourceVectorRotation(azimuth, dip):
    """hvfem.py:2303-2344."""
    a, b = np.deg2rad(azimuth), np.deg2rad(dip)
    M1 = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    M2 = np.array([[np.cos(b), 0.0, -np.sin(b)], [0.0, 1.0, 0.0], [np.sin(b), 0.0, np.cos(b)]])
    return M1 @ M2 @ np.array([1.0, 0.0, 0.0])

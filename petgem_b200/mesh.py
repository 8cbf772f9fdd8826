"""Tetrahedral mesh topology with PETGEM's numbering, vectorised.

Same function names, arguments and results as ``petgem/mesh.py`` (and the helpers
of ``petgem/vectors.py`` it relies on); the numbering is bit-identical:
edge / face ids are the lexicographic rank of the sorted node tuple
(``vectors.py:15-64``), the edges of a face are listed in the local-face order of
the lowest-index element containing it (``mesh.py:92-125``).  The reference walks
Python loops over faces and elements (minutes at 5 M tets); here every step is a
sort/unique on packed integer keys.
"""
from __future__ import annotations

import numpy as np

EDGE_NODES = np.array([[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]])  # mesh.py:27-32
FACE_NODES = np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3], [0, 2, 3]])  # mesh.py:70-73
FACE_EDGES = np.array([[0, 1, 2], [0, 4, 3], [1, 5, 4], [2, 5, 3]])  # mesh.py:116-123


def _rank_rows(rows: np.ndarray):
    """(unique sorted rows, inverse, first occurrence) of rows whose entries are sorted
    along axis 1; the rank is lexicographic like np.unique on a structured view."""
    rows = np.asarray(rows, dtype=np.int64)
    nn = int(rows.max()) + 1 if rows.size else 1
    k = rows.shape[1]
    if float(nn) ** k < 2.0 ** 62:
        key = rows[:, 0].copy()
        for c in range(1, k):
            key *= nn
            key += rows[:, c]
        ukey, first, inv = np.unique(key, return_index=True, return_inverse=True)
        uniq = np.empty((ukey.size, k), dtype=np.int64)
        rem = ukey
        for c in range(k - 1, 0, -1):
            uniq[:, c] = rem % nn
            rem = rem // nn
        uniq[:, 0] = rem
        return uniq, inv.reshape(-1), first
    uniq, first, inv = np.unique(rows, axis=0, return_index=True, return_inverse=True)
    return uniq, inv.reshape(-1), first


def computeEdges(elemsN, nElems):
    """mesh.py:18-53 -> (edges [nElems,6], edgesNodes [nEdges,2])."""
    elemsN = np.asarray(elemsN, dtype=np.int64)
    pairs = elemsN[:, EDGE_NODES].reshape(nElems * 6, 2)
    uniq, inv, _ = _rank_rows(np.sort(pairs, axis=1))
    return inv.reshape(nElems, 6), uniq


def computeFaces(elemsN, nElems):
    """mesh.py:56-89 -> (elemsF [nElems,4], facesN [nFaces,3]); facesN keeps the node order
    of the first element face that defines it, like the reference."""
    elemsN = np.asarray(elemsN, dtype=np.int64)
    tri = elemsN[:, FACE_NODES].reshape(nElems * 4, 3)
    _, inv, first = _rank_rows(np.sort(tri, axis=1))
    return inv.reshape(nElems, 4), tri[first]


def computeFacesEdges(elemsF, elemsE, nFaces, nElems):
    """mesh.py:92-125 -> facesE [nFaces,3]."""
    flat = np.asarray(elemsF).reshape(-1)
    _, first = np.unique(flat, return_index=True)  # first (lowest element) occurrence of each face
    t, k = first // 4, first % 4
    return np.asarray(elemsE)[t[:, None], FACE_EDGES[k]].astype(np.int64)


def computeBoundaryFaces(elemsF, facesN):
    """mesh.py:149-221 -> (bfacesN [3,nb], bFaces [nb], nb): faces owned by one element."""
    nFaces = int(np.max(elemsF)) + 1
    count = np.bincount(np.asarray(elemsF).reshape(-1), minlength=nFaces)
    bFaces = np.nonzero(count == 1)[0]
    return np.asarray(facesN)[bFaces].T.copy(), bFaces, bFaces.size


def computeBoundaryEdges(edgesN, bfacesN):
    """mesh.py:224-277 -> ids of the edges of boundary faces (ascending)."""
    edgesN = np.asarray(edgesN, dtype=np.int64)
    b = np.asarray(bfacesN, dtype=np.int64)
    pairs = np.sort(np.concatenate([b[[0, 1]].T, b[[1, 2]].T, b[[2, 0]].T], axis=0), axis=1)
    nn = int(edgesN.max()) + 1
    key_b = np.unique(pairs[:, 0] * nn + pairs[:, 1])
    key_all = edgesN[:, 0] * nn + edgesN[:, 1]  # ascending by construction
    return np.searchsorted(key_all, key_b)


def computeBoundaries(dof_connectivity, dof_edges, dof_faces, bEdges, bFaces, Nord):
    """mesh.py:280-321 -> (inner dofs, boundary dofs)."""
    parts = [np.asarray(dof_edges)[bEdges].reshape(-1)]
    if np.asarray(dof_faces).size:
        parts.append(np.asarray(dof_faces)[bFaces].reshape(-1))
    bd = np.concatenate(parts).astype(np.int64)
    total = int(np.max(dof_connectivity)) + 1
    mask = np.ones(total, dtype=bool)
    mask[bd] = False
    return np.nonzero(mask)[0], bd


def computeBoundaryElements(elemsF, bFaces, nFaces):
    """mesh.py:128-146 -> element owning each boundary face."""
    flat = np.asarray(elemsF).reshape(-1)
    _, first = np.unique(flat, return_index=True)
    bnd = (first // 4)[bFaces]
    return bnd, bnd.size


def computeFacePlane(nodes, bFaces, bFacesN):
    """mesh.py:324-428 -> plane flag 0..5 (bottom,left,front,right,back,top) per boundary face."""
    nodes = np.asarray(nodes)
    lo, hi = nodes.min(axis=0), nodes.max(axis=0)
    cen = nodes[np.asarray(bFacesN)].sum(axis=0) / 3.0  # [nb,3]
    ext = hi - lo
    dist = np.stack(
        [
            np.abs(cen[:, 2] - lo[2]) * ext[0] * ext[1],
            np.abs(cen[:, 0] - lo[0]) * ext[1] * ext[2],
            np.abs(cen[:, 1] - lo[1]) * ext[0] * ext[2],
            np.abs(cen[:, 0] - hi[0]) * ext[1] * ext[2],
            np.abs(cen[:, 1] - hi[1]) * ext[0] * ext[2],
            np.abs(cen[:, 2] - hi[2]) * ext[0] * ext[1],
        ],
        axis=1,
    )
    # a boundary-face centroid lies on exactly one side of the box: pick the plane it is on
    return np.argmin(dist, axis=1)

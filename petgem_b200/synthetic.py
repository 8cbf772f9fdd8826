"""Deterministic synthetic layered-earth tetrahedral meshes.

Structured box of m^3 hexahedra, each split into 6 Kuhn tetrahedra (T = 6 m^3):
m=55 -> 998 250 tets (C2), m=94 -> 4 983 504 (C3), m=69 -> 1 971 054 (C4),
m=32 -> 196 608 (C5).  Interior nodes are jittered, node ids randomly permuted and
the local vertex order of every tet randomly permuted (keeping detJ > 0 like Gmsh
output) so that all edge/face orientation codes of computeElementOrientation
(hvfem.py:122-220) occur.  Conductivity follows the case1 layering
(examples/case1/params.yaml:10-11) scaled to the box.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import mesh as pmesh

# case1: water 3.3333 / sediments 1 / oil 0.01 / sediments 1 (top to bottom)
CASE1_LAYERS = (
    # (z_top_fraction, z_bottom_fraction, sigma)   z fraction measured from the top of the box
    (0.0, 2.0 / 7.0, 3.3333),
    (2.0 / 7.0, 4.0 / 7.0, 1.0),
    (4.0 / 7.0, 4.2 / 7.0, 0.01),
    (4.2 / 7.0, 1.0, 1.0),
)


def kuhn_box(m: int, length: float = 3500.0, jitter: float = 0.2, seed: int = 1234, shuffle: bool = True):
    """-> (nodes [Nn,3] f64, elemsN [T,4] int64) of the jittered, shuffled Kuhn mesh."""
    rng = np.random.default_rng(seed)
    h = length / m
    g = np.arange(m + 1)
    I, J, K = np.meshgrid(g, g, g, indexing="ij")
    nodes = np.stack([I, J, K], axis=-1).reshape(-1, 3).astype(np.float64) * h
    interior = ((I > 0) & (I < m) & (J > 0) & (J < m) & (K > 0) & (K < m)).reshape(-1)
    nodes[interior] += rng.uniform(-jitter * h, jitter * h, size=(int(interior.sum()), 3))
    nodes[:, 2] -= length  # z in [-length, 0] like the marine CSEM examples

    def nid(i, j, k):
        return (i * (m + 1) + j) * (m + 1) + k

    c = np.arange(m)
    CI, CJ, CK = np.meshgrid(c, c, c, indexing="ij")
    CI, CJ, CK = CI.reshape(-1), CJ.reshape(-1), CK.reshape(-1)
    tets = []
    for perm in itertools.permutations(range(3)):
        off = np.zeros(3, dtype=np.int64)
        verts = [nid(CI, CJ, CK)]
        for axis in perm:
            off[axis] += 1
            verts.append(nid(CI + off[0], CJ + off[1], CK + off[2]))
        tets.append(np.stack(verts, axis=1))
    # element order: all 6 tets of a hex are consecutive (spatially coherent element ids)
    elemsN = np.stack(tets, axis=1).reshape(-1, 4)

    if shuffle:
        rng2 = np.random.default_rng(seed + 3087)  # 4321 for the default seed
        new_id = rng2.permutation(nodes.shape[0])
        nodes_p = np.empty_like(nodes)
        nodes_p[new_id] = nodes
        nodes = nodes_p
        elemsN = new_id[elemsN]
        # random local vertex order per element
        order = np.argsort(rng2.random(elemsN.shape), axis=1)
        elemsN = np.take_along_axis(elemsN, order, axis=1)
    # positive signed volume (hvfem.py:265 keeps the sign; Gmsh meshes are positive)
    X = nodes[elemsN]
    det = np.linalg.det(X[:, 1:] - X[:, :1])
    neg = det < 0
    elemsN[neg, 2], elemsN[neg, 3] = elemsN[neg, 3].copy(), elemsN[neg, 2].copy()
    return nodes, elemsN.astype(np.int64)


def layered_sigma(nodes, elemsN, layers=CASE1_LAYERS, vti_ratio: float = 1.0):
    """(sigma_h, sigma_v) per element from the centroid depth -> [T,2]."""
    z = nodes[elemsN][:, :, 2].mean(axis=1)
    ztop, zbot = nodes[:, 2].max(), nodes[:, 2].min()
    frac = (ztop - z) / (ztop - zbot)
    sig = np.full(z.shape, layers[-1][2])
    for a, b, s in layers:
        sig[(frac >= a) & (frac < b)] = s
    return np.stack([sig, sig * vti_ratio], axis=1)


def mesh_tables(nodes, elemsN):
    """All topology tables Preprocessing.run writes (preprocessing.py:134-312)."""
    T = elemsN.shape[0]
    elemsE, edgesNodes = pmesh.computeEdges(elemsN, T)
    elemsF, facesN = pmesh.computeFaces(elemsN, T)
    nFaces = facesN.shape[0]
    facesE = pmesh.computeFacesEdges(elemsF, elemsE, nFaces, T)
    bFacesN, bFaces, _ = pmesh.computeBoundaryFaces(elemsF, facesN)
    bEdges = pmesh.computeBoundaryEdges(edgesNodes, bFacesN)
    return dict(nodes=nodes, elemsN=elemsN, elemsE=elemsE, edgesNodes=edgesNodes, elemsF=elemsF, facesN=facesN,
                facesE=facesE, bFaces=bFaces, bEdges=bEdges, nEdges=edgesNodes.shape[0], nFaces=nFaces)


def csem_box(m: int, length: float = 3500.0, seed: int = 1234, vti_ratio: float = 1.0):
    """Synthetic layered-earth CSEM model: mesh tables + sigma + source/receivers."""
    nodes, elemsN = kuhn_box(m, length=length, seed=seed)
    tab = mesh_tables(nodes, elemsN)
    tab["sigma"] = layered_sigma(nodes, elemsN, vti_ratio=vti_ratio)
    seafloor = -length * 2.0 / 7.0
    tab["source"] = dict(frequency=2.0, position=np.array([length / 2, length / 2, seafloor + 25.0]), azimuth=0.0,
                         dip=0.0, current=1.0, length=1.0)
    xs = np.linspace(0.15 * length, 0.85 * length, 58)
    tab["receivers"] = np.stack([xs, np.full_like(xs, length / 2), np.full_like(xs, seafloor + 10.0)], axis=1)
    return tab

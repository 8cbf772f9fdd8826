"""Thin, vectorised stand-in for ``petgem/preprocessing.py`` (rank-0 serial stage).

Produces exactly the scratch files ``Solver.setup`` reads, in PETSc
binary format, from a Gmsh 2.2 ASCII mesh -- without meshio/h5py/petsc4py (absent here).
Topology and numbering come from ``petgem_b200.mesh`` / ``hvfem`` (bit-identical to the
reference), including the MT boundary-element table ``boundaryElements.dat``
(preprocessing.py:314-381) that feeds the MT right-hand side.
"""
from __future__ import annotations

from struct import error as struct_error

import numpy as np

from . import hvfem
from . import mesh as pmesh
from .common import Print, Timers
from .parallel import (MPIEnvironment, createSequentialDenseMatrixWithArray, createSequentialVectorWithArray,
                       writeParallelDenseMatrix, writePetscVector)


def read_gmsh22(path):
    """Gmsh 2.2 ASCII, tetrahedra only -> (points [Nn,3], tets [T,4] 0-based, physical tag [T])."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    try:
        i = lines.index("$Nodes")
        nn = int(lines[i + 1])
        pts = np.array([ln.split()[1:4] for ln in lines[i + 2:i + 2 + nn]], dtype=np.float64)
        i = lines.index("$Elements")
        ne = int(lines[i + 1])
    except ValueError:
        Print.master("     %s is not a Gmsh 2.2 ASCII mesh" % path)
        exit(-1)
    tets, tags = [], []
    for ln in lines[i + 2:i + 2 + ne]:
        f = ln.split()
        if len(f) > 2 and f[1] == "4":
            ntags = int(f[2])
            tags.append(int(f[3]))
            tets.append(f[3 + ntags:3 + ntags + 4])
    return pts, np.array(tets, dtype=np.int64) - 1, np.array(tags, dtype=np.int64)


def read_receivers(path, cols=3, rows=None):
    """Receiver positions [n, 3] (or another [rows, cols] float64 table such as the conductivity model):
    .npy / text, or the 'data' dataset of PETGEM's .h5 files (preprocessing.py:399-407).  h5py is used when it is
    installed; otherwise the file is parsed by h5lite.read_classic (the layout h5py writes by default: symbol-table
    root group, contiguous dataset).  The shape is checked against `cols` (and `rows` when known): a table of the
    wrong shape fails loudly instead of being reinterpreted."""
    if path.endswith(".npy"):
        data = np.load(path)
    elif path.endswith(".h5"):
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File(path, "r") as fh:
                data = fh["data"][()]
        else:
            from . import h5lite
            try:
                data = h5lite.read_classic(path)["data"]
            except (ValueError, KeyError, IndexError, struct_error) as err:
                Print.master("     %s: cannot be read without h5py (%s); install h5py or convert the file to .npy"
                             % (path, err))
                exit(-1)
    else:
        data = np.loadtxt(path)
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1 and data.size % cols == 0:
        data = data.reshape(-1, cols)
    if data.ndim != 2 or data.shape[1] < cols or (rows is not None and data.shape[0] != rows):
        Print.master("     %s: expected a table of %s x %d values, found shape %s"
                     % (path, "n" if rows is None else str(rows), cols, data.shape))
        exit(-1)
    return data[:, :cols]


def locate_points(nodes, elemsN, points, tol=1.0e-12):
    """Containing element of each point (lowest index), -1 if none: vectorised stand-in for
    Delaunay.find_simplex with the mesh connectivity (preprocessing.py:414-420)."""
    points = np.atleast_2d(points)
    X0 = nodes[elemsN[:, 0]]
    Jt = np.stack([nodes[elemsN[:, k]] - X0 for k in (1, 2, 3)], axis=2)
    lo, hi = nodes[elemsN].min(axis=1), nodes[elemsN].max(axis=1)
    out = np.full(points.shape[0], -1, dtype=np.int64)
    for i, pt in enumerate(points):
        cand = np.nonzero(((pt >= lo - 1e-9) & (pt <= hi + 1e-9)).all(axis=1))[0]
        if cand.size == 0:
            continue
        loc = np.linalg.solve(Jt[cand], np.broadcast_to(pt - X0[cand], (cand.size, 3))[..., None])[..., 0]
        ok = (loc >= -tol).all(axis=1) & (1.0 - loc.sum(axis=1) >= -tol)
        if ok.any():
            out[i] = cand[np.argmax(ok)]
    return out


class Preprocessing():
    """Class for preprocessing."""

    def __init__(self):
        return

    def run(self, inputSetup):
        Timers()["Preprocessing"].start()
        parEnv = MPIEnvironment()
        if parEnv.rank == 0:
            self._run_master(inputSetup)
        try:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.barrier()
        except Exception:
            pass
        Timers()["Preprocessing"].stop()

    def _run_master(self, inputSetup):
        model, run, output = inputSetup.model, inputSetup.run, inputSetup.output
        out_dir = output.get('directory_scratch')
        p = run.get('nord')
        mode = model.get('mode')
        data_model = model.get(mode)
        points, cells, tags = read_gmsh22(model.get('mesh'))
        nElems = cells.shape[0]

        def dump(name, arr):
            arr = np.asarray(arr, dtype=np.float64).reshape(nElems, -1)
            writeParallelDenseMatrix(out_dir + '/' + name,
                                     createSequentialDenseMatrixWithArray(arr.shape[0], arr.shape[1], arr))

        Print.master('     Nodal coordinates')
        dump('nodes.dat', points[cells])
        Print.master('     Mesh connectivity')
        dump('meshConnectivity.dat', cells)
        Print.master('     Edges connectivity')
        elemsE, edgesNodes = pmesh.computeEdges(cells, nElems)
        dump('edges.dat', elemsE)
        dump('edgesNodes.dat', edgesNodes[elemsE])
        Print.master('     Faces connectivity')
        elemsF, facesN = pmesh.computeFaces(cells, nElems)
        nFaces = facesN.shape[0]
        dump('faces.dat', elemsF)
        Print.master('     Faces-edges connectivity')
        facesE = pmesh.computeFacesEdges(elemsF, elemsE, nFaces, nElems)
        dump('facesEdges.dat', facesE[elemsF])
        Print.master('     DOFs connectivity')
        dofs, dof_edges, dof_faces, _, total_num_dofs = hvfem.computeConnectivityDOFS(elemsE, elemsF, p)
        dump('dofs.dat', dofs)
        Print.master('     Conductivity model')
        i_model = data_model.get('sigma')
        if run.get('conductivity_from_file'):
            sig_file = i_model.get('file')
            conductivityModel = np.load(sig_file) if sig_file.endswith('.npy') else read_receivers(sig_file, cols=2, rows=nElems)
        else:
            elemsS = tags - 1
            conductivityModel = np.stack([np.asarray(i_model.get('horizontal'), dtype=np.float64)[elemsS],
                                          np.asarray(i_model.get('vertical'), dtype=np.float64)[elemsS]], axis=1)
        dump('conductivityModel.dat', conductivityModel)
        Print.master('     Boundaries')
        bFacesN, bFaces, nbFaces = pmesh.computeBoundaryFaces(elemsF, facesN)
        if mode == 'csem':
            bEdges = pmesh.computeBoundaryEdges(edgesNodes, bFacesN)
            _, bd = pmesh.computeBoundaries(dofs, dof_edges, dof_faces, bEdges, bFaces, p)
            writePetscVector(out_dir + '/boundaries.dat', createSequentialVectorWithArray(bd.astype(np.float64)))
            src = np.asarray(data_model.get('source').get('position'), dtype=np.float64)
            srcElem = int(locate_points(points, cells, src)[0])
            if srcElem < 0:
                Print.master('        Source no located in the computational domain. Please, verify source position or improve the mesh quality.')
                exit(-1)
            data_source = np.concatenate([cells[srcElem], points[cells[srcElem]].reshape(-1), elemsF[srcElem],
                                          facesE[elemsF[srcElem]].reshape(-1), elemsE[srcElem],
                                          edgesNodes[elemsE[srcElem]].reshape(-1), dofs[srcElem]]).astype(np.float64)
            writePetscVector(out_dir + '/source.dat', createSequentialVectorWithArray(data_source))
        else:
            # one row per boundary face with everything the MT right-hand side needs (preprocessing.py:314-381)
            planeFace = pmesh.computeFacePlane(points, bFaces, bFacesN)
            bElems, numbElems = pmesh.computeBoundaryElements(elemsF, bFaces, nFaces)
            if nbFaces != numbElems:
                Print.master('     Number of boundary faces is not consistent.')
                exit(-1)
            data_boundaries = boundary_element_rows(points, cells, elemsE, edgesNodes, elemsF, facesE, dofs,
                                                    conductivityModel, bFaces, bElems, planeFace)
            writeParallelDenseMatrix(out_dir + '/boundaryElements.dat',
                                     createSequentialDenseMatrixWithArray(data_boundaries.shape[0],
                                                                          data_boundaries.shape[1], data_boundaries))
        valence = np.array([50, 200, 400, 800, 1400, 2500])  # preprocessing.py:470-473
        writePetscVector(out_dir + '/nnz.dat',
                         createSequentialVectorWithArray(np.full(total_num_dofs, valence[p - 1], dtype=np.float64)))
        # tables the post-processing needs (the reference recomputes them from the mesh)
        np.savez(out_dir + '/mesh_tables.npz', nodes=points, elemsN=cells, elemsE=elemsE, edgesNodes=edgesNodes,
                 elemsF=elemsF, facesE=facesE, dofs=dofs)
        Print.master('     Number of elements: %d, dofs: %d' % (nElems, total_num_dofs))


def boundary_element_rows(points, cells, elemsE, edgesNodes, elemsF, facesE, dofs, conductivityModel, bFaces,
                          bElems, planeFace):
    """Rows of boundaryElements.dat (preprocessing.py:326-367), one per boundary face: nodes 0:4,
    coordinates 4:16, faces 16:20, edges of the faces 20:32, edges 32:38, nodes of the edges 38:50, plane
    flag 50, global face id 51, horizontal sigma 52, dofs 53:."""
    points, cells = np.asarray(points), np.asarray(cells)
    t = np.asarray(bElems, dtype=np.int64)
    nb, n = t.size, np.asarray(dofs).shape[1]
    rows = np.zeros((nb, 53 + n), dtype=np.float64)
    rows[:, 0:4] = cells[t]
    rows[:, 4:16] = points[cells[t]].reshape(nb, 12)
    rows[:, 16:20] = np.asarray(elemsF)[t]
    rows[:, 20:32] = np.asarray(facesE)[np.asarray(elemsF)[t]].reshape(nb, 12)
    rows[:, 32:38] = np.asarray(elemsE)[t]
    rows[:, 38:50] = np.asarray(edgesNodes)[np.asarray(elemsE)[t]].reshape(nb, 12)
    rows[:, 50] = planeFace
    rows[:, 51] = bFaces
    rows[:, 52] = np.asarray(conductivityModel)[t, 0]
    rows[:, 53:] = np.asarray(dofs)[t]
    return rows


def unitary_test():
    """Unitary test for preprocessing.py script."""

#!/usr/bin/env python3
"""Golden vectors for the MT right-hand side, recorded from the UNMODIFIED reference.

Test infrastructure, container-only (imports /root/reference through oracle/refshim.py).
``petgem/solver.py`` itself cannot be imported (petsc4py), so its MT branch (solver.py:318-512) is
driven here call by call with the reference's own functions -- compute2DGaussPoints,
transform2Dto3DInReferenceElement, getRealFromReference, getFaceByLocalNodes, mt1d.eval_MT1D,
computeJacobian, computeElementOrientation, getNormalVector, get2DJacobDet, getNeumannBCface,
computeBasisFunctionsReferenceElement -- on the boundary rows that preprocessing.py:326-367 would
write for tests/data/test_mesh.msh (mesh.computeBoundaryFaces / computeFacePlane /
computeBoundaryElements).  The 1-D problem uses N1D nodes instead of the hard-wired 1e6
(solver.py:403) to keep the reference's Python loops short; the product takes the same parameter.

    python oracle/make_golden_mt.py           ->  tests/golden/mt_rhs.npz
    python oracle/make_golden_mt.py --rules   ->  tests/golden/triangle_rules.npz
    python oracle/make_golden_mt.py --impedance -> tests/golden/mt_impedance.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import refshim  # noqa: E402

N1D = 4001
FREQ = 2.0
SIGMA_BY_TAG = np.array([1.0, 0.01, 1.0, 3.3333])


def main():
    hvfem, mesh, _ = refshim.load()
    from petgem import mt1d

    topo = dict(np.load(os.path.join(GOLD, "test_mesh_topology.npz")))
    nodes = topo["nodes"]
    elemsN = topo["elemsN"].astype(np.int64)
    elemsE, elemsF = topo["elemsE"].astype(np.int64), topo["elemsF"].astype(np.int64)
    edgesNodes, facesE, facesN = topo["edgesNodes"].astype(np.int64), topo["facesE"].astype(np.int64), topo["facesN"].astype(np.int64)
    nF = facesN.shape[0]
    sigma_h = SIGMA_BY_TAG[topo["tags"] - 1]
    omega, mu = 2 * np.pi * FREQ, 4e-7 * np.pi

    bFacesN, bFaces, nbFaces = mesh.computeBoundaryFaces(elemsF, facesN)
    planeFace = mesh.computeFacePlane(nodes, bFaces, bFacesN)
    bElems, numbElems = mesh.computeBoundaryElements(elemsF, bFaces, nF)
    assert nbFaces == numbElems
    out = dict(planeFace=planeFace.astype(np.int8), bElems=bElems.astype(np.int32), n1d=N1D, freq=FREQ,
               sigma_by_tag=SIGMA_BY_TAG)

    # 1-D solver alone: solution at all nodes and at scattered points, two sizes
    rng = np.random.default_rng(5)
    x0 = np.sort(rng.uniform(nodes[:, 2].min() + 20, nodes[:, 2].max() - 20, size=40))
    s0 = rng.choice(SIGMA_BY_TAG, size=40)
    pts = rng.uniform(nodes[:, 2].min(), nodes[:, 2].max(), size=(7, 5))
    for n1 in (501, N1D):
        u_nodes = mt1d.eval_MT1D(nodes[:, 2].max(), nodes[:, 2].min(), 1.0, 0.0, s0, x0, omega, mu, n1, 1, None)
        u_pts = mt1d.eval_MT1D(nodes[:, 2].max(), nodes[:, 2].min(), 1.0, 0.0, s0, x0, omega, mu, n1, 1, pts)
        out["mt1d_u_nodes_%d" % n1], out["mt1d_u_pts_%d" % n1] = u_nodes, u_pts
    out["mt1d_x0"], out["mt1d_sigma0"], out["mt1d_pts"] = x0, s0, pts

    for p in (1, 2, 3):
        n = p * (p + 2) * (p + 3) // 2
        dofs, _, _, _, total = hvfem.computeConnectivityDOFS(elemsE, elemsF, p)
        # boundary rows, preprocessing.py:326-367
        rows = np.zeros((nbFaces, 53 + n))
        for i in range(nbFaces):
            t = bElems[i]
            rows[i, 0:4] = elemsN[t]
            rows[i, 4:16] = nodes[elemsN[t]].flatten()
            rows[i, 16:20] = elemsF[t]
            rows[i, 20:32] = facesE[elemsF[t]].flatten()
            rows[i, 32:38] = elemsE[t]
            rows[i, 38:50] = edgesNodes[elemsE[t]].flatten()
            rows[i, 50] = planeFace[i]
            rows[i, 51] = bFaces[i]
            rows[i, 52] = sigma_h[t]
            rows[i, 53:] = dofs[t]
        # solver.py:322-404
        g2, Wi = hvfem.compute2DGaussPoints(2 * p)
        ng = g2.shape[0]
        interp = np.zeros((nbFaces, ng))
        cz, sz = [], []
        for i in range(nbFaces):
            coord = rows[i, 4:16].reshape(4, 3)
            facesEle = rows[i, 16:20].astype(int)
            floc = np.where(facesEle == int(rows[i, 51]))[0][0]
            for j in range(ng):
                g3 = hvfem.transform2Dto3DInReferenceElement(g2[j, :], floc)
                interp[i, j] = hvfem.getRealFromReference(g3, coord)[2]
            if int(rows[i, 50]) == 3:
                nf = hvfem.getFaceByLocalNodes(floc)
                cz.append(np.sum(coord[nf], axis=0)[2] / 3.0)
                sz.append(rows[i, 52])
        za, zb = nodes[elemsN].reshape(-1, 3)[:, 2].max(), nodes[elemsN].reshape(-1, 3)[:, 2].min()
        u_raw = mt1d.eval_MT1D(za, zb, 1.0, 0.0, np.asarray(sz), np.asarray(cz), omega, mu, N1D, 1, interp)
        # Rounding artefact of the reference: a Gauss point of a TOP face whose z comes out a few ulp above
        # z_max is never assigned by mt1d.linearInterp1D and gets u = 0 instead of u(z_max) = 1.  Which
        # points are hit depends on the last bits of the quadrature table, so it cannot be reproduced by an
        # independent implementation; the vectors recorded for the parity test therefore feed the same
        # reference function with the points clipped to z_max ("clip"), and the raw outcome is kept too.
        u_clip = mt1d.eval_MT1D(za, zb, 1.0, 0.0, np.asarray(sz), np.asarray(cz), omega, mu, N1D, 1,
                                np.minimum(interp, za))
        out["artefact_points_p%d" % p] = int(np.count_nonzero(u_raw != u_clip))
        const = 1j * omega * mu
        # solver.py:407-512
        variants = [("x", 1, u_clip, ""), ("y", 2, u_clip, "")]
        if p <= 2:  # the raw (artefact) outcome is recorded at the two lowest orders only
            variants += [("x", 1, u_raw, "_raw"), ("y", 2, u_raw, "_raw")]
        for pol, mode, u, tag in variants:
            b = np.zeros(total, dtype=np.complex128)
            for i in range(nbFaces):
                nodesEle = rows[i, 0:4].astype(int)
                coord = rows[i, 4:16].reshape(4, 3)
                facesEle = rows[i, 16:20].astype(int)
                edgesFace = rows[i, 20:32].astype(int).reshape(4, 3)
                edgesEle = rows[i, 32:38].astype(int)
                edgesNodesEle = rows[i, 38:50].astype(int).reshape(6, 2)
                ftype = int(rows[i, 50])
                floc = np.where(facesEle == int(rows[i, 51]))[0][0]
                dofsEle = rows[i, 53:].astype(int)
                _, invj = hvfem.computeJacobian(coord)
                eo, fo = hvfem.computeElementOrientation(edgesEle, nodesEle, edgesNodesEle, edgesFace)
                nv = hvfem.getNormalVector(floc, invj)
                nu = nv / np.linalg.norm(nv)
                det2 = hvfem.get2DJacobDet(coord, floc)
                ex, ey, ez = hvfem.getNeumannBCface(ftype, mode, u)
                contrib = np.zeros(n, dtype=np.complex128)
                for k in range(ng):
                    g3 = hvfem.transform2Dto3DInReferenceElement(g2[k, :], floc)
                    bases = hvfem.computeBasisFunctionsReferenceElement(eo, fo, p, g3)
                    real = np.matmul(invj, bases[:, :, 0])
                    exc = np.array([ex[i, k], ey[i, k], ez[i, k]], dtype=np.complex128)
                    integrand = np.zeros(n, dtype=np.complex128)
                    for l in range(n):
                        tang = np.cross(np.cross(nu, real[:, l]), nu)
                        integrand[l] = np.dot(tang, exc)
                    contrib += Wi[k] * integrand * det2
                b[dofsEle] += contrib * const
            nz = np.nonzero(b)[0]
            out["b_%s_p%d%s_idx" % (pol, p, tag)] = nz.astype(np.int32)
            out["b_%s_p%d%s_val" % (pol, p, tag)] = b[nz]
            print("p=%d pol=%s%s: %d boundary faces, %d nonzero entries, |b| = %.6e" % (p, pol, tag, nbFaces, nz.size, np.linalg.norm(b)))
        out["u_p%d" % p] = u_clip
        out["gauss2d_p%d_pts" % p], out["gauss2d_p%d_w" % p] = np.asarray(g2), np.asarray(Wi)
    np.savez_compressed(os.path.join(GOLD, "mt_rhs.npz"), **out)
    print("wrote", os.path.join(GOLD, "mt_rhs.npz"))


def impedance():
    """computeImpedance of the reference's postprocessing.py on random fields.  The module itself cannot be
    imported (h5py, meshio, petsc4py): the one pure-numpy function is compiled from its source text."""
    import ast

    refshim.load()
    path = os.path.join(refshim.REFERENCE_ROOT, "petgem", "postprocessing.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "computeImpedance"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    rng = np.random.default_rng(11)
    fields = [rng.normal(size=(9, 6)) + 1j * rng.normal(size=(9, 6)) for _ in range(2)]
    omega, mu = 2 * np.pi * FREQ, 4e-7 * np.pi
    res, phase, tipper, imp = ns["computeImpedance"](fields, omega, mu)
    np.savez_compressed(os.path.join(GOLD, "mt_impedance.npz"), f0=fields[0], f1=fields[1], omega=omega, mu=mu,
                        apparent_resistivity=np.stack(res), phase=np.stack(phase), tipper=np.stack(tipper),
                        impedance=np.stack(imp))
    print("wrote", os.path.join(GOLD, "mt_impedance.npz"))


def rules():
    """The reference's 2-D rules of order 2p, p = 1..6 (hvfem.compute2DGaussPoints) ->
    tests/golden/triangle_rules.npz"""
    hvfem, _, _ = refshim.load()
    out = {}
    for deg in (2, 4, 6, 8, 10, 12):
        pts, w = hvfem.compute2DGaussPoints(deg)
        out["pts_%d" % deg], out["w_%d" % deg] = np.asarray(pts, dtype=np.float64), np.asarray(w, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "triangle_rules.npz"), **out)
    print("wrote", os.path.join(GOLD, "triangle_rules.npz"))


if __name__ == "__main__":
    if "--rules" in sys.argv:
        rules()
    elif "--impedance" in sys.argv:
        impedance()
    else:
        main()

"""Generate tests/golden/*.npz from the UNMODIFIED reference (container-only).

Run once in the build container:  python oracle/make_golden.py [--slow]
It imports /root/reference through oracle/refshim.py, calls the reference's own
functions and stores inputs + outputs as small fixtures.  The fixtures travel to
the GPU box; this script and the reference do not need to.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
REF_DATA = os.path.join(refshim.REFERENCE_ROOT, "tests", "data")


def read_msh22(path):
    """Minimal Gmsh 2.2 ASCII reader (tets only), 0-based node ids."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    i = lines.index("$Nodes")
    nn = int(lines[i + 1])
    nodes = np.array([[float(v) for v in ln.split()[1:4]] for ln in lines[i + 2:i + 2 + nn]])
    ids = np.array([int(ln.split()[0]) for ln in lines[i + 2:i + 2 + nn]])
    assert (ids == np.arange(1, nn + 1)).all()
    i = lines.index("$Elements")
    ne = int(lines[i + 1])
    tets, tags = [], []
    for ln in lines[i + 2:i + 2 + ne]:
        f = [int(v) for v in ln.split()]
        if f[1] == 4:
            ntags = f[2]
            tags.append(f[3])
            tets.append(f[3 + ntags:3 + ntags + 4])
    return nodes, np.array(tets, dtype=np.int64) - 1, np.array(tags, dtype=np.int64)


def read_petsc_mat(path):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype=">i4", count=4)
    assert hdr[0] == 1211216
    M, N, nnz = int(hdr[1]), int(hdr[2]), int(hdr[3])
    off = 16
    rowlens = np.frombuffer(raw, dtype=">i4", count=M, offset=off).astype(np.int64)
    off += 4 * M
    cols = np.frombuffer(raw, dtype=">i4", count=nnz, offset=off).astype(np.int32)
    off += 4 * nnz
    vals = np.frombuffer(raw, dtype=">f8", count=2 * nnz, offset=off).astype(np.float64).view(np.complex128)
    return np.concatenate([[0], np.cumsum(rowlens)]), cols, vals, (M, N)


def read_petsc_vec(path):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype=">i4", count=2)
    assert hdr[0] == 1211214
    return np.frombuffer(raw, dtype=">f8", count=2 * int(hdr[1]), offset=8).astype(np.float64).view(np.complex128)


def random_tet(rng, negative=False):
    while True:
        X = rng.normal(size=(4, 3)) * rng.uniform(0.5, 50.0)
        J = X[1:] - X[0]
        d = np.linalg.det(J)
        if abs(d) > 0.05 * np.abs(J).max() ** 3:
            break
    if (d < 0) != negative:
        X[[2, 3]] = X[[3, 2]]
    return X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slow", action="store_true", help="also p=5,6 element matrices (minutes)")
    args = ap.parse_args()
    hvfem, mesh, vectors = refshim.load()
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # --- quadrature facts the oracle/product rely on ------------------------------
    from math import factorial as fact
    quad = {}
    for p in range(1, 7):
        g, w = hvfem.compute3DGaussPoints(2 * p)
        worst = 0.0
        deg = 2 * p
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                for c in range(deg + 1 - a - b):
                    ex = fact(a) * fact(b) * fact(c) / fact(a + b + c + 3)
                    worst = max(worst, abs((w * g[:, 0] ** a * g[:, 1] ** b * g[:, 2] ** c).sum() - ex) / ex)
        quad[p] = (g.shape[0], w.sum(), worst)
        print("quadrature p=%d npts=%d sum=%.17g exactness(deg 2p)=%.2e" % (p, *quad[p]))
    np.savez(os.path.join(GOLD, "quadrature_facts.npz"),
             npts=np.array([quad[p][0] for p in range(1, 7)]),
             wsum=np.array([quad[p][1] for p in range(1, 7)]),
             exactness=np.array([quad[p][2] for p in range(1, 7)]))

    # --- shape functions -----------------------------------------------------------
    out = {}
    for p in range(1, 7):
        n = p * (p + 2) * (p + 3) // 2
        pts = rng.dirichlet(np.ones(4), size=3)[:, :3]
        eos = np.concatenate([np.zeros((1, 6), int), np.ones((1, 6), int), rng.integers(0, 2, (6, 6))])
        fos = np.concatenate([np.zeros((1, 4), int), np.array([[1, 2, 3, 4]]), np.array([[5, 0, 5, 3]]),
                              rng.integers(0, 6, (5, 4))])
        shp = np.zeros((len(eos), len(pts), 2, 3, n))
        for c, (eo, fo) in enumerate(zip(eos, fos)):
            for ip, X in enumerate(pts):
                nd, S, C = hvfem.shape3DETet(X, np.ones(11, dtype=int) * p, eo, fo)
                assert nd == n
                shp[c, ip, 0], shp[c, ip, 1] = S[:, :n], C[:, :n]
        out["pts_%d" % p], out["eo_%d" % p], out["fo_%d" % p], out["shape_%d" % p] = pts, eos, fos, shp
    np.savez_compressed(os.path.join(GOLD, "hvfem_shape.npz"), **out)
    print("shape functions done")

    # --- element matrices ------------------------------------------------------------
    counts = {1: 24, 2: 12, 3: 6, 4: 2}
    if args.slow:
        counts.update({5: 1, 6: 1})
    for p, cnt in counts.items():
        n = p * (p + 2) * (p + 3) // 2
        coords = np.zeros((cnt, 4, 3))
        eos = rng.integers(0, 2, (cnt, 6))
        fos = rng.integers(0, 6, (cnt, 4))
        sig = np.zeros((cnt, 2))
        Me = np.zeros((cnt, n, n))
        Ke = np.zeros((cnt, n, n))
        t0 = time.time()
        for i in range(cnt):
            coords[i] = random_tet(rng, negative=(i % 3 == 2))
            sh = 10.0 ** rng.uniform(-3, 1)
            sig[i] = (sh, sh if i % 2 == 0 else sh * rng.uniform(0.2, 0.9))
            if i < 6:
                fos[i] = i  # every face code on every face at least once
            J, Ji = hvfem.computeJacobian(coords[i])
            Me[i], Ke[i] = hvfem.computeElementalMatrices(eos[i], fos[i], J, Ji, p, sig[i])
        print("element matrices p=%d x%d: %.1f s" % (p, cnt, time.time() - t0))
        np.savez_compressed(os.path.join(GOLD, "hvfem_elemental_p%d.npz" % p),
                            coords=coords, eo=eos, fo=fos, sigma=sig, Me=Me, Ke=Ke)

    # --- mesh topology / numbering on the reference's test mesh ---------------------
    nodes, elemsN, tags = read_msh22(os.path.join(REF_DATA, "test_mesh.msh"))
    T = elemsN.shape[0]
    elemsE, edgesNodes = mesh.computeEdges(elemsN, T)
    elemsF, facesN = mesh.computeFaces(elemsN, T)
    nE, nF = edgesNodes.shape[0], facesN.shape[0]
    facesE = mesh.computeFacesEdges(elemsF, elemsE, nF, T)
    # known answers of the reference's own tests/test_mesh.py:31-36
    assert (T, nodes.shape[0], nF, nE) == (9453, 2163, 20039, 12748)
    assert list(elemsE[0]) == [10591, 10600, 10831, 10832, 10601, 11465]
    assert list(elemsF[0]) == [17369, 17370, 17400, 17977]
    bFacesN, bFaces, nbF = mesh.computeBoundaryFaces(elemsF, facesN)
    bEdges = mesh.computeBoundaryEdges(edgesNodes, bFacesN)
    orient = np.zeros((T, 10), dtype=np.uint8)
    for t in range(T):
        eo, fo = hvfem.computeElementOrientation(elemsE[t], elemsN[t], edgesNodes[elemsE[t]], facesE[elemsF[t]])
        orient[t, :6], orient[t, 6:] = eo, fo
    topo = dict(nodes=nodes, elemsN=elemsN.astype(np.int32), tags=tags.astype(np.int8),
                elemsE=elemsE.astype(np.int32), edgesNodes=edgesNodes.astype(np.int32),
                elemsF=elemsF.astype(np.int32), facesN=facesN.astype(np.int32),
                facesE=facesE.astype(np.int32), bFaces=bFaces.astype(np.int32),
                bEdges=bEdges.astype(np.int32), orient=orient)
    for p in (1, 2, 3):
        dofs, dof_edges, dof_faces, _, total = hvfem.computeConnectivityDOFS(elemsE, elemsF, p)
        _, bd = mesh.computeBoundaries(dofs, dof_edges, dof_faces, bEdges, bFaces, p)
        # full tables are a pure function of (elemsE, elemsF, p): keep samples + totals
        sel = np.array([0, 1, 17, 4096, T - 1])
        topo["dofs_rows_p%d" % p] = dofs[sel]
        topo["dofs_sel"] = sel
        topo["total_dofs_p%d" % p] = total
        topo["dofs_sum_p%d" % p] = dofs.sum(axis=0)
        topo["boundary_dofs_p%d" % p] = bd.astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "test_mesh_topology.npz"), **topo)
    print("topology done: T=%d E=%d F=%d bF=%d bE=%d" % (T, nE, nF, nbF, bEdges.size))

    # receivers of case1 (58 x 3 float64 at byte offset 2048)
    raw = open(os.path.join(REF_DATA, "receiver_pos.h5"), "rb").read()
    rec = np.frombuffer(raw, dtype="<f8", count=58 * 3, offset=2048).reshape(58, 3)
    assert np.allclose(rec[:, 1], 1750.0) and np.allclose(rec[:, 2], -990.0)
    np.save(os.path.join(GOLD, "case1_receivers.npy"), rec)
    # the file itself (3.4 KB of DATA written by h5py, no source code) as a fixture of the classic-layout HDF5
    # reader petgem_b200/h5lite.read_classic (tests/test_host.py)
    import shutil
    shutil.copyfile(os.path.join(REF_DATA, "receiver_pos.h5"), os.path.join(GOLD, "receiver_pos_reference.h5"))

    # --- reference element loop on the test mesh (solver.py:191-224) -----------------
    sig_table = np.array([1.0, 0.01, 1.0, 3.3333])  # examples/case1 params.yaml:10-11
    omega, mu = 2.0 * np.pi * 2.0, 4e-7 * np.pi
    for p, nsel in ((1, T), (2, 160), (3, 24)):
        n = p * (p + 2) * (p + 3) // 2
        sel = np.arange(T) if nsel == T else np.sort(rng.choice(T, nsel, replace=False))
        Ae = np.zeros((sel.size, n, n), dtype=np.complex128)
        t0 = time.time()
        for c, t in enumerate(sel):
            J, Ji = hvfem.computeJacobian(nodes[elemsN[t]])
            eo, fo = orient[t, :6].astype(int), orient[t, 6:].astype(int)
            s = sig_table[tags[t] - 1]
            M, K = hvfem.computeElementalMatrices(eo, fo, J, Ji, p, np.array([s, s]))
            Ae[c] = K - 1j * omega * mu * M
        print("reference element loop p=%d x%d: %.1f s" % (p, sel.size, time.time() - t0))
        if p == 1:
            # global A = sum of cliques; store y = A x for a seeded x and the diagonal
            dofs = elemsE  # p=1: dof id == edge id (hvfem.py:54-59)
            x = rng.normal(size=nE) + 1j * rng.normal(size=nE)
            y = np.zeros(nE, dtype=np.complex128)
            diag = np.zeros(nE, dtype=np.complex128)
            for c in range(T):
                y[dofs[c]] += Ae[c] @ x[dofs[c]]
                diag[dofs[c]] += np.diag(Ae[c])
            np.savez_compressed(os.path.join(GOLD, "test_mesh_system_p1.npz"), x=x, y=y, diag=diag,
                                Ae_first=Ae[:64], omega=omega, mu=mu, sigma=sig_table)
        else:
            np.savez_compressed(os.path.join(GOLD, "test_mesh_elements_p%d.npz" % p), sel=sel.astype(np.int32),
                                Ae=Ae, omega=omega, mu=mu, sigma=sig_table)

    # --- the reference's PETSc fixture system (tests/test_petsc.py:15-35) ------------
    rowptr, cols, vals, shape = read_petsc_mat(os.path.join(REF_DATA, "matrix-A.dat"))
    b = read_petsc_vec(os.path.join(REF_DATA, "vector-b.dat"))
    assert shape == (4184, 4184) and cols.size == 49931
    np.savez_compressed(os.path.join(GOLD, "petsc_fixture_system.npz"), rowptr=rowptr, colidx=cols, vals=vals, b=b)
    print("all fixtures written to", os.path.abspath(GOLD))


if __name__ == "__main__":
    main()

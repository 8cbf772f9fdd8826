"""Pin the CPU oracle (oracle/petgem_oracle.py) against the reference's own outputs.

The fixtures were produced by oracle/make_golden.py, which imports the UNMODIFIED
reference in the build container; tests/test_mesh.py:31-36 known answers of the
reference are asserted too.  No GPU involved.
"""
import numpy as np
import pytest

from conftest import golden


def test_reference_mesh_known_answers(topo):
    # /root/reference/tests/test_mesh.py:31-36
    assert topo["elemsN"].shape[0] == 9453
    assert topo["nodes"].shape[0] == 2163
    assert topo["facesN"].shape[0] == 20039
    assert topo["edgesNodes"].shape[0] == 12748
    assert list(topo["elemsE"][0]) == [10591, 10600, 10831, 10832, 10601, 11465]
    assert list(topo["elemsF"][0]) == [17369, 17370, 17400, 17977]


def test_quadrature_facts():
    q = golden("quadrature_facts.npz")
    assert list(q["npts"]) == [4, 11, 24, 43, 126, 210]  # rule "2p" sizes (hvfem.py:251)
    assert np.allclose(q["wsum"], 1.0 / 6.0, rtol=1e-13)
    assert (q["exactness"] < 2e-13).all()  # exact for degree 2p -> any exact rule reproduces Me/Ke


def test_oracle_topology(oracle, topo):
    elemsN = topo["elemsN"].astype(np.int64)
    elemsE, edgesNodes = oracle.compute_edges(elemsN)
    elemsF, facesN = oracle.compute_faces(elemsN)
    assert np.array_equal(elemsE, topo["elemsE"])
    assert np.array_equal(edgesNodes, topo["edgesNodes"])
    assert np.array_equal(elemsF, topo["elemsF"])
    assert np.array_equal(facesN, topo["facesN"])
    facesE = oracle.compute_faces_edges(elemsF, elemsE, facesN.shape[0])
    assert np.array_equal(facesE, topo["facesE"])
    bfacesN, bFaces = oracle.compute_boundary_faces(elemsF, facesN)
    assert np.array_equal(bFaces, topo["bFaces"])
    bEdges = oracle.compute_boundary_edges(edgesNodes, bfacesN)
    assert np.array_equal(bEdges, topo["bEdges"])


@pytest.mark.parametrize("p", [1, 2, 3])
def test_oracle_dofs_and_boundaries(oracle, topo, p):
    elemsE, elemsF = topo["elemsE"].astype(np.int64), topo["elemsF"].astype(np.int64)
    dofs, dof_edges, dof_faces, _, total = oracle.compute_connectivity_dofs(elemsE, elemsF, p)
    assert total == int(topo["total_dofs_p%d" % p])
    assert np.array_equal(dofs[topo["dofs_sel"]], topo["dofs_rows_p%d" % p])
    assert np.array_equal(dofs.sum(axis=0), topo["dofs_sum_p%d" % p])
    bd = oracle.compute_boundaries(dof_edges, dof_faces, topo["bEdges"], topo["bFaces"])
    assert np.array_equal(bd, topo["boundary_dofs_p%d" % p])


def test_oracle_orientation(oracle, topo):
    elemsN, elemsE, elemsF = topo["elemsN"], topo["elemsE"], topo["elemsF"]
    edgesNodes, facesE = topo["edgesNodes"], topo["facesE"]
    for t in list(range(0, 9453, 97)) + [9452]:
        eo, fo = oracle.compute_element_orientation(elemsE[t], elemsN[t], edgesNodes[elemsE[t]], facesE[elemsF[t]])
        assert np.array_equal(np.concatenate([eo, fo]), topo["orient"][t])
    assert set(np.unique(topo["orient"][:, 6:])) == {0, 1, 2, 3, 4, 5}  # every face code occurs


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_oracle_shape_functions(oracle, p):
    g = golden("hvfem_shape.npz")
    pts, eos, fos, ref = g["pts_%d" % p], g["eo_%d" % p], g["fo_%d" % p], g["shape_%d" % p]
    for c in range(eos.shape[0]):
        for ip in range(pts.shape[0]):
            N, Cu = oracle.shape3d_etet(pts[ip], p, eos[c], fos[c])
            assert np.abs(N - ref[c, ip, 0]).max() <= 1e-14 * max(1.0, np.abs(ref[c, ip, 0]).max())
            assert np.abs(Cu - ref[c, ip, 1]).max() <= 1e-14 * max(1.0, np.abs(ref[c, ip, 1]).max())


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_oracle_element_matrices(oracle, p):
    g = golden("hvfem_elemental_p%d.npz" % p)
    worst = 0.0
    for i in range(g["coords"].shape[0]):
        J, Ji = oracle.compute_jacobian(g["coords"][i])
        Me, Ke = oracle.compute_elemental_matrices(g["eo"][i], g["fo"][i], J, Ji, p, g["sigma"][i])
        worst = max(worst, np.abs(Me - g["Me"][i]).max() / np.abs(g["Me"][i]).max(),
                    np.abs(Ke - g["Ke"][i]).max() / np.abs(g["Ke"][i]).max())
    assert worst < 1e-12, worst  # north_star: element values within 1e-12 (norm-relative)


def test_oracle_global_system_p1(oracle, topo):
    """Oracle assembly (scipy restatement of MatSetValues/ADD_VALUES) against the reference's
    element loop accumulated densely: y = A x and diag(A) on the 9453-tet test mesh."""
    g = golden("test_mesh_system_p1.npz")
    nodes, elemsN, elemsE, elemsF = topo["nodes"], topo["elemsN"], topo["elemsE"], topo["elemsF"]
    edgesNodes, facesE, tags = topo["edgesNodes"], topo["facesE"], topo["tags"]
    T = elemsN.shape[0]
    omega, mu, sig = float(g["omega"]), float(g["mu"]), g["sigma"]
    Ae = np.zeros((T, 6, 6), dtype=np.complex128)
    for t in range(T):
        s = sig[tags[t] - 1]
        Ae[t] = oracle.element_system(nodes[elemsN[t]], elemsN[t], elemsE[t], edgesNodes[elemsE[t]],
                                      facesE[elemsF[t]], np.array([s, s]), 1, omega, mu)
    assert np.abs(Ae[:64] - g["Ae_first"]).max() <= 1e-12 * np.abs(g["Ae_first"]).max()
    dofs, *_ = oracle.compute_connectivity_dofs(elemsE.astype(np.int64), elemsF.astype(np.int64), 1)
    N = 12748
    rowptr, colidx, vals = oracle.assemble_global(Ae, dofs, N)
    assert vals.size == 189700  # SURVEY 6 [probe]
    rp2, ci2 = oracle.csr_pattern(dofs, N)
    assert np.array_equal(rowptr, rp2) and np.array_equal(colidx, ci2)
    y = oracle.spmv(rowptr, colidx, vals, g["x"])
    assert np.abs(y - g["y"]).max() <= 1e-12 * np.abs(g["y"]).max()
    A = oracle.to_scipy(rowptr, colidx, vals)
    assert np.abs(A.diagonal() - g["diag"]).max() <= 1e-12 * np.abs(g["diag"]).max()
    assert abs(A - A.T).max() <= 1e-15 * np.abs(vals).max()  # complex symmetric


@pytest.mark.parametrize("p", [2, 3])
def test_oracle_test_mesh_elements(oracle, topo, p):
    g = golden("test_mesh_elements_p%d.npz" % p)
    nodes, elemsN, elemsE, elemsF = topo["nodes"], topo["elemsN"], topo["elemsE"], topo["elemsF"]
    edgesNodes, facesE, tags = topo["edgesNodes"], topo["facesE"], topo["tags"]
    omega, mu, sig = float(g["omega"]), float(g["mu"]), g["sigma"]
    sel = g["sel"][: (40 if p == 2 else 8)]
    for c, t in enumerate(sel):
        s = sig[tags[t] - 1]
        Ae = oracle.element_system(nodes[elemsN[t]], elemsN[t], elemsE[t], edgesNodes[elemsE[t]],
                                   facesE[elemsF[t]], np.array([s, s]), p, omega, mu)
        assert np.abs(Ae - g["Ae"][c]).max() <= 1e-12 * np.abs(g["Ae"][c]).max()


def test_oracle_krylov_on_petsc_fixture(oracle):
    """The reference's tests/test_petsc.py solves matrix-A.dat x = vector-b.dat and asserts
    nothing; we check the oracle's GMRES/BiCGStab against a direct solve (parity of the
    PETSc half is otherwise UNPINNED, see the oracle header)."""
    import scipy.sparse.linalg as spla

    g = golden("petsc_fixture_system.npz")
    A = oracle.to_scipy(g["rowptr"], g["colidx"], g["vals"])
    b = g["b"]
    xd = spla.spsolve(A.tocsc(), b)
    dinv = 1.0 / A.diagonal()
    # GMRES(30)+Jacobi stagnates on this non-symmetric fixture (1e-5 after 3000 its); restart 100 converges
    x, its, hist = oracle.gmres(lambda v: A @ v, b, rtol=1e-10, restart=100, maxit=4000, pc=lambda v: dinv * v)
    assert hist[-1] <= 1e-10 * hist[0] and its < 4000
    assert np.linalg.norm(x - xd) <= 1e-6 * np.linalg.norm(xd), (its, hist[-1])
    x2, its2, hist2 = oracle.bicgstab(lambda v: A @ v, b, rtol=1e-6, maxit=8000, pc=lambda v: dinv * v)
    assert np.linalg.norm(b - A @ x2) <= 1e-5 * np.linalg.norm(b), (its2, hist2[-1])


def test_oracle_cocg_cocr_on_assembled_system(oracle):
    """COCG / COCR restatements (the complex symmetric solvers of the GPU path) against a direct solve of
    a small system assembled by the oracle itself (shuffled Kuhn box, p = 2, Dirichlet applied)."""
    import scipy.sparse.linalg as spla

    from petgem_b200 import synthetic

    omega, mu, p = 2 * np.pi * 2.0, 4e-7 * np.pi, 2
    nodes, elemsN = synthetic.kuhn_box(3, length=700.0, seed=7)
    tab = synthetic.mesh_tables(nodes, elemsN)
    sigma = synthetic.layered_sigma(nodes, elemsN, vti_ratio=0.5)
    T = elemsN.shape[0]
    Ae = np.zeros((T, 20, 20), dtype=np.complex128)
    for t in range(T):
        Ae[t] = oracle.element_system(nodes[elemsN[t]], elemsN[t], tab["elemsE"][t],
                                      tab["edgesNodes"][tab["elemsE"][t]], tab["facesE"][tab["elemsF"][t]],
                                      sigma[t], p, omega, mu)
    dofs, dof_edges, dof_faces, _, N = oracle.compute_connectivity_dofs(tab["elemsE"], tab["elemsF"], p)
    rp, ci, v = oracle.assemble_global(Ae, dofs, N)
    bd = oracle.compute_boundaries(dof_edges, dof_faces, tab["bEdges"], tab["bFaces"])
    v = oracle.zero_rows_columns(rp, ci, v, bd, 1.0)
    A = oracle.to_scipy(rp, ci, v).tocsr()
    assert abs(A - A.T).max() <= 1e-15 * abs(A).max()  # complex symmetric, not Hermitian
    rng = np.random.default_rng(0)
    b = rng.normal(size=N) + 1j * rng.normal(size=N)
    b[bd] = 0.0
    xd = spla.spsolve(A.tocsc(), b)
    dinv = 1.0 / A.diagonal()
    xg, itg, hg = oracle.cocg(lambda y: A @ y, b, rtol=1e-11, maxit=5000, dinv=dinv)
    xr, itr, hr = oracle.cocr(lambda y: A @ y, b, rtol=1e-11, maxit=5000, dinv=dinv)
    assert itg < 5000 and itr < 5000
    assert np.linalg.norm(xg - xd) <= 1e-8 * np.linalg.norm(xd)
    assert np.linalg.norm(xr - xd) <= 1e-8 * np.linalg.norm(xd)
    assert itr <= itg  # the residual-minimising variant is never slower here

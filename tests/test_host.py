"""CPU tests of the host-side product code (no GPU, no oracle in the product path):
topology numbering, reference-element tables, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden
from petgem_b200 import basis, hvfem
from petgem_b200 import mesh as pmesh


def test_mesh_numbering_bit_exact(topo):
    elemsN = topo["elemsN"].astype(np.int64)
    T = elemsN.shape[0]
    elemsE, edgesNodes = pmesh.computeEdges(elemsN, T)
    elemsF, facesN = pmesh.computeFaces(elemsN, T)
    assert np.array_equal(elemsE, topo["elemsE"]) and np.array_equal(edgesNodes, topo["edgesNodes"])
    assert np.array_equal(elemsF, topo["elemsF"]) and np.array_equal(facesN, topo["facesN"])
    facesE = pmesh.computeFacesEdges(elemsF, elemsE, facesN.shape[0], T)
    assert np.array_equal(facesE, topo["facesE"])
    bFacesN, bFaces, nb = pmesh.computeBoundaryFaces(elemsF, facesN)
    assert nb == 2266 and np.array_equal(bFaces, topo["bFaces"])
    bEdges = pmesh.computeBoundaryEdges(edgesNodes, bFacesN)
    assert bEdges.size == 3399 and np.array_equal(bEdges, topo["bEdges"])
    for p in (1, 2, 3):
        dofs, de, df, dv, total = hvfem.computeConnectivityDOFS(elemsE, elemsF, p)
        assert total == int(topo["total_dofs_p%d" % p])
        assert np.array_equal(dofs[topo["dofs_sel"]], topo["dofs_rows_p%d" % p])
        assert np.array_equal(dofs.sum(axis=0), topo["dofs_sum_p%d" % p])
        _, bd = pmesh.computeBoundaries(dofs, de, df, bEdges, bFaces, p)
        assert np.array_equal(bd, topo["boundary_dofs_p%d" % p])


def test_mesh_rank_rows_fallback_matches_packed_keys():
    rng = np.random.default_rng(3)
    rows = np.sort(rng.integers(0, 50, size=(400, 3)), axis=1)
    u1, inv1, f1 = pmesh._rank_rows(rows)
    u2, f2, inv2 = np.unique(rows, axis=0, return_index=True, return_inverse=True)
    assert np.array_equal(u1, u2) and np.array_equal(inv1, inv2.reshape(-1)) and np.array_equal(f1, f2)
    big = rows.astype(np.int64) + (1 << 40)  # forces the lexicographic path
    u3, inv3, _ = pmesh._rank_rows(big)
    assert np.array_equal(u3 - (1 << 40), u2) and np.array_equal(inv3, inv1)


def test_orientation_host_matches_reference(topo):
    for t in range(0, 9453, 211):
        eo, fo = hvfem.computeElementOrientation(topo["elemsE"][t], topo["elemsN"][t],
                                                 topo["edgesNodes"][topo["elemsE"][t]], topo["facesE"][topo["elemsF"][t]])
        assert np.array_equal(np.concatenate([eo, fo]), topo["orient"][t])
        code = hvfem.pack_orientation(eo, fo)
        eo2, fo2 = hvfem.unpack_orientation(code)
        assert np.array_equal(eo, eo2) and np.array_equal(fo, fo2)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_basis_matches_reference_shape_functions(p):
    g = golden("hvfem_shape.npz")
    pts, eos, fos, ref = g["pts_%d" % p], g["eo_%d" % p], g["fo_%d" % p], g["shape_%d" % p]
    N, C = basis.evaluate_expanded(p, pts)
    for c in range(eos.shape[0]):
        J, S = basis.local_to_expanded(p, eos[c], fos[c])
        mine_n = np.moveaxis(N[J] * S[:, None, None], 0, -1)  # [pts, 3, n]
        mine_c = np.moveaxis(C[J] * S[:, None, None], 0, -1)
        assert np.abs(mine_n - ref[c, :, 0]).max() <= 1e-14 * max(1.0, np.abs(ref[c, :, 0]).max())
        assert np.abs(mine_c - ref[c, :, 1]).max() <= 1e-14 * max(1.0, np.abs(ref[c, :, 1]).max())
        n_, Sh, Cu = hvfem.shape3DETet(pts[0], np.ones(11, dtype=int) * p, eos[c], fos[c])
        assert n_ == basis.ndof_element(p) and np.allclose(Sh, ref[c, 0, 0], atol=1e-13)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_tables_reproduce_reference_element_matrices(p):
    """Host check of the contraction the kernels implement: Me = sum_c gM[c] SM[c][J,J] s s^T."""
    g = golden("hvfem_elemental_p%d.npz" % p)
    SM, SK = basis.element_tables(p)
    worst = 0.0
    for i in range(g["coords"].shape[0]):
        X = g["coords"][i]
        Jm = X[1:] - X[0]
        gf = hvfem.geometric_factors(Jm, g["sigma"][i])
        J, S = basis.local_to_expanded(p, g["eo"][i], g["fo"][i])
        ss = np.outer(S, S)
        Ke = np.einsum("c,cjk->jk", gf[:6], SK[:, J][:, :, J]) * ss
        Me = np.einsum("c,cjk->jk", gf[6:], SM[:, J][:, :, J]) * ss
        worst = max(worst, np.abs(Me - g["Me"][i]).max() / np.abs(g["Me"][i]).max(),
                    np.abs(Ke - g["Ke"][i]).max() / np.abs(g["Ke"][i]).max())
    assert worst < 1e-12, worst


def test_tet_quadrature_exact():
    from math import factorial as f
    for deg in (3, 7, 13):
        pts, w = basis.tet_quadrature(deg)
        assert (w > 0).all() and abs(w.sum() - 1 / 6) < 1e-15
        for a, b, c in ((deg, 0, 0), (1, deg - 2, 1), (deg // 3, deg // 3, deg - 2 * (deg // 3))):
            exact = f(a) * f(b) * f(c) / f(a + b + c + 3)
            assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum() - exact) <= 1e-13 * exact


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every symbol include/petgem_b200.h declares."""
    from petgem_b200 import _lib

    path = _lib.build()
    handle = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "petgem_b200.h")).read()
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(handle, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    assert L.pg_version() >= 100
    for p, n in zip(range(1, 7), (6, 20, 45, 84, 140, 216)):
        assert L.pg_ndof_element(p) == n == basis.ndof_element(p)
        assert L.pg_nexp(p) == basis.expanded_layout(p)["nexp"]


def test_synthetic_mesh_covers_all_orientation_codes():
    from petgem_b200 import synthetic

    nodes, elemsN = synthetic.kuhn_box(4)
    assert elemsN.shape == (6 * 64, 4)
    X = nodes[elemsN]
    assert (np.linalg.det(X[:, 1:] - X[:, :1]) > 0).all()
    tab = synthetic.mesh_tables(nodes, elemsN)
    codes = set()
    for t in range(elemsN.shape[0]):
        eo, fo = hvfem.computeElementOrientation(tab["elemsE"][t], elemsN[t], tab["edgesNodes"][tab["elemsE"][t]],
                                                 tab["facesE"][tab["elemsF"][t]])
        codes |= set(fo.tolist())
    assert codes == {0, 1, 2, 3, 4, 5}
    sig = synthetic.layered_sigma(nodes, elemsN)
    assert set(np.unique(sig[:, 0])) <= {3.3333, 1.0, 0.01}


def test_petsc_options_parser():
    from petgem_b200.krylov import parse_petsc_options

    o = parse_petsc_options("# Solver options for PETSc\n-ksp_type gmres\n-pc_type sor\n-ksp_rtol 1e-8\n-ksp_monitor\n")
    assert o == {"ksp_type": "gmres", "pc_type": "sor", "ksp_rtol": "1e-8", "ksp_monitor": True}


def test_p2_reduced_face_functions():
    """At p=2 the two functions of a face under any of the 6 orientations are +-2 of the 3 base
    functions (orientation 0 family 0/1, orientation 1 family 1): the table the p=2 kernel keeps in
    shared memory (pg_assemble.cu kFaceLut2)."""
    lut = [0, 1, 1, 2, 2, 0, 6, 5, 4, 6, 5, 4]  # entry o*2+fam: base | neg<<2
    lay = basis.expanded_layout(2)
    pts = np.random.default_rng(1).dirichlet(np.ones(4), size=6)[:, :3]
    N, C = basis.evaluate_expanded(2, pts)
    X = lambda f, o, fam: lay["face_off"] + (f * 6 + o) * 2 + fam  # noqa: E731
    for f in range(4):
        base = [X(f, 0, 0), X(f, 0, 1), X(f, 1, 1)]
        for o in range(6):
            for fam in range(2):
                e = lut[o * 2 + fam]
                s = -1.0 if e & 4 else 1.0
                assert np.abs(N[X(f, o, fam)] - s * N[base[e & 3]]).max() < 1e-14
                assert np.abs(C[X(f, o, fam)] - s * C[base[e & 3]]).max() < 1e-14


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
def test_reference_tensors_are_small_integers(p):
    """SK*DK and SM*DM are exact integers: what lets the kernels hold the table as 16-bit (p<=2) or
    32-bit (p=3..5) codes; the rounding distance must be far below 1/2."""
    nM, nK, err = basis.integer_tables(p)
    assert err < 1e-5
    assert max(np.abs(nM).max(), np.abs(nK).max()) < (32768 if p <= 3 else 2 ** 31)
    if p == 2:
        assert np.abs(nM).max() == 252 and np.abs(nK).max() == 160


def test_batched_orientation_equals_reference_table(topo):
    """hvfem.computeElementOrientation_batch on the whole test mesh == the reference's per-element codes."""
    from petgem_b200 import hvfem

    eo, fo = hvfem.computeElementOrientation_batch(topo["elemsE"], topo["elemsN"], topo["edgesNodes"][topo["elemsE"]],
                                                   topo["facesE"][topo["elemsF"]])
    assert np.array_equal(eo, topo["orient"][:, :6]) and np.array_equal(fo, topo["orient"][:, 6:])
    assert len(set(map(int, fo.reshape(-1)))) == 6  # all six face codes occur on this mesh


def test_pc_substitution_is_announced_and_direct_solves_refused():
    """-pc_type values of the reference's option files that have no counterpart are replaced with a notice;
    direct factorisations raise (no silent rewriting)."""
    from petgem_b200 import krylov

    said = []
    assert krylov.resolve_pc({"ksp_type": "gmres", "pc_type": "sor"}, notify=said.append) == "hiptmair"
    assert len(said) == 1 and "sor" in said[0] and "hiptmair" in said[0]
    assert krylov.resolve_pc({"pc_type": "gamg"}, notify=said.append, have_mesh=False) == "jacobi"
    assert krylov.resolve_pc({"pc_type": "jacobi"}, notify=said.append) == "jacobi" and len(said) == 2
    for bad in ({"ksp_type": "preonly", "pc_type": "lu"}, {"pc_type": "cholesky"}, {"ksp_type": "preonly"}):
        with pytest.raises(krylov.UnsupportedSolverError):
            krylov.resolve_pc(bad, notify=said.append)


def test_common_star_import():
    """`from petgem.common import *` is the reference's import style (solver.py:16)."""
    ns = {}
    exec("from petgem_b200.common import *", ns)
    assert "Print" in ns and "InputParameters" in ns and "Timers" in ns


def test_h5lite_roundtrip_and_checksum(tmp_path):
    """The HDF5 subset of the results file (postprocessing.py:341-461): lookup3 known answers (Jenkins'
    lookup3.c self-test strings), and write -> read of every value kind the schema uses."""
    from petgem_b200 import h5lite

    s = b"Four score and seven years ago"
    assert h5lite.lookup3(s, 0) == 0x17770551 and h5lite.lookup3(s, 1) == 0xCD628161
    assert h5lite.lookup3(b"", 0) == 0xDEADBEEF
    tree = {"machine": {"machine": "node. 6.1. x86_64", "num_processors": 8, "petgem_version": "1.0"},
            "model": {"nord": 2, "cuda": True, "vtk": False, "run-time (s)": 1.25, "mode": "mt", "polarization": "xy",
                      "source_position (m)": np.array([1750.0, 1750.0, -975.0]),
                      "E-fields_mode_x": {"x": np.arange(7) * (1.5 - 2j), "y": np.zeros(7, dtype=complex)},
                      "apparent_resistivity": {"xx": np.linspace(0, 1, 7)},
                      "wide": {"k%02d" % i: i for i in range(24)}}}  # more links than libhdf5's compact default
    f = str(tmp_path / "r.h5")
    h5lite.write(f, tree)
    raw = open(f, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and len(raw) % 8 == 0
    r = h5lite.read(f)
    assert r["machine"] == tree["machine"]
    m = r["model"]
    assert m["cuda"] and not m["vtk"] and m["nord"] == 2 and m["run-time (s)"] == 1.25 and m["polarization"] == "xy"
    assert np.array_equal(m["source_position (m)"], tree["model"]["source_position (m)"])
    assert np.array_equal(m["E-fields_mode_x"]["x"], tree["model"]["E-fields_mode_x"]["x"])
    assert m["E-fields_mode_x"]["x"].dtype == np.complex128
    assert [m["wide"]["k%02d" % i] for i in range(24)] == list(range(24))
    # the same tree in the classic layout (what Postprocessing writes; the reader is the one that parses the
    # reference's h5py-written files)
    g = str(tmp_path / "c.h5")
    h5lite.write_classic(g, tree)
    rawc = open(g, "rb").read()
    assert rawc[:8] == raw[:8] and rawc[8] == 0 and len(rawc) == int.from_bytes(rawc[40:48], "little")
    c = h5lite.read_classic(g)
    assert c["machine"] == tree["machine"]
    mc = c["model"]
    assert mc["cuda"] and not mc["vtk"] and mc["nord"] == 2 and mc["run-time (s)"] == 1.25 and mc["polarization"] == "xy"
    assert np.array_equal(mc["E-fields_mode_x"]["x"], tree["model"]["E-fields_mode_x"]["x"])
    assert mc["E-fields_mode_x"]["y"].dtype == np.complex128 and mc["apparent_resistivity"]["xx"].dtype == np.float64
    assert [mc["wide"]["k%02d" % i] for i in range(24)] == list(range(24))
    # a flipped byte in an object header is caught by its checksum
    bad = bytearray(raw)
    bad[raw.index(b"OHDR") + 12] ^= 0x40
    open(f, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        h5lite.read(f)


def test_h5lite_reads_the_reference_receiver_file(tmp_path):
    """h5lite.read_classic on tests/golden/receiver_pos_reference.h5 -- the reference's own tests/data/receiver_pos.h5,
    written by h5py (classic layout: version-0 superblock, symbol-table group, contiguous float64 dataset) -- and
    preprocessing.read_receivers on top of it (preprocessing.py:399-407) without h5py."""
    from petgem_b200 import h5lite
    from petgem_b200.preprocessing import read_receivers

    f = os.path.join(GOLDEN, "receiver_pos_reference.h5")
    r = h5lite.read_classic(f)
    rec = golden("case1_receivers.npy")
    assert set(r) == {"data"} and r["data"].dtype == np.float64 and np.array_equal(r["data"], rec)
    assert np.array_equal(read_receivers(f), rec)
    with pytest.raises(SystemExit):  # a [58, 3] table is not a [T, 2] conductivity model
        read_receivers(f, cols=2, rows=9453)
    # a file written by h5lite.write is read through the same entry point
    g = str(tmp_path / "w.h5")
    h5lite.write(g, {"data": rec})
    assert np.array_equal(h5lite.read_classic(g)["data"], rec) and np.array_equal(read_receivers(g), rec)


def test_peer_exchange_offsets_without_a_gpu():
    """PeerExchange (peer.py): where each rank's segment starts inside the receive area of its destinations, the
    segment table and the reader mask, from the send/recv counts alone (no CUDA: a stand-in communicator)."""
    torch = pytest.importorskip("torch")
    from petgem_b200.peer import PeerExchange

    # recv[d][s] = entries rank d receives from rank s (3 ranks); send[s][d] = recv[d][s]
    recv = [[0, 4, 2], [3, 0, 0], [5, 1, 0]]

    class Comm:
        def __init__(self, rank):
            self.rank, self.world, self._c = rank, 3, 0

        def channel(self):
            self._c += 1
            return self._c - 1

        def all_gather_object(self, obj):
            return recv

    for s in range(3):
        send_splits = [recv[d][s] for d in range(3)]
        ex = PeerExchange(Comm(s), torch.arange(sum(send_splits)), send_splits, recv[s])
        assert list(ex.seg) == [0] + list(np.cumsum(send_splits))
        assert ex.from_mask == sum(1 << r for r in range(3) if recv[s][r] > 0) and ex.n_recv == sum(recv[s])
        # my segment in rank d's receive area starts after the segments of the lower ranks
        assert ex.dst_entry == [sum(recv[d][:s]) for d in range(3)]

        class Buf:
            addr = [1000, 2000, 3000]
        table = ex.target(Buf, [64, 128, 256], k=2)
        for d in range(3):
            want = (Buf.addr[d] + [64, 128, 256][d] + ex.dst_entry[d] * 2 * 16) if send_splits[d] else None
            assert table[d] == want


def test_results_tree_has_the_reference_schema():
    """postprocessing.results_tree: the groups / datasets of the reference's results file
    (postprocessing.py:346-458) for a CSEM and an MT run, from stand-in input objects (no GPU)."""
    from petgem_b200 import postprocessing as post

    class Setup:
        pass

    rec = 7
    f = (np.arange(rec * 6).reshape(rec, 6) * (1 + 0.5j)).astype(np.complex128)
    s = Setup()
    s.model = {"mode": "csem", "mesh": "m.msh", "receivers": "r.h5",
               "csem": {"sigma": {"horizontal": [1.0, 0.01], "vertical": [1.0, 0.01]},
                        "source": {"frequency": 2.0, "position": [1750.0, 1750.0, -975.0], "azimuth": 0.0, "dip": 0.0,
                                   "current": 1.0, "length": 1.0}}}
    s.run = {"nord": 2, "cuda": True, "num_polarizations": 1, "conductivity_from_file": False}
    s.output = {"vtk": False}
    t = post.results_tree(s, {"fields_0": f, "run_time_s": 1.5}, total_num_dofs=65574, solver_type="cr")
    assert set(t) == {"machine", "model"} and set(t["machine"]) == {"machine", "num_processors", "petgem_version"}
    m = t["model"]
    for key in ("date", "mesh_file", "receivers_file", "nord", "dof", "cuda", "vtk", "mode", "num-polarizations", "solver",
                "run-time (s)", "sigma_horizontal (S/m)", "sigma_vertical (S/m)", "frequency (Hz)", "source_position (m)",
                "source_azimuth (deg)", "source_dip (deg)", "source_current (Am)", "source_length (m)", "E-fields",
                "H-fields"):
        assert key in m, key
    assert m["dof"] == 65574 and m["solver"] == "cr" and np.array_equal(m["E-fields"]["y"], f[:, 1])
    assert np.array_equal(m["H-fields"]["z"], f[:, 5])
    # MT: fields per polarization + impedance, apparent resistivity, phase, tipper
    s.model = {"mode": "mt", "mesh": "m.msh", "receivers": "r.h5",
               "mt": {"sigma": {"horizontal": [1e-10, 0.01], "vertical": [1e-10, 0.01]}, "frequency": 2.0,
                      "polarization": "xy"}}
    s.run["num_polarizations"] = 2
    four = np.arange(4 * rec).reshape(4, rec).astype(np.complex128)
    out = {"fields_0": f, "fields_1": 2 * f, "impedance": four, "apparent_resistivity": four.real, "phase": four.real,
           "tipper": four[:2], "run_time_s": 2.0}
    m = post.results_tree(s, out, total_num_dofs=10)["model"]
    assert {"E-fields_mode_x", "H-fields_mode_x", "E-fields_mode_y", "H-fields_mode_y", "impedance",
            "apparent_resistivity", "phase", "tipper", "polarization", "frequency (Hz)"} <= set(m)
    assert set(m["impedance"]) == {"xx", "xy", "yx", "yy"} and np.array_equal(m["impedance"]["yx"], four[2])
    assert np.array_equal(m["E-fields_mode_y"]["x"], 2 * f[:, 0]) and set(m["tipper"]) == {"x", "y"}

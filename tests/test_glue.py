"""Drop-in surface around the hot path: params file, PETSc-binary scratch files, preprocessing
(CPU), and the kernel.py command line end to end (GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden


def write_msh22(path, nodes, elemsN, tags):
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % nodes.shape[0])
        for i, (x, y, z) in enumerate(nodes):
            fh.write("%d %.17g %.17g %.17g\n" % (i + 1, x, y, z))
        fh.write("$EndNodes\n$Elements\n%d\n" % elemsN.shape[0])
        for i, (t, tag) in enumerate(zip(elemsN, tags)):
            fh.write("%d 4 2 %d %d %d %d %d %d\n" % (i + 1, tag, tag, t[0] + 1, t[1] + 1, t[2] + 1, t[3] + 1))
        fh.write("$EndElements\n")


PARAMS = """
model:
  mode: csem
  csem:
    sigma:
      horizontal: [1., 0.01, 1., 3.3333]
      vertical: [1., 0.01, 1., 3.3333]
    source:
      frequency: 2.
      position: [1750., 1750., -975.]
      azimuth: 0.
      dip: 0.
      current: 1.
      length: 1.
  mesh: %(mesh)s
  receivers: %(rec)s
run:
  nord: %(nord)d
  cuda: True
output:
  vtk: False
  directory: %(out)s
  directory_scratch: %(tmp)s
  remove_scratch: False
"""


def make_case(tmp_path, topo, nord=1):
    mesh = str(tmp_path / "case.msh")
    write_msh22(mesh, topo["nodes"], topo["elemsN"], topo["tags"])
    rec = str(tmp_path / "receivers.npy")
    np.save(rec, golden("case1_receivers.npy"))
    params = str(tmp_path / "params.yaml")
    with open(params, "w") as fh:
        fh.write(PARAMS % dict(mesh=mesh, rec=rec, nord=nord, out=str(tmp_path / "out"), tmp=str(tmp_path / "tmp")))
    opts = str(tmp_path / "petsc.opts")
    with open(opts, "w") as fh:
        fh.write("# Solver options for PETSc\n-ksp_type gmres\n-pc_type sor\n-ksp_rtol 1e-11\n-ksp_max_it 30000\n")
    return params, opts


def test_input_parameters_schema(tmp_path, topo):
    from petgem_b200.common import InputParameters

    params, _ = make_case(tmp_path, topo)
    s = InputParameters(params)
    assert s.model["mode"] == "csem" and s.run["nord"] == 1 and s.run["cuda"] is True
    assert s.run["num_polarizations"] == 1 and s.run["conductivity_from_file"] is False
    assert os.path.isdir(s.output["directory"]) and os.path.isdir(s.output["directory_scratch"])
    bad = str(tmp_path / "bad.yaml")
    open(bad, "w").write(open(params).read().replace("nord: 1", "nord: 9"))
    with pytest.raises(SystemExit):
        InputParameters(bad)


def test_petsc_binary_roundtrip_and_reference_fixture(tmp_path):
    from petgem_b200 import parallel as par

    rng = np.random.default_rng(0)
    a = rng.normal(size=(7, 5))
    f = str(tmp_path / "m.dat")
    par.writeParallelDenseMatrix(f, par.createSequentialDenseMatrixWithArray(7, 5, a))
    m = par.readPetscMatrix(f)
    assert m.getSize() == (7, 5) and np.array_equal(m.array.real, a) and not m.array.imag.any()
    cols, row = m.getRow(3)
    assert np.array_equal(row.real, a[3]) and np.array_equal(cols, np.arange(5))
    raw = open(f, "rb").read()
    assert np.frombuffer(raw[:16], dtype=">i4").tolist() == [1211216, 7, 5, 35]  # PETSc MAT_FILE_CLASSID layout
    v = rng.normal(size=11) + 1j * rng.normal(size=11)
    g = str(tmp_path / "v.dat")
    par.writePetscVector(g, par.createSequentialVectorWithArray(v))
    assert np.array_equal(par.readPetscVector(g).getArray(), v)
    assert np.frombuffer(open(g, "rb").read()[:8], dtype=">i4").tolist() == [1211214, 11]
    # same bytes as the reference's own fixture reader would expect: write the golden system and re-read it
    gold = golden("petsc_fixture_system.npz")
    par.writePetscVector(g, par.createSequentialVectorWithArray(gold["b"]))
    assert np.array_equal(par.readPetscVector(g).getArray(), gold["b"])


def test_preprocessing_writes_reference_scratch_files(tmp_path, topo):
    from petgem_b200 import parallel as par
    from petgem_b200.common import InputParameters
    from petgem_b200.preprocessing import Preprocessing

    params, _ = make_case(tmp_path, topo, nord=2)
    setup = InputParameters(params)
    Preprocessing().run(setup)
    tmp = setup.output["directory_scratch"]
    T = 9453
    assert np.array_equal(par.readPetscMatrix(tmp + "/edges.dat").array.real.astype(np.int64), topo["elemsE"])
    assert np.array_equal(par.readPetscMatrix(tmp + "/faces.dat").array.real.astype(np.int64), topo["elemsF"])
    fe = par.readPetscMatrix(tmp + "/facesEdges.dat").array.real.astype(np.int64)
    assert np.array_equal(fe, topo["facesE"][topo["elemsF"]].reshape(T, 12))
    dofs = par.readPetscMatrix(tmp + "/dofs.dat").array.real.astype(np.int64)
    assert dofs.shape == (T, 20) and np.array_equal(dofs[topo["dofs_sel"]], topo["dofs_rows_p2"])
    bd = par.readPetscVector(tmp + "/boundaries.dat").getArray().real.astype(np.int64)
    assert np.array_equal(bd, topo["boundary_dofs_p2"])
    src = par.readPetscVector(tmp + "/source.dat").getArray().real
    assert src.size == 70 and np.array_equal(src[:4].astype(np.int64), topo["elemsN"][8946])
    nnz = par.readPetscVector(tmp + "/nnz.dat").getArray().real
    assert nnz.size == int(topo["total_dofs_p2"]) and (nnz == 200).all()
    sig = par.readPetscMatrix(tmp + "/conductivityModel.dat").array.real
    assert np.array_equal(sig[:, 0], np.array([1.0, 0.01, 1.0, 3.3333])[topo["tags"] - 1])


def test_cpu_matrix_type_fails_loudly():
    from petgem_b200 import parallel as par

    with pytest.raises(SystemExit):
        par.createParallelMatrix(10, 10, None, False)
    with pytest.raises(SystemExit):
        par.createParallelVector(10, False)


@pytest.mark.gpu
def test_kernel_cli_two_ranks_match_one(tmp_path, topo):
    """torchrun -n 2 kernel.py (row blocks over 2 GPUs, NCCL halo + allreduce) gives the same receiver
    fields as the single-GPU run (p=2)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fields = []
    for world in (1, 2):
        d = tmp_path / ("w%d" % world)
        d.mkdir()
        params, opts = make_case(d, topo, nord=2)
        # -pc_type sor (the reference's shipped choice) -> Hiptmair preconditioner, distributed over the ranks
        open(opts, "w").write("-ksp_type cr\n-pc_type sor\n-ksp_rtol 1e-11\n-ksp_max_it 40000\n")
        cmd = [sys.executable, os.path.join(ROOT, "kernel.py"), "-options_file", opts, params]
        if world > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                   "--master-addr", "127.0.0.1", "--master-port", "29533"] + cmd[1:]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
        fields.append(np.load(str(d / "out" / "fields.npz"))["fields_0"])
    scale = np.abs(fields[0][:, :3]).max()
    assert np.abs(fields[0][:, :3] - fields[1][:, :3]).max() <= 1e-6 * scale
    gold = golden("test_mesh_fields_p2.npz")["fields"]
    for f in fields:
        assert np.abs(f[:, :3] - gold[:, :3]).max() <= 1e-6 * np.abs(gold[:, :3]).max()


@pytest.mark.gpu
def test_kernel_cli_p2_with_the_shipped_solver_options(tmp_path, topo):
    """kernel.py at p = 2 with the solver options PETGEM ships (examples/case1/petsc.opts: gmres + sor, tight
    rtol here): the SOR request is announced as replaced by the Hiptmair preconditioner and the receiver
    fields match the golden fields of the reference pipeline (oracle/make_golden_fields.py) to 1e-6."""
    params, opts = make_case(tmp_path, topo, nord=2)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "kernel.py"), "-options_file", opts, params],
                         capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "-pc_type sor is not available" in res.stdout and "hiptmair" in res.stdout
    F = np.load(str(tmp_path / "out" / "fields.npz"))["fields_0"]
    gold = golden("test_mesh_fields_p2.npz")["fields"]
    assert np.abs(F[:, :3] - gold[:, :3]).max() <= 1e-6 * np.abs(gold[:, :3]).max()
    assert np.abs(F[:, 3:] - gold[:, 3:]).max() <= 1e-6 * np.abs(gold[:, 3:]).max()


@pytest.mark.gpu
def test_kernel_cli_refuses_direct_solver_options(tmp_path, topo):
    """-ksp_type preonly -pc_type lu (examples/case4/petsc.opts, MUMPS) has no counterpart: fail loudly, do
    not silently iterate."""
    params, opts = make_case(tmp_path, topo, nord=1)
    open(opts, "w").write("-ksp_type preonly\n-pc_type lu\n-pc_factor_mat_solver_type mumps\n")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "kernel.py"), "-options_file", opts, params],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode != 0
    assert "direct factorisation" in res.stdout
    assert not os.path.exists(str(tmp_path / "tmp" / "x0.dat"))


@pytest.mark.gpu
def test_kernel_cli_end_to_end(tmp_path, topo, oracle):
    """python3 kernel.py -options_file petsc.opts params.yaml on the reference's test mesh (case1
    physics, p=1): receiver E-fields within 1e-6 of a direct solve of the oracle's system."""
    import scipy.sparse.linalg as spla

    params, opts = make_case(tmp_path, topo, nord=1)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "kernel.py"), "-options_file", opts, params],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    out = np.load(str(tmp_path / "out" / "fields.npz"))
    E = out["fields_0"]
    # oracle: assemble with the reference algorithm restated on the CPU, direct solve
    p, omega, mu = 1, 2 * np.pi * 2.0, 4e-7 * np.pi
    nodes, elemsN, elemsE, elemsF = topo["nodes"], topo["elemsN"], topo["elemsE"], topo["elemsF"]
    edgesNodes, facesE, tags = topo["edgesNodes"], topo["facesE"], topo["tags"]
    T = elemsN.shape[0]
    sig = np.array([1.0, 0.01, 1.0, 3.3333])
    Ae = np.zeros((T, 6, 6), dtype=np.complex128)
    for t in range(T):
        s = sig[tags[t] - 1]
        Ae[t] = oracle.element_system(nodes[elemsN[t]], elemsN[t], elemsE[t], edgesNodes[elemsE[t]],
                                      facesE[elemsF[t]], np.array([s, s]), p, omega, mu)
    dofs, *_ , N = oracle.compute_connectivity_dofs(elemsE.astype(np.int64), elemsF.astype(np.int64), p)
    rp, ci, v = oracle.assemble_global(Ae, dofs, N)
    bd = topo["boundary_dofs_p1"]
    v = oracle.zero_rows_columns(rp, ci, v, bd, 1.0)
    src = np.array([1750.0, 1750.0, -975.0])
    t = 8946
    b = oracle.csem_rhs(N, nodes[elemsN[t]], elemsN[t], elemsE[t], edgesNodes[elemsE[t]], facesE[elemsF[t]], dofs[t],
                        p, src, 0.0, 0.0, 1.0, 1.0, omega, mu)
    b[bd] = 0.0
    xd = spla.spsolve(oracle.to_scipy(rp, ci, v).tocsc(), b)
    rec = golden("case1_receivers.npy")
    Ed = oracle.field_interpolator(xd, nodes, elemsN, elemsE, edgesNodes, elemsF, facesE, dofs, rec, p, omega, mu)
    assert E.shape == Ed.shape
    assert np.abs(E[:, :3] - Ed[:, :3]).max() <= 1e-6 * np.abs(Ed[:, :3]).max()
    assert np.abs(E[:, 3:] - Ed[:, 3:]).max() <= 1e-6 * np.abs(Ed[:, 3:]).max()
    # the reference's results file (postprocessing.py:341-461): <mode>_petgemV<version>_<date>.h5 with the
    # groups machine / model and the receiver responses per component
    import glob

    from petgem_b200 import h5lite
    h5 = glob.glob(str(tmp_path / "out" / "csem_petgemV*.h5"))
    assert len(h5) == 1
    r = h5lite.read_classic(h5[0])
    assert open(h5[0], "rb").read()[8] == 0  # version-0 superblock: the layout h5py writes
    assert set(r) == {"machine", "model"} and r["machine"]["petgem_version"] == "1.0"
    m = r["model"]
    assert m["mode"] == "csem" and m["nord"] == 1 and m["dof"] == N and bool(m["cuda"])
    assert m["solver"] == "gmres" and m["frequency (Hz)"] == 2.0 and m["num-polarizations"] == 1
    assert np.array_equal(m["source_position (m)"], src)
    for i, c in enumerate("xyz"):
        assert np.array_equal(m["E-fields"][c], E[:, i]) and np.array_equal(m["H-fields"][c], E[:, 3 + i])
    # x0.dat is a PETSc binary Vec the reference Postprocessing could read (solver.py:593-594)
    from petgem_b200.parallel import readPetscVector
    x = readPetscVector(str(tmp_path / "tmp" / "x0.dat")).getArray()
    assert np.linalg.norm(x - xd) <= 1e-6 * np.linalg.norm(xd)


PARAMS_MT = """
model:
  mode: mt
  mt:
    sigma:
      horizontal: [1., 0.01, 1., 3.3333]
      vertical: [1., 0.01, 1., 3.3333]
    frequency: 2.
    polarization: 'xy'
  mesh: %(mesh)s
  receivers: %(rec)s
run:
  nord: %(nord)d
  cuda: True
output:
  vtk: False
  directory: %(out)s
  directory_scratch: %(tmp)s
  remove_scratch: False
"""


def make_mt_case(tmp_path, topo, nord=1):
    params, opts = make_case(tmp_path, topo, nord=nord)
    with open(params, "w") as fh:
        fh.write(PARAMS_MT % dict(mesh=str(tmp_path / "case.msh"), rec=str(tmp_path / "receivers.npy"), nord=nord,
                                  out=str(tmp_path / "out"), tmp=str(tmp_path / "tmp")))
    with open(opts, "w") as fh:
        fh.write("-ksp_type cr\n-pc_type jacobi\n-ksp_rtol 1e-9\n-ksp_max_it 40000\n")
    return params, opts


def test_mt_preprocessing_writes_boundary_elements(tmp_path, topo):
    """mode: mt -> boundaryElements.dat with the reference's row layout (preprocessing.py:326-381)."""
    from petgem_b200 import parallel as par
    from petgem_b200.common import InputParameters
    from petgem_b200.preprocessing import Preprocessing

    params, _ = make_mt_case(tmp_path, topo, nord=2)
    setup = InputParameters(params)
    assert setup.run["num_polarizations"] == 2
    Preprocessing().run(setup)
    tmp = setup.output["directory_scratch"]
    rows = par.readPetscMatrix(tmp + "/boundaryElements.dat").array.real
    gold = golden("mt_rhs.npz")
    nb = gold["bElems"].size
    assert rows.shape == (nb, 53 + 20)
    t = gold["bElems"].astype(np.int64)
    assert np.array_equal(rows[:, 0:4].astype(np.int64), topo["elemsN"][t])
    assert np.array_equal(rows[:, 4:16], topo["nodes"][topo["elemsN"][t]].reshape(nb, 12))
    assert np.array_equal(rows[:, 16:20].astype(np.int64), topo["elemsF"][t])
    assert np.array_equal(rows[:, 50].astype(np.int64), gold["planeFace"])
    assert np.array_equal(rows[:, 51].astype(np.int64), topo["bFaces"])
    assert np.array_equal(rows[:, 52], gold["sigma_by_tag"][topo["tags"][t] - 1])
    assert not os.path.exists(tmp + "/boundaries.dat")  # MT has natural boundary conditions only


@pytest.mark.gpu
@pytest.mark.parametrize("nord", [1, 2])
def test_kernel_cli_mt_two_polarizations(tmp_path, topo, nord):
    """kernel.py in MT mode (p=1, 2): both polarizations assembled from the 1-D excitation, solved in
    lockstep, and each x{i}.dat solves its system: ||b_i - A x_i|| <= 1e-7 ||b_i|| with A, b from the
    oracle-checked pieces (fused assembly without Dirichlet rows, mt_rhs)."""
    import torch

    from petgem_b200.parallel import readPetscMatrix, readPetscVector

    params, opts = make_mt_case(tmp_path, topo, nord=nord)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "kernel.py"), "-options_file", opts, params],
                         capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    from petgem_b200 import mt
    from petgem_b200.device import AssemblyPlan, CSRMatrix, ElementData

    tmp = str(tmp_path / "tmp")
    rows = readPetscMatrix(tmp + "/boundaryElements.dat").array.real
    omega, mu = 2 * np.pi * 2.0, 4e-7 * np.pi
    elem_z = topo["nodes"][topo["elemsN"]][:, :, 2]
    N = int(topo["total_dofs_p%d" % nord])
    bs = mt.mt_rhs(rows, elem_z.max(), elem_z.min(), nord, omega, mu, ["x", "y"], N)
    sig = np.array([1.0, 0.01, 1.0, 3.3333])[topo["tags"] - 1]
    el = ElementData.from_mesh(topo["nodes"], topo["elemsN"], topo["elemsE"], topo["edgesNodes"], topo["elemsF"],
                               topo["facesE"], np.stack([sig, sig], axis=1))
    plan = AssemblyPlan(el, nord, order="reference")
    geo, code = el.geometry()
    A = CSRMatrix(*plan.csr(), plan.assemble(geo, code, omega, mu), plan.N)
    for i in range(2):
        x = readPetscVector(tmp + "/x%d.dat" % i).getArray()
        b = torch.as_tensor(bs[i], device=A.vals.device)
        r = b - A.mult(torch.as_tensor(x, device=A.vals.device))
        assert float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(b)) <= 1e-7
    assert os.path.exists(str(tmp_path / "out" / "fields.npz"))
    # the MT results file (postprocessing.py:424-458): fields per polarization, impedance, apparent resistivity,
    # phase and tipper
    import glob

    from petgem_b200 import h5lite
    h5 = glob.glob(str(tmp_path / "out" / "mt_petgemV*.h5"))
    assert len(h5) == 1
    m = h5lite.read_classic(h5[0])["model"]
    F = np.load(str(tmp_path / "out" / "fields.npz"))
    assert m["mode"] == "mt" and m["num-polarizations"] == 2 and m["nord"] == nord
    assert set(m["impedance"]) == {"xx", "xy", "yx", "yy"} and set(m["tipper"]) == {"x", "y"}
    pol = [k for k in m if k.startswith("E-fields_mode_")]
    assert len(pol) == 2
    assert np.array_equal(m["impedance"]["xy"], F["impedance"][1])
    assert np.array_equal(m["apparent_resistivity"]["yx"], F["apparent_resistivity"][2])

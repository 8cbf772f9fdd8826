"""A C program (tests/c_abi/solve_case.c, no Python, no torch) drives the whole hot path through the C ABI of
include/petgem_b200.h on the reference's test mesh with the case1 physics: pg_tables_init ->
pg_element_geometry -> pg_plan_create/set_dirichlet/csr -> pg_assemble -> pg_csem_rhs -> pg_krylov_solve
(COCR, BiCGStab, GMRES) -> pg_interpolate_fields.  Its receiver fields must match the golden fields of the
reference pipeline (oracle/make_golden_fields.py) to 1e-6 (north_star; postprocessing.py:566-614)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
SRC = os.path.join(ROOT, "tests", "c_abi", "solve_case.c")


def _compile(out):
    lib_dir = os.path.join(ROOT, "petgem_b200")
    cmd = ["gcc", "-O2", "-std=gnu99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           SRC, "-o", out, "-L", lib_dir, "-lpetgem_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_c_caller_compiles_against_the_header(tmp_path):
    """CPU: the C program builds with gcc against include/petgem_b200.h and links the library (no run)."""
    from petgem_b200 import _lib

    _lib.build()
    _compile(str(tmp_path / "solve_case"))


@pytest.mark.gpu
@pytest.mark.parametrize("method,name", [(1, "cocr"), (2, "bcgs"), (3, "gmres")])
def test_c_caller_solves_case1(tmp_path, topo, method, name):
    p = 2
    exe = _compile(str(tmp_path / "solve_case"))
    d = tmp_path / "case"
    d.mkdir()
    T = topo["elemsN"].shape[0]
    nE, nF = topo["edgesNodes"].shape[0], topo["facesE"].shape[0]
    sig = np.array([1.0, 0.01, 1.0, 3.3333])[topo["tags"] - 1]
    rec = golden("case1_receivers.npy")
    bd = np.zeros(nE + nF, dtype=np.uint8)
    bd[topo["bEdges"]] = 1
    bd[nE + topo["bFaces"]] = 1
    arrays = {
        "nodes": topo["nodes"][topo["elemsN"]].reshape(T, 12).astype(np.float64),
        "sigma": np.stack([sig, sig], axis=1).astype(np.float64),
        "elemsN": topo["elemsN"].astype(np.int32), "elemsE": topo["elemsE"].astype(np.int32),
        "edgesNodes": topo["edgesNodes"][topo["elemsE"]].reshape(T, 12).astype(np.int32),
        "facesEdges": topo["facesE"][topo["elemsF"]].reshape(T, 12).astype(np.int32),
        "elemsF": topo["elemsF"].astype(np.int32), "bd_entity": bd, "receivers": rec.astype(np.float64),
    }
    for k, v in arrays.items():
        np.ascontiguousarray(v).tofile(str(d / (k + ".bin")))
    omega, mu = 2 * np.pi * 2.0, 4e-7 * np.pi
    rtol = 1e-12 if method == 1 else 1e-11
    with open(d / "manifest.txt", "w") as fh:  # examples/case1 params.yaml:13-18: x-directed unit dipole
        fh.write("%d %d %d %d %d %d %d %.17g %.17g 1750.0 1750.0 -975.0 1.0 0.0 0.0 %g\n"
                 % (T, topo["nodes"].shape[0], nE, nF, p, rec.shape[0], int(bd.sum()), omega, mu, rtol))
    res = subprocess.run([exe, str(d), str(method)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    its, rel, N, nnz, true_res = res.stdout.split()
    assert int(N) == 65574 and float(rel) <= rtol and float(true_res) <= 1e-7, res.stdout
    F = np.fromfile(str(d / "fields.bin"), dtype=np.complex128).reshape(-1, 6)
    gold = golden("test_mesh_fields_p2.npz")["fields"]
    assert np.abs(F[:, :3] - gold[:, :3]).max() <= 1e-6 * np.abs(gold[:, :3]).max(), (name, its)
    assert np.abs(F[:, 3:] - gold[:, 3:]).max() <= 1e-6 * np.abs(gold[:, 3:]).max(), (name, its)

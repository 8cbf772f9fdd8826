"""Mat / Vec objects and PETSc-binary I/O: the plug point of the B200 backend.

Mirrors ``petgem/parallel.py``: same function names and argument order.  The
reference switches PETSc object types with ``run.cuda`` (``parallel.py:164-167``,
``:192-195``: ``aijcusparse`` / ``cuda``); here that switch selects :class:`B200Mat`
and :class:`B200Vec`, device-resident objects whose arithmetic is the C ABI of
include/petgem_b200.h.  There is no CPU matrix type: ``matrix_type=False`` fails.

The scratch files written by ``Preprocessing`` and read by ``Solver.setup``
 and the solution files ``x{i}.dat`` keep the PETSc binary
format (big-endian; Mat classid 1211216: header, row lengths, column indices,
complex values; Vec classid 1211214), so files are interchangeable with a real
PETGEM/PETSc run in a complex-scalar build.
"""
from __future__ import annotations

import os

import numpy as np

from .common import Print

MAT_FILE_CLASSID = 1211216
VEC_FILE_CLASSID = 1211214


class MPIEnvironment(object):
    """parallel.py:17-40: rank / num_proc of the SPMD job (one process per GPU)."""

    def __init__(self):
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.rank, self.num_proc = dist.get_rank(), dist.get_world_size()
                self.comm = dist.group.WORLD
                self.machine_name = os.uname().nodename
                return
        except Exception:
            pass
        self.rank = int(os.environ.get("RANK", "0"))
        self.num_proc = int(os.environ.get("WORLD_SIZE", "1"))
        self.comm = None
        self.machine_name = os.uname().nodename


# ---------------------------------------------------------------------------
# host containers with the few petsc4py methods the PETGEM code calls
# ---------------------------------------------------------------------------
class HostMat(object):
    """Dense scratch table (rows x cols, complex PETSc scalars holding reals in .real)."""

    def __init__(self, array):
        self.array = np.ascontiguousarray(array)

    def getSize(self):
        return self.array.shape

    def getSizes(self):
        return (self.array.shape[0], self.array.shape[0]), (self.array.shape[1], self.array.shape[1])

    def getOwnershipRange(self):
        return 0, self.array.shape[0]

    def getRow(self, i):
        return np.arange(self.array.shape[1], dtype=np.int32), self.array[i].astype(np.complex128)

    def getLocalSize(self):
        return self.array.shape


class HostVec(object):
    def __init__(self, array):
        self.array = np.ascontiguousarray(array)

    def getArray(self):
        return self.array

    def getSizes(self):
        return self.array.size, self.array.size

    def getSize(self):
        return self.array.size

    def getOwnershipRange(self):
        return 0, self.array.size

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)


def createSequentialDenseMatrixWithArray(dimension1, dimension2, data):
    """parallel.py:47-62."""
    return HostMat(np.asarray(data).reshape(dimension1, dimension2))


def createSequentialVectorWithArray(data):
    """parallel.py:82-93."""
    return HostVec(np.asarray(data))


def writeParallelDenseMatrix(output_file, data, communicator=None):
    """parallel.py:65-79: dense Mat through a default binary viewer = AIJ layout on disk with
    every entry explicit."""
    a = np.asarray(data.array if isinstance(data, HostMat) else data)
    M, N = a.shape
    with open(output_file, "wb") as fh:
        np.array([MAT_FILE_CLASSID, M, N, M * N], dtype=">i4").tofile(fh)
        np.full(M, N, dtype=">i4").tofile(fh)
        np.tile(np.arange(N, dtype=">i4"), M).tofile(fh)
        np.ascontiguousarray(a, dtype=np.complex128).view(np.float64).astype(">f8").tofile(fh)


def writePetscVector(output_file, data, communicator=None):
    """parallel.py:96-111."""
    if hasattr(data, "to_host"):
        a = data.to_host()
    else:
        a = np.asarray(data.array if isinstance(data, HostVec) else data)
    with open(output_file, "wb") as fh:
        np.array([VEC_FILE_CLASSID, a.size], dtype=">i4").tofile(fh)
        np.ascontiguousarray(a, dtype=np.complex128).view(np.float64).astype(">f8").tofile(fh)


def read_petsc_aij(input_file):
    """PETSc binary Mat -> (rowptr int64, colidx int32, vals complex128, (M, N))."""
    raw = np.fromfile(input_file, dtype=np.uint8)
    hdr = raw[:16].view(">i4")
    if int(hdr[0]) != MAT_FILE_CLASSID:
        Print.master("     %s is not a PETSc binary matrix" % input_file)
        exit(-1)
    M, N, nnz = int(hdr[1]), int(hdr[2]), int(hdr[3])
    off = 16
    rowlens = raw[off:off + 4 * M].view(">i4").astype(np.int64)
    off += 4 * M
    cols = raw[off:off + 4 * nnz].view(">i4").astype(np.int32)
    off += 4 * nnz
    vals = raw[off:off + 16 * nnz].view(">f8").astype(np.float64).view(np.complex128)
    return np.concatenate([[0], np.cumsum(rowlens)]), cols, vals, (M, N)


def readPetscMatrix(input_file, communicator=None):
    """parallel.py:114-129 -> HostMat (the scratch tables are dense: every row has N entries)."""
    rowptr, cols, vals, (M, N) = read_petsc_aij(input_file)
    if vals.size == M * N:
        return HostMat(vals.reshape(M, N))
    dense = np.zeros((M, N), dtype=np.complex128)
    rows = np.repeat(np.arange(M), np.diff(rowptr))
    dense[rows, cols] = vals
    return HostMat(dense)


def readPetscVector(input_file, communicator=None):
    """parallel.py:132-147."""
    raw = np.fromfile(input_file, dtype=np.uint8)
    hdr = raw[:8].view(">i4")
    if int(hdr[0]) != VEC_FILE_CLASSID:
        Print.master("     %s is not a PETSc binary vector" % input_file)
        exit(-1)
    n = int(hdr[1])
    return HostVec(raw[8:8 + 16 * n].view(">f8").astype(np.float64).view(np.complex128))


# ---------------------------------------------------------------------------
# device objects
# ---------------------------------------------------------------------------
class B200Vec(object):
    """Complex128 vector in HBM (reference numbering at the API)."""

    def __init__(self, size, device=None):
        import torch

        from .device import _dev

        self.t = torch.zeros((int(size),), dtype=torch.complex128, device=_dev(device))

    def setValues(self, indices, values, addv=None):
        import torch

        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=self.t.device)
        val = torch.as_tensor(np.asarray(values, dtype=np.complex128), device=self.t.device)
        if addv in (True, "ADD_VALUES", 2):
            self.t.index_add_(0, idx, val)
        else:
            self.t[idx] = val

    def getValues(self, indices):
        import torch

        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=self.t.device)
        return self.t[idx].cpu().numpy()

    def assemblyBegin(self):
        pass

    def assemblyEnd(self):
        pass

    def getArray(self):
        return self.t.cpu().numpy()

    to_host = getArray

    def getSizes(self):
        return self.t.numel(), self.t.numel()

    def getSize(self):
        return self.t.numel()

    def getOwnershipRange(self):
        return 0, self.t.numel()


class B200Mat(object):
    """Complex128 CSR matrix in HBM; filled by the fused assembly kernel, not by setValues."""

    def __init__(self, dimension1, dimension2, nnz=None):
        self.shape = (int(dimension1), int(dimension2))
        self.csr = None      # petgem_b200.device.CSRMatrix in the numbering in use
        self.plan = None     # AssemblyPlan that produced it
        self.perm = None     # reference dof -> numbering in use (device int32)

    def getSize(self):
        return self.shape

    def assemblyBegin(self):
        pass

    def assemblyEnd(self):
        pass

    def setValues(self, rows, cols, values, addv=None):
        Print.master("     B200Mat is assembled by Solver.assembly (fused kernel); per-element setValues is not "
                     "part of the B200 path")
        exit(-1)

    def zeroRowsColumns(self, rows, diag=1.0):
        """A.zeroRowsColumns(rows) with reference dof ids (solver.py:562)."""
        rows = np.asarray(rows, dtype=np.int64)
        if self.perm is not None:
            rows = self.perm.cpu().numpy().astype(np.int64)[rows]
        self.csr.zeroRowsColumns(rows, diag)

    def mult(self, x, y):
        """y = A x with B200Vec arguments in reference numbering."""
        import torch

        if self.perm is None:
            self.csr.mult(x.t, y.t)
        else:
            p = self.perm.to(torch.int64)
            xi = torch.empty_like(x.t)
            xi[p] = x.t
            y.t.copy_(self.csr.mult(xi)[p])


def createParallelMatrix(dimension1, dimension2, nnz, matrix_type, communicator=None):
    """parallel.py:150-177; ``matrix_type`` is ``run['cuda']`` (solver.py:188)."""
    if matrix_type is not True:
        Print.master("     petgem_b200 implements the device path only: set run.cuda: True (no CPU fallback)")
        exit(-1)
    return B200Mat(dimension1, dimension2, nnz)


def createParallelVector(size, vector_type, communicator=None):
    """parallel.py:180-203; ``vector_type`` is ``run['cuda']`` (solver.py:243-244)."""
    if vector_type is not True:
        Print.master("     petgem_b200 implements the device path only: set run.cuda: True (no CPU fallback)")
        exit(-1)
    return B200Vec(size)


def unitary_test():
    """Unitary test for parallel.py script."""

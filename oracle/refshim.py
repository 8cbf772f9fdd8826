"""Import shim for the UNMODIFIED reference (test infrastructure, container-only).

``/root/reference/petgem/{hvfem,mesh,vectors}.py`` import and run under numpy>=1.24
once the removed aliases ``np.float/np.int/np.complex`` are restored and the
modules the reference imports but this image lacks (mpi4py, petsc4py, colorama,
singleton_decorator) are stubbed.  Nothing is copied from
the reference; this file only makes it importable so that ``make_golden.py`` can
record its outputs.  It must never be imported by the product, by ``bench.py``
or by any ``-m gpu`` test: /root/reference does not exist on the GPU box.
"""
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def load():
    np.float = float
    np.int = int
    np.complex = complex

    def _mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class _Comm:
        def Get_rank(self):
            return 0

        def Get_size(self):
            return 1

        def barrier(self):
            pass

    mpi = _mod("mpi4py.MPI", COMM_WORLD=_Comm(), Get_processor_name=lambda: "oracle")
    _mod("mpi4py", MPI=mpi)
    petsc = _mod("petsc4py.PETSc")
    _mod("petsc4py", PETSc=petsc)

    class _Fore:
        def __getattr__(self, key):
            return ""

    _mod("colorama", Fore=_Fore())

    def singleton(cls):
        inst = {}

        def get(*a, **k):
            if cls not in inst:
                inst[cls] = cls(*a, **k)
            return inst[cls]

        return get

    _mod("singleton_decorator", singleton=singleton)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from petgem import hvfem, mesh, vectors  # noqa: E402

    return hvfem, mesh, vectors

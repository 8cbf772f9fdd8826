"""ctypes binding of libpetgem_b200.so (the C ABI declared in include/petgem_b200.h).

The library is built in-tree by :func:`build` (nvcc, sm_100a only).  There is no
fallback: if the shared object is missing or a CUDA device is absent, the product
path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libpetgem_b200.so")
SOURCES = ["pg_element.cu", "pg_plan.cu", "pg_assemble.cu", "pg_linalg.cu", "pg_spmv_blocked.cu", "pg_multi.cu", "pg_krylov.cu", "pg_aux.cu", "pg_tables.cu", "pg_comm.cu", "pg_dmma.cu"]
HEADERS = ["pg_common.cuh", "pg_plan.cuh", "pg_basis.cuh", os.path.join("..", "..", "include", "petgem_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared",
]


LINK_LIBS = []


class PetgemB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


OBJ_DIR = os.path.join(HERE, "build")


def _newer(path: str, deps) -> bool:
    """True if `path` is missing or older than any of `deps`."""
    if not os.path.exists(path):
        return True
    built = os.path.getmtime(path)
    return any(os.path.getmtime(d) > built for d in deps)


def _stale() -> bool:
    return _newer(LIB_PATH, [os.path.join(CSRC, f) for f in SOURCES + HEADERS])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into petgem_b200/libpetgem_b200.so: one object per source
    (compiled in parallel, rebuilt only when the source or a header changed), then one link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in HEADERS]
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if force or _newer(obj, [path] + headers):
            cmd = [nvcc] + compile_flags + ["-c", "-o", obj, path]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + LINK_LIBS
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_i64, _i32, _p, _d = C.c_int64, C.c_int, C.c_void_p, C.c_double

# name -> (restype, argtypes); must list every symbol include/petgem_b200.h declares
SIGNATURES = {
    "pg_version": (C.c_int, []),
    "pg_last_error": (C.c_char_p, []),
    "pg_ndof_element": (C.c_int, [_i32]),
    "pg_ndof_edge": (C.c_int, [_i32]),
    "pg_ndof_face": (C.c_int, [_i32]),
    "pg_ndof_volume": (C.c_int, [_i32]),
    "pg_nexp": (C.c_int, [_i32]),
    "pg_element_geometry": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pg_element_matrices": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p, _p]),
    "pg_element_systems": (C.c_int, [_i64, _i32, _p, _p, _p, _d, _p, _p]),
    "pg_connectivity_dofs": (C.c_int, [_i64, _i32, _p, _p, _i64, _i64, _p, _p]),
    "pg_plan_create": (C.c_int, [_i64, _i32, _p, _p, _i64, _i64, _p, _i64, _i64, C.POINTER(_p), _p]),
    "pg_plan_destroy": (None, [_p]),
    "pg_plan_locality_order": (C.c_int, [_i64, _i32, _p, _p, _i64, _i64, _p, _p]),
    "pg_plan_ranked_order": (C.c_int, [_i64, _i32, _p, _p, _i64, _i64, _p, _p, _p]),
    "pg_plan_num_dofs": (_i64, [_p]),
    "pg_plan_num_entities": (_i64, [_p]),
    "pg_plan_local_rows": (_i64, [_p]),
    "pg_plan_row_begin": (_i64, [_p]),
    "pg_plan_nnz": (_i64, [_p]),
    "pg_plan_contributions": (_i64, [_p]),
    "pg_plan_max_row_length": (C.c_int, [_p]),
    "pg_plan_element_range": (C.c_int, [_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "pg_plan_csr": (C.c_int, [_p, _p, _p, _p]),
    "pg_plan_dof_permutation": (C.c_int, [_p, _p, _p]),
    "pg_plan_entity_aligned_row": (_i64, [_p, _i64]),
    "pg_plan_set_dirichlet": (C.c_int, [_p, _p, _p, _p, _p]),
    "pg_assemble": (C.c_int, [_p, _p, _p, _p, _d, _i32, _d, _p, _p]),
    "pg_zero_rows_columns": (C.c_int, [_i64, _i64, _p, _p, _p, _d, _p, _p]),
    "pg_spmv": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p]),
    "pg_spmv_scaled": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p]),
    "pg_spmv_blocked": (C.c_int, [_p, _p, _p, _p, _p, _p, _p]),
    "pg_spmv_blocked_range": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "pg_spmm_blocked_range": (C.c_int, [_p, _i64, _i64, _p, _p, _i32, _p, _p, _p, _p]),
    "pg_plan_halo_split": (C.c_int, [_p, _p, _i64, C.POINTER(_i64), C.POINTER(_i64), _p]),
    "pg_plan_num_column_entities": (_i64, [_p]),
    "pg_plan_column_starts": (C.c_int, [_p, _p, _p]),
    "pg_csr_diagonal": (C.c_int, [_i64, _i64, _p, _p, _p, _p, _p]),
    "pg_zaxpy": (C.c_int, [_i64, _p, _p, _p, _p]),
    "pg_zaypx": (C.c_int, [_i64, _p, _p, _p, _p]),
    "pg_zaxpbypcz": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pg_zscal": (C.c_int, [_i64, _p, _i32, _p, _p]),
    "pg_zpointwise_mult": (C.c_int, [_i64, _p, _p, _p, _p]),
    "pg_zdotc": (C.c_int, [_i64, _p, _p, _p, _p, _p]),
    "pg_zdotu": (C.c_int, [_i64, _p, _p, _p, _p, _p]),
    "pg_zmdotc": (C.c_int, [_i64, _i32, _p, _i64, _p, _p, _p, _p]),
    "pg_zmaxpy": (C.c_int, [_i64, _i32, _p, _d, _p, _i64, _p, _p]),
    "pg_zmaxpy_nrm2sq": (C.c_int, [_i64, _i32, _p, _d, _p, _i64, _p, _p, _p, _p]),
    "pg_zdiv": (C.c_int, [_p, _p, _i32, _p, _p]),
    "pg_zcopy_scaled": (C.c_int, [_i64, _p, _i32, _p, _p, _p]),
    "pg_dznrm2sq": (C.c_int, [_i64, _p, _p, _p, _p]),
    "pg_reduce_workspace_bytes": (_i64, [_i32]),
    "pg_spmm": (C.c_int, [_i64, _p, _p, _p, _i32, _p, _p, _p, _p]),
    "pg_spmm_blocked": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p]),
    "pg_zbaxpy": (C.c_int, [_i64, _i32, _p, _p, _p, _p]),
    "pg_zbaypx": (C.c_int, [_i64, _i32, _p, _p, _p, _p]),
    "pg_zbscale_rows": (C.c_int, [_i64, _i32, _p, _p, _p, _p]),
    "pg_zbdotu": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p]),
    "pg_zbdotu_w": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p, _p]),
    "pg_cocr_update": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "pg_cocr_direction": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p, _p]),
    "pg_zbnrm2sq": (C.c_int, [_i64, _i32, _p, _p, _p, _p]),
    "pg_zbdiv": (C.c_int, [_i32, _p, _p, _p, _p]),
    "pg_krylov_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "pg_krylov_solve": (C.c_int, [_i64, _p, _p, _p, _p, _p, _i32, _i32, _i32, _d, _i32, _i32, _p, C.POINTER(C.c_int),
                                  C.POINTER(C.c_double), _p]),
    "pg_table_size": (_i64, [_i32]),
    "pg_tables_init": (C.c_int, [_i32, _p, _p]),
    "pg_shape_functions_host": (C.c_int, [_i32, C.c_uint32, _p, _p, _p]),
    "pg_locate_points": (C.c_int, [_i64, _p, _i64, _p, _d, _p, _p]),
    "pg_interpolate_fields": (C.c_int, [_i64, _p, _p, _i32, _p, _p, _p, _p, _i64, _i64, _p, _p, _d, _d, _p, _p]),
    "pg_csem_rhs": (C.c_int, [_i32, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _i64, _d, _d, _p, _p]),
    "pg_rcsr_apply": (C.c_int, [_i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "pg_galerkin_diagonal": (C.c_int, [_i64, _p, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "pg_masked_reciprocal": (C.c_int, [_i64, _p, _p, _p]),
    "pg_graph_begin": (C.c_int, [_p]),
    "pg_graph_end": (C.c_int, [_p, C.POINTER(_p)]),
    "pg_graph_launch": (C.c_int, [_p, _p]),
    "pg_graph_destroy": (None, [_p]),
    "pg_halo_columns": (C.c_int, [_i64, _p, _i64, _i64, _p, C.POINTER(_i64), _p]),
    "pg_halo_remap": (C.c_int, [_i64, _p, _i64, _i64, _p, _i64, _p]),
    "pg_ipc_alloc": (C.c_int, [_i64, C.POINTER(_p), _p]),
    "pg_ipc_open": (C.c_int, [_p, C.POINTER(_p)]),
    "pg_ipc_close": (C.c_int, [_p]),
    "pg_ipc_free": (C.c_int, [_p]),
    "pg_comm_ctrl_bytes": (_i64, []),
    "pg_comm_create": (C.c_int, [_i32, _i32, C.POINTER(_p), _d, C.POINTER(_p)]),
    "pg_comm_destroy": (None, [_p]),
    "pg_comm_allreduce": (C.c_int, [_p, _i32, _p, _p, _p]),
    "pg_comm_push": (C.c_int, [_p, _i32, _i32, _p, _p, C.POINTER(_i64), C.POINTER(_p), _p]),
    "pg_comm_wait": (C.c_int, [_p, _i32, C.c_uint32, _p]),
    "pg_comm_ack": (C.c_int, [_p, _i32, C.c_uint32, _p]),
    "pg_comm_status": (C.c_int, [_p, _p]),
    "pg_l2_persist": (C.c_int, [_p, _i64, _d, _p]),
    "pg_l2_persist_capacity": (_i64, []),
    "pg_l2_fetch_granularity": (C.c_int, [_i32]),
    "pg_tune_spmv_hints": (C.c_int, [_i32]),
    "pg_tune_spmm_prefetch": (C.c_int, [_i32]),
    "pg_tune_spmv_chunk": (C.c_int, [_i32]),
    "pg_phi_gemm_workspace_doubles": (_i64, [_i64, _i32, _i32]),
    "pg_element_matrices_phi_gemm": (C.c_int, [_i64, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "pg_cocr_direction_dot": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pg_spmv_dot_workspace_bytes": (_i64, [_p, _i32]),
    "pg_spmm_blocked_dot": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "pg_cocg_step": (C.c_int, [_i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
}

_LIB = None


def lib() -> C.CDLL:
    """Load the shared library (raises if it has not been built)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PetgemB200Error(
                "libpetgem_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pg_last_error().decode(errors="replace")
        raise PetgemB200Error("%s failed with status %d: %s" % (what or "petgem_b200 call", rc, msg))


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr():
    """The current torch CUDA stream as a cudaStream_t."""
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise PetgemB200Error("petgem_b200 needs a CUDA device (B200); there is no CPU fallback")

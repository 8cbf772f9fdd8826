#!/usr/bin/env python3
"""A/B of the two forms of computeElementalMatrices (hvfem.py:223-316) at p = 3..6 on one B200:
  table:    pg_element_matrices, the K = 12 contraction of per-element geometric factors with the
            reference-element table (what pg_assemble fuses with the scatter; CUDA cores, 24 n^2 flops);
  phi-gemm: pg_element_matrices_phi_gemm, Me = +-U^T U, Ke = +-V^T V on the FP64 tensor pipe (DMMA
            m8n8k4), 4 n^2 3 ngauss flops, operands built from the expanded basis at the Gauss points.
Both write the same [T, n, n] Me, Ke; parity is checked (<= 1e-11 of max).  One JSON object on stdout."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from petgem_b200 import basis  # noqa: E402
from petgem_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from petgem_b200.device import ElementData, element_matrices  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--orders", default="3,4,5,6")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
L = lib()
REF_NGAUSS = {1: 4, 2: 11, 3: 24, 4: 43, 5: 126, 6: 210}  # rule "order 2p" of hvfem.py:1055-1610
COUNT = {3: 4096, 4: 2048, 5: 1024, 6: 512}
out = {"device": torch.cuda.get_device_name(0), "orders": {}}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for p in [int(v) for v in args.orders.split(",")]:
    T = COUNT[p]
    tab = bench.build_case(10, p, vti=0.5)  # 6000 tets, VTI conductivity
    rows = bench.host_rows(tab)
    sel = slice(0, T)
    el = ElementData(rows["nodes"][sel], rows["elemsN"][sel], rows["elemsE"][sel], rows["edgesNodes"][sel],
                     rows["facesEdges"][sel], rows["elemsF"][sel], rows["sigma"][sel], tab["nEdges"], tab["nFaces"],
                     device=dev)
    geo, code = el.geometry()
    n = basis.ndof_element(p)
    pts, wts = basis.tet_quadrature(2 * p)
    ng = pts.shape[0]
    N, C = basis.evaluate_expanded(p, pts)
    phiN = torch.as_tensor(np.ascontiguousarray(N), device=dev)
    phiC = torch.as_tensor(np.ascontiguousarray(C), device=dev)
    w = torch.as_tensor(np.ascontiguousarray(wts), device=dev)
    work = torch.empty((L.pg_phi_gemm_workspace_doubles(T, p, ng),), dtype=torch.float64, device=dev)
    Me2 = torch.empty((T, n, n), dtype=torch.float64, device=dev)
    Ke2 = torch.empty_like(Me2)

    def phi(stage):
        check(L.pg_element_matrices_phi_gemm(T, p, ptr(el.nodes), ptr(el.sigma), ptr(code), ng, ptr(phiN), ptr(phiC),
                                             ptr(w), ptr(work), stage, ptr(Me2), ptr(Ke2), stream_ptr()), "phi_gemm")

    Me, Ke = element_matrices(p, geo, code)
    phi(0)
    torch.cuda.synchronize()
    eM = float((Me - Me2).abs().max() / Me.abs().max())
    eK = float((Ke - Ke2).abs().max() / Ke.abs().max())
    t_table = timed(lambda: element_matrices(p, geo, code), args.reps)
    t_ops = timed(lambda: phi(1), args.reps)
    t_gemm = timed(lambda: phi(2), args.reps)
    kpad = (3 * ng + 3) // 4 * 4
    ld = (n + 15) // 16 * 16
    nb = ld // 16
    flops_done = 2.0 * T * (nb * (nb + 1) // 2) * 256 * kpad * 2   # DMMA flops issued (upper blocks, both matrices)
    flops_full = 2.0 * T * 2 * n * n * 3 * ng                       # 4 n^2 3 ngauss per element
    out["orders"][str(p)] = {
        "elements": T, "n": n, "ngauss_conical_rule": ng, "ngauss_reference_rule": REF_NGAUSS[p],
        "parity_max_rel": {"Me": eM, "Ke": eK},
        "table_us_per_element": 1e3 * t_table / T,
        "phi_gemm_us_per_element": {"operands": 1e3 * t_ops / T, "dmma_gemms": 1e3 * t_gemm / T,
                                    "total": 1e3 * (t_ops + t_gemm) / T,
                                    "dmma_gemms_scaled_to_reference_ngauss": 1e3 * t_gemm / T * REF_NGAUSS[p] / ng},
        "dmma_tflops_issued": flops_done / (t_gemm * 1e-3) / 1e12,
        "flops_per_element": {"table_24n2": 24.0 * n * n, "phi_gemm_4n2_3ng_reference_rule": 4.0 * n * n * 3 * REF_NGAUSS[p],
                              "phi_gemm_symmetric_half_issued_here": flops_done / T},
        "phi_gemm_over_table": (t_ops + t_gemm) / t_table,
        "phi_gemm_at_dmma_peak_37tflops_reference_rule_us": flops_full * REF_NGAUSS[p] / ng / 2 / T / 37e12 * 1e6,
    }
    del Me, Ke, Me2, Ke2, work, phiN, phiC
    torch.cuda.empty_cache()
print(json.dumps(out))

"""Golden receiver fields for p = 2 and p = 3 on the reference's test mesh (TEST INFRASTRUCTURE).

Run once in the build container:  python oracle/make_golden_fields.py 2 3
The whole case1 pipeline (examples/case1: f = 2 Hz, x-directed unit dipole at (1750, 1750, -975),
sigma by physical tag, Dirichlet on the boundary dofs) is evaluated with the oracle only -- the
restatement of the reference pinned by tests/test_oracle_golden.py -- on the topology recorded from
the unmodified reference (tests/golden/test_mesh_topology.npz):
  element systems (solver.py:214-224) -> MatSetValues/assembly (solver.py:230-235) ->
  zeroRowsColumns (solver.py:562-574) -> solve -> fieldInterpolator (postprocessing.py:479-616).
The solve is a sparse direct solve (scipy splu, minutes at p = 2) or, at p = 3 where the LU fill is out
of reach of this container, the oracle's COCR+Jacobi run to 1e-13 with the true residual recorded.
Stored: the receiver fields [58, 6] (Ex, Ey, Ez, Hx, Hy, Hz), ||x||, the residual of the stored solve
and 64 sampled solution entries.  The GPU parity test compares its own assembly + Krylov solve with
these fields to 1e-6 (north_star).
"""
import os
import sys
import time

import numpy as np
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import petgem_oracle as oracle  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")


def main(orders):
    topo = dict(np.load(os.path.join(GOLD, "test_mesh_topology.npz")))
    rec = np.load(os.path.join(GOLD, "case1_receivers.npy"))
    omega, mu = 2 * np.pi * 2.0, 4e-7 * np.pi
    sig = np.array([1.0, 0.01, 1.0, 3.3333])[topo["tags"] - 1]
    sigma = np.stack([sig, sig], axis=1)
    elemsN, elemsE, elemsF = topo["elemsN"], topo["elemsE"].astype(np.int64), topo["elemsF"].astype(np.int64)
    T = elemsN.shape[0]
    for p in orders:
        t0 = time.time()
        n = p * (p + 2) * (p + 3) // 2
        dofs, *_, N = oracle.compute_connectivity_dofs(elemsE, elemsF, p)
        Ae = np.zeros((T, n, n), dtype=np.complex128)
        for t in range(T):
            Ae[t] = oracle.element_system(topo["nodes"][elemsN[t]], elemsN[t], elemsE[t], topo["edgesNodes"][elemsE[t]],
                                          topo["facesE"][elemsF[t]], sigma[t], p, omega, mu)
        rp, ci, v = oracle.assemble_global(Ae, dofs, N)
        del Ae
        bd = topo["boundary_dofs_p%d" % p]
        v = oracle.zero_rows_columns(rp, ci, v, bd)
        A = oracle.to_scipy(rp, ci, v)
        src = np.array([1750.0, 1750.0, -975.0])
        te = int(oracle.locate_points(topo["nodes"], elemsN, src[None, :])[0])
        b = oracle.csem_rhs(N, topo["nodes"][elemsN[te]], elemsN[te], elemsE[te], topo["edgesNodes"][elemsE[te]],
                            topo["facesE"][elemsF[te]], dofs[te], p, src, 0.0, 0.0, 1.0, 1.0, omega, mu)
        b[bd] = 0.0
        print("p=%d N=%d nnz=%d assembled in %.0f s" % (p, N, v.size, time.time() - t0), flush=True)
        t0 = time.time()
        if p <= 2:
            x = spla.splu(A.tocsc()).solve(b)
            how, its = "splu", 0
        else:
            dinv = 1.0 / A.diagonal()
            x, its, hist = oracle.cocr(lambda u: A @ u, b, rtol=1e-13, maxit=200000, dinv=dinv)
            how = "oracle cocr+jacobi rtol 1e-13"
        res = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        print("   solve (%s, %d its) %.0f s, true residual %.2e" % (how, its, time.time() - t0, res), flush=True)
        F = oracle.field_interpolator(x, topo["nodes"], elemsN, elemsE, topo["edgesNodes"], elemsF, topo["facesE"], dofs,
                                      rec, p, omega, mu)
        sel = np.linspace(0, N - 1, 64).astype(np.int64)
        np.savez_compressed(os.path.join(GOLD, "test_mesh_fields_p%d.npz" % p), fields=F, xnorm=np.linalg.norm(x),
                            residual=res, solver=np.array(how), iterations=its, x_sel=sel, x_val=x[sel],
                            omega=omega, mu=mu, source=src)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [2, 3])

"""Drop-in ``Solver`` for the B200 path: same call sequence as ``petgem/solver.py``
(``kernel.py:64-73``): ``Solver().setup(inputSetup)``, ``.assembly(inputSetup)``,
``.run(inputSetup)``.

setup    reads the scratch files of ``Preprocessing`` (solver.py:66-123) and moves
         the per-element rows to HBM once;
assembly replaces the Python element loop + MatSetValues (solver.py:191-235) by the
         geometry kernel and the fused element-matrix / row-gather kernel, and
         builds the CSEM right-hand side (solver.py:247-316) for the one source
         element on the host;
run      applies MatZeroRowsColumns (solver.py:562-574) and solves with the Krylov
         drivers configured from the PETSc options file (solver.py:584-590), then
         writes ``x{i}.dat`` (solver.py:593-594) for ``Postprocessing``.
Leaves ``self.A``, ``self.b[i]``, ``self.x[i]`` like the reference.
"""
from __future__ import annotations

import numpy as np

from . import hvfem, krylov
from .common import Print, Timers
from .parallel import (MPIEnvironment, createParallelMatrix, createParallelVector, readPetscMatrix,
                       readPetscVector, writePetscVector)


class Solver():
    """Class for solver."""

    def __init__(self):
        self.petsc_options = {"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": "1e-8"}
        self.ksp_results = []

    def setOptions(self, options):
        """Options of the PETSc options file given on the command line (kernel.py:15)."""
        self.petsc_options = dict(options)

    # ------------------------------------------------------------------------------
    def setup(self, inputSetup):
        Timers()["Setup"].start()
        output = inputSetup.output
        out_dir = output.get('directory_scratch')
        Print.master('     Importing files')
        parEnv = MPIEnvironment()

        def real_table(name, dtype):
            return np.ascontiguousarray(readPetscMatrix(out_dir + '/' + name).array.real.astype(dtype))

        self.nodes = real_table('nodes.dat', np.float64)                # [T,12]
        self.elemsN = real_table('meshConnectivity.dat', np.int32)      # [T,4]
        self.elemsE = real_table('edges.dat', np.int32)                 # [T,6]
        self.edgesNodes = real_table('edgesNodes.dat', np.int32)        # [T,12]
        self.elemsF = real_table('faces.dat', np.int32)                 # [T,4]
        self.facesEdges = real_table('facesEdges.dat', np.int32)        # [T,12]
        self.dofs = real_table('dofs.dat', np.int64)                    # [T,n]
        self.sigmaModel = real_table('conductivityModel.dat', np.float64)
        tmp = readPetscVector(out_dir + '/nnz.dat')
        self.nnz = tmp.getArray().real.astype(np.int64)
        self.total_num_dofs = tmp.getSizes()[1]
        mode = inputSetup.model.get('mode')
        if mode == 'csem':
            self.boundaries = readPetscVector(out_dir + '/boundaries.dat')
            # every rank keeps replicated b/x vectors, so every rank reads the (tiny) source record
            self.source_data = readPetscVector(out_dir + '/source.dat')
        elif mode == 'mt':
            self.boundaries = readPetscMatrix(out_dir + '/boundaryElements.dat')

        # consistency of the numbering the kernels rebuild from (edges.dat, faces.dat)
        p = inputSetup.run.get('nord')
        nE, nF = int(self.elemsE.max()) + 1, int(self.elemsF.max()) + 1
        k = min(64, self.dofs.shape[0])
        ref = hvfem.dofs_of_elements(self.elemsE[:k], self.elemsF[:k], np.arange(k), nE, nF, p)
        if not np.array_equal(self.dofs[:k], ref):
            Print.master('     dofs.dat is not consistent with edges.dat/faces.dat')
            exit(-1)

        from .device import ElementData
        self.elems = ElementData(self.nodes, self.elemsN, self.elemsE, self.edgesNodes, self.facesEdges, self.elemsF,
                                 self.sigmaModel, nE, nF)
        Timers()["Setup"].stop()
        return

    # ------------------------------------------------------------------------------
    def assembly(self, inputSetup):
        import torch

        from .device import MU0, AssemblyPlan, CSRMatrix

        Timers()["Assembly"].start()
        model, run = inputSetup.model, inputSetup.run
        Print.master('     Assembling linear system')
        parEnv = MPIEnvironment()
        basis_order = run.get('nord')
        num_polarizations = run.get('num_polarizations')
        mode = model.get('mode')
        data_model = model.get(mode)
        frequency = data_model.get('source').get('frequency') if mode == 'csem' else data_model.get('frequency')
        omega = frequency * 2. * np.pi
        mu = MU0
        Const = 1j * omega * mu
        self.omega, self.mu = omega, mu

        # ---- LHS: symbolic phase once, numeric phase fused (solver.py:188-235) ----
        order = run.get('b200_order', 'locality')
        self.plan = AssemblyPlan(self.elems, basis_order, order=order)
        if self.plan.N != self.total_num_dofs:
            Print.master('     Number of DOFs is not consistent')
            exit(-1)
        self.row_begins = [0]
        if parEnv.num_proc > 1:
            # PETSc-style contiguous row blocks (aligned to entities): each rank assembles the rows it owns
            N, world = self.plan.N, parEnv.num_proc
            self.row_begins = [0] + [self.plan.entity_aligned_row(N * r // world) for r in range(1, world)]
            ends = self.row_begins[1:] + [N]
            order_host = self.plan.order_host if self.plan.order_host is not None else 'reference'
            self.plan = None  # release the global symbolic plan before the owned-rows plan is built
            torch.cuda.empty_cache()
            self.plan = AssemblyPlan(self.elems, basis_order, order=order_host,
                                     row_range=(self.row_begins[parEnv.rank], ends[parEnv.rank]))
        geo, code = self.elems.geometry(self.plan.element_range)
        vals = self.plan.assemble(geo, code, omega, mu)
        rowptr, colidx = self.plan.csr()
        self.A = createParallelMatrix(self.total_num_dofs, self.total_num_dofs, self.nnz, run.get('cuda'))
        self.A.plan = self.plan
        self.A.csr = CSRMatrix(rowptr, colidx, vals, self.plan.N, self.plan.row_begin, plan=self.plan)
        self.A.perm = self.plan.dof_permutation() if order != 'reference' else None

        # ---- RHS ----
        self.b, self.x = [], []
        for i in np.arange(num_polarizations):
            self.b.append(createParallelVector(self.total_num_dofs, run.get('cuda')))
            self.x.append(createParallelVector(self.total_num_dofs, run.get('cuda')))
        if mode == 'csem':
            src = data_model.get('source')
            position = np.asarray(src.get('position'), dtype=np.float64)
            rot = hvfem.computeSourceVectorRotation(src.get('azimuth'), src.get('dip'))
            moment = src.get('current') * src.get('length')
            field = rot[0] * np.array([moment, 0., 0.]) + rot[1] * np.array([0., moment, 0.]) \
                + rot[2] * np.array([0., 0., moment])
            # b and x are replicated on every rank (the reference inserts on rank 0 into a distributed Vec);
            # the basis of the source element is evaluated at the dipole position on the device (pg_csem_rhs)
            from .device import csem_rhs
            sd = self.source_data.getArray().real
            nodesEle = sd[0:4].astype(np.int64)
            dofsSource = sd[50:].astype(np.int64)
            # source.dat holds the rows of the source element (preprocessing.py:404-464): find it in the mesh
            hit = np.nonzero((self.elemsN == nodesEle[None, :]).all(axis=1))[0]
            if hit.size == 0 or not np.array_equal(self.dofs[hit[0]], dofsSource):
                Print.master('     source.dat is not consistent with the mesh files')
                exit(-1)
            te = int(hit[0])
            t0, t1 = self.plan.element_range
            if not (t0 <= te < t1):  # orientation codes exist for this rank's elements only
                self.elems.geometry((te, te + 1), out=(geo, code))
            csem_rhs(self.elems, basis_order, code, te, position, field, omega, self.b[0].t, mu)
        elif mode == 'mt':
            # Neumann excitation from the 1-D layered-earth solution on the box sides (solver.py:318-512):
            # host numpy on every rank (b is replicated), like the reference's serial loops
            from . import mt
            rows = self.boundaries.array.real
            if rows.shape[1] != 53 + self.dofs.shape[1]:
                Print.master('     boundaryElements.dat is not consistent with the basis order')
                exit(-1)
            za, zb = float(self.nodes[:, 2::3].max()), float(self.nodes[:, 2::3].min())  # solver.py:381-401
            pols = data_model.get('polarization')
            for tmp in pols:
                if tmp not in ('x', 'y'):
                    Print.master('     MT polarization mode not supported.')
                    exit(-1)
            rhs = mt.mt_rhs(rows, za, zb, basis_order, omega, mu, list(pols), self.total_num_dofs)
            for i in np.arange(num_polarizations):
                self.b[i].t.copy_(torch.as_tensor(rhs[i], device=self.b[i].t.device))
        for i in np.arange(num_polarizations):
            self.b[i].assemblyBegin()
            self.b[i].assemblyEnd()
        torch.cuda.synchronize()
        Timers()["Assembly"].stop()
        return

    # ------------------------------------------------------------------------------
    def run(self, inputSetup):
        import torch

        model, run, output = inputSetup.model, inputSetup.run, inputSetup.output
        out_dir = output.get('directory_scratch')
        num_polarizations = run.get('num_polarizations')
        mode = model.get('mode')
        Print.master('     Solving linear system')
        if mode == 'csem':
            Timers()["SetBoundaries"].start()
            bd = np.real(self.boundaries.getArray()).astype(np.int64)
            self.A.zeroRowsColumns(bd)                                   # solver.py:562
            self.b[0].setValues(bd, np.zeros(bd.size, dtype=np.complex128))  # solver.py:565-567
            self.A.assemblyBegin()
            self.A.assemblyEnd()
            Timers()["SetBoundaries"].stop()

        Timers()["Solver"].start()
        self.ksp_results = []
        parEnv = MPIEnvironment()
        perm = self.A.perm.to(torch.int64) if self.A.perm is not None else None
        ctx = krylov.DistContext(self.row_begins, self.plan.N) if parEnv.num_proc > 1 else None
        lo, hi = self.plan.row_begin, self.plan.row_begin + self.plan.local_rows

        def to_internal(b):
            if perm is None:
                return b
            bi = torch.empty_like(b)
            bi[perm] = b
            return bi

        def collect(xi):
            if ctx is not None:  # collect the owned blocks into the replicated solution vector
                send = torch.zeros((ctx.max_rows,), dtype=torch.complex128, device=xi.device)
                full = torch.zeros((ctx.world * ctx.max_rows,), dtype=torch.complex128, device=xi.device)
                ctx.gather(xi, send, full)
                xi = torch.cat([full[r * ctx.max_rows:r * ctx.max_rows + ctx.sizes[r]] for r in range(ctx.world)])
            return xi[perm] if perm is not None else xi

        try:
            if num_polarizations > 1:
                # the right-hand sides share A (the two MT polarizations): one pass over the matrix per
                # iteration for all of them when the solver type allows it (krylov.solve_multi)
                B = torch.stack([to_internal(self.b[i].t)[lo:hi] for i in np.arange(num_polarizations)],
                                dim=1).contiguous()
                X, results = krylov.solve_multi(self.A.csr, B, self.petsc_options, ctx=ctx)   # solver.py:589
                self.ksp_results = list(results)
                for res in results:
                    if not np.all(res.converged):
                        Print.master('     KSP did not converge: %s after %d iterations' % (res.reason, res.iterations))
                xs = [X[:, i].contiguous() for i in range(num_polarizations)]
            else:
                res = krylov.solve(self.A.csr, to_internal(self.b[0].t)[lo:hi].contiguous(), self.petsc_options,
                                   ctx=ctx)                                                     # solver.py:589
                self.ksp_results.append(res)
                if not res.converged:
                    Print.master('     KSP did not converge: %s after %d iterations' % (res.reason, res.iterations))
                xs = [res.x]
        except krylov.UnsupportedSolverError as err:
            Print.master('     ' + str(err))
            exit(-1)
        for i in np.arange(num_polarizations):
            self.x[i].t.copy_(collect(xs[i]))
            if parEnv.rank == 0:
                writePetscVector(out_dir + '/x' + str(i) + '.dat', self.x[i])
        torch.cuda.synchronize()
        Timers()["Solver"].stop()
        return


def unitary_test():
    """Unitary test for solver.py script."""

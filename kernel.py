#!/usr/bin/env python3
"""PETGEM kernel on the B200 path: same command line and stage order as the reference
(/root/reference/kernel.py:6-92):

    python3 kernel.py -options_file <petsc.opts> <params.yaml>          (run.cuda: True)

The PETSc options file (ksp_type, pc_type, ksp_rtol, ...) is parsed by petgem_b200.krylov.
"""
if __name__ == '__main__':
    import os
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from petgem_b200.common import InputParameters, Print, Timers
    from petgem_b200.krylov import parse_petsc_options
    from petgem_b200.parallel import MPIEnvironment
    from petgem_b200.postprocessing import Postprocessing
    from petgem_b200.preprocessing import Preprocessing
    from petgem_b200.solver import Solver

    args = sys.argv[1:]
    options = {}
    if '-options_file' in args:
        k = args.index('-options_file')
        options = parse_petsc_options(args[k + 1])
        del args[k:k + 2]
    if not args:
        print('usage: kernel.py -options_file <petsc.opts> <params.yaml>')
        sys.exit(-1)

    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')

    par_env = MPIEnvironment()
    input_setup = InputParameters(args[-1], par_env)   # sys.argv[3] in the reference (kernel.py:35)
    input_setup.petsc_options = options  # the results file records -ksp_type (postprocessing.py:381-383)
    Timers(input_setup.output['directory'])
    Print.header()
    Print.master(' ')
    Print.master('  Data preprocessing')
    preprocessing = Preprocessing()
    preprocessing.run(input_setup)
    Print.master(' ')
    Print.master('  Run modelling')
    solver = Solver()
    if options:
        solver.setOptions(options)
    solver.setup(input_setup)
    solver.assembly(input_setup)
    solver.run(input_setup)
    for i, res in enumerate(solver.ksp_results):
        # one entry per solve; a lockstep solve of several right-hand sides reports its worst residual
        rel = np.max(np.asarray(res.residuals[-1]) / np.maximum(np.asarray(res.residuals[0]), 1e-300))
        Print.master('     KSP %d: %s, %d iterations, |r|/|r0| = %.3e' % (i, res.reason, res.iterations, rel))
    del solver
    Print.master(' ')
    Print.master('  Data postprocessing')
    postprocessing = Postprocessing()
    postprocessing.run(input_setup)
    Print.master('  Timers (s): ' + ', '.join('%s %.3f' % kv for kv in Timers().items().items()))

"""Thin stand-in for ``petgem/postprocessing.py``: receiver fields from ``x{i}.dat``.

``fieldInterpolator`` follows postprocessing.py:479-616 (locate receivers, evaluate the basis at
the receiver, E = sum_j x_j N_j, H = sum_j x_j curl N_j / (i omega mu)), vectorised with the
product's basis tables on the device (pg_locate_points, pg_interpolate_fields).  Output: the reference's
results file ``<directory>/<mode>_petgemV<version>_<date>.h5`` with its schema (postprocessing.py:341-461:
groups ``machine`` and ``model``, E/H fields per component, MT impedance / apparent resistivity / phase /
tipper), written by ``h5lite.write_classic`` in the layout h5py uses (h5py is not a dependency), plus ``<directory>/fields.npz`` with the same
arrays; the scratch files are removed afterwards like the reference does (postprocessing.py:463-472) unless
the parameter file says ``remove_scratch: False``.
"""
from __future__ import annotations

import numpy as np

from .common import Print, Timers
from .parallel import MPIEnvironment, readPetscVector
from .preprocessing import read_receivers


def fieldInterpolator(solution_vector, nodes, elemsN, elemsE, edgesN, elemsF, facesE, dof_connectivity, points,
                      inputSetup):
    """postprocessing.py:479-616 -> fields [nPoints, 6] complex (Ex,Ey,Ez,Hx,Hy,Hz)."""
    model, run = inputSetup.model, inputSetup.run
    p = run.get('nord')
    mode = model.get('mode')
    data_model = model.get(mode)
    frequency = data_model.get('source').get('frequency') if mode == 'csem' else data_model.get('frequency')
    omega, mu = frequency * 2. * np.pi, 4. * np.pi * 1e-7
    import torch

    from .device import ElementData, interpolate_fields, locate_points

    x = np.asarray(solution_vector.getArray() if hasattr(solution_vector, 'getArray') else solution_vector)
    points = np.atleast_2d(points)
    # receivers are located and the basis evaluated at them on the device (pg_locate_points,
    # pg_interpolate_fields); x is in the reference dof numbering, like x{i}.dat
    elemsN = np.asarray(elemsN)
    el = ElementData.from_mesh(nodes, elemsN, elemsE, edgesN, elemsF, facesE, np.ones((elemsN.shape[0], 2)))
    idx_dev = locate_points(el, points)
    idx = idx_dev.cpu().numpy()
    lost = np.nonzero(idx < 0)[0]
    if lost.size:
        Print.master('        The following receivers were not located and will not be taken into account ' + str(lost))
        points = points[idx >= 0]
        idx_dev = idx_dev[torch.as_tensor(idx >= 0, device=idx_dev.device)].contiguous()
        if points.shape[0] == 0:
            Print.master('     No point has been found. Nothing to do. Aborting')
            exit(-1)
    _, code = el.geometry()
    xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.complex128), device=el.device)
    fields = interpolate_fields(el, p, code, xd, points, idx_dev, omega, mu).cpu().numpy()
    return fields


def computeImpedance(fields, omega, mu):
    """postprocessing.py:619-688 -> (apparent resistivity [4], phase [4], tipper [2], impedance [4]), each
    entry an array over the receivers, components ordered xx, xy, yx, yy (tipper: x, y).  The impedance
    tensor solves E = Z H for the two polarizations, the tipper Hz = T (Hx, Hy): 2x2 systems per receiver."""
    f1, f2 = np.asarray(fields[0]), np.asarray(fields[1])
    E = np.stack([np.stack([f1[:, 0], f2[:, 0]], axis=-1), np.stack([f1[:, 1], f2[:, 1]], axis=-1)], axis=-2)
    H = np.stack([np.stack([f1[:, 3], f2[:, 3]], axis=-1), np.stack([f1[:, 4], f2[:, 4]], axis=-1)], axis=-2)
    Hz = np.stack([f1[:, 5], f2[:, 5]], axis=-1)[:, None, :]
    Hinv = np.linalg.inv(H)
    Z = E @ Hinv
    T = (Hz @ Hinv)[:, 0, :]
    impedance = [Z[:, 0, 0], Z[:, 0, 1], Z[:, 1, 0], Z[:, 1, 1]]
    apparent_resistivity = [np.abs(z) ** 2 / (mu * omega) for z in impedance]
    phase = [(np.degrees(np.arctan(-np.imag(z) / np.real(z)))) % 360 for z in impedance]
    return apparent_resistivity, phase, [T[:, 0], T[:, 1]], impedance


def results_tree(inputSetup, out, total_num_dofs=None, solver_type='None', date=None):
    """The groups and datasets of the reference's results file (postprocessing.py:346-458) as a nested dict."""
    import platform
    from datetime import datetime

    model, run, output = inputSetup.model, inputSetup.run, inputSetup.output
    mode = model.get('mode')
    data_model = model.get(mode)
    npol = run.get('num_polarizations')
    uname = platform.uname()
    machine = {'machine': uname.node + '. ' + uname.release + '. ' + uname.processor,
               'num_processors': MPIEnvironment().num_proc, 'petgem_version': CODE_VERSION}
    m = {'date': (date or datetime.today()).isoformat(), 'mesh_file': str(model.get('mesh')),
         'receivers_file': str(model.get('receivers')), 'nord': run.get('nord'),
         'dof': int(total_num_dofs) if total_num_dofs is not None else -1, 'cuda': bool(run.get('cuda')),
         'vtk': bool(output.get('vtk')), 'mode': mode, 'num-polarizations': npol, 'solver': solver_type,
         'run-time (s)': float(out.get('run_time_s', 0.0))}
    if run.get('conductivity_from_file'):
        m['sigma-file'] = str(data_model.get('sigma').get('file'))
    else:
        m['sigma_horizontal (S/m)'] = np.asarray(data_model.get('sigma').get('horizontal'), dtype=np.float64)
        m['sigma_vertical (S/m)'] = np.asarray(data_model.get('sigma').get('vertical'), dtype=np.float64)
    comp = ('x', 'y', 'z')
    if mode == 'csem':
        src = data_model.get('source')
        m['frequency (Hz)'] = float(src.get('frequency'))
        m['source_position (m)'] = np.asarray(src.get('position'), dtype=np.float64)
        m['source_azimuth (deg)'] = float(src.get('azimuth'))
        m['source_dip (deg)'] = float(src.get('dip'))
        m['source_current (Am)'] = float(src.get('current'))
        m['source_length (m)'] = float(src.get('length'))
        f = out['fields_0']
        m['E-fields'] = {c: f[:, i] for i, c in enumerate(comp)}
        m['H-fields'] = {c: f[:, 3 + i] for i, c in enumerate(comp)}
    else:
        m['frequency (Hz)'] = float(data_model.get('frequency'))
        pol = data_model.get('polarization')
        m['polarization'] = str(pol)
        for i in range(npol):
            f = out['fields_%d' % i]
            m['E-fields_mode_' + pol[i]] = {c: f[:, j] for j, c in enumerate(comp)}
            m['H-fields_mode_' + pol[i]] = {c: f[:, 3 + j] for j, c in enumerate(comp)}
        if 'impedance' in out:
            four = ('xx', 'xy', 'yx', 'yy')
            m['impedance'] = {k: out['impedance'][i] for i, k in enumerate(four)}
            m['apparent_resistivity'] = {k: out['apparent_resistivity'][i] for i, k in enumerate(four)}
            m['phase'] = {k: out['phase'][i] for i, k in enumerate(four)}
            m['tipper'] = {'x': out['tipper'][0], 'y': out['tipper'][1]}
    return {'machine': machine, 'model': m}


CODE_VERSION = '1.0'  # postprocessing.py:333 (code_version)


def write_results_h5(inputSetup, out):
    """<directory>/<mode>_petgemV<version>_<date>.h5 (postprocessing.py:342-344) -> its path."""
    from datetime import datetime

    from . import h5lite

    mode = inputSetup.model.get('mode')
    path = (inputSetup.output.get('directory') + '/' + mode + '_petgemV' + CODE_VERSION + '_'
            + str(datetime.today().strftime('%Y-%m-%d')) + '.h5')
    opts = getattr(inputSetup, 'petsc_options', None) or {}
    # classic layout (version-0 superblock, symbol-table groups): what h5py itself writes by default
    h5lite.write_classic(path, results_tree(inputSetup, out, total_num_dofs=out.get('total_num_dofs'),
                                            solver_type=str(opts.get('ksp_type', 'None'))))
    return path


class Postprocessing():
    """Class for postprocessing."""

    def __init__(self):
        return

    def run(self, inputSetup):
        Timers()["Postprocessing"].start()
        if MPIEnvironment().rank == 0:
            out_dir = inputSetup.output.get('directory_scratch')
            tab = np.load(out_dir + '/mesh_tables.npz')
            receivers = read_receivers(inputSetup.model.get('receivers'))
            out = {}
            for i in range(inputSetup.run.get('num_polarizations')):
                x = readPetscVector(out_dir + '/x%d.dat' % i)
                out['fields_%d' % i] = fieldInterpolator(x, tab['nodes'], tab['elemsN'], tab['elemsE'], tab['edgesNodes'],
                                                         tab['elemsF'], tab['facesE'], tab['dofs'], receivers, inputSetup)
            if inputSetup.model.get('mode') == 'mt' and inputSetup.run.get('num_polarizations') == 2:
                freq = inputSetup.model.get('mt').get('frequency')
                res, phase, tipper, imp = computeImpedance([out['fields_0'], out['fields_1']],
                                                           2. * np.pi * freq, 4. * np.pi * 1e-7)
                out.update(apparent_resistivity=np.stack(res), phase=np.stack(phase), tipper=np.stack(tipper),
                           impedance=np.stack(imp))
            out['receiver_coordinates'] = receivers
            out['total_num_dofs'] = int(np.asarray(tab['dofs']).max()) + 1
            out['run_time_s'] = Timers().elapsed('Assembly') + Timers().elapsed('Solver')
            np.savez(inputSetup.output.get('directory') + '/fields.npz', **out)
            self.fields = out
            self.output_file = write_results_h5(inputSetup, out)
            # scratch clean-up (postprocessing.py:463-472): the whole scratch directory when it was given
            # in the parameter file, else only the PETSc binary files written next to the results
            import os
            import shutil
            if inputSetup.output.get('remove_scratch'):
                shutil.rmtree(out_dir, ignore_errors=True)
            elif os.path.abspath(out_dir) == os.path.abspath(inputSetup.output.get('directory')):
                for name in os.listdir(out_dir):
                    if name.endswith('.dat') or name.endswith('.info'):
                        os.remove(os.path.join(out_dir, name))
        Timers()["Postprocessing"].stop()


def unitary_test():
    """Unitary test for postprocessing.py script."""

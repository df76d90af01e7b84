"""Fan-out of independent solves and the adjoint-state gradient: batch.solve (the
task of the reference's process pool, _multiprocessing.py:72-153), batch.gradient (solver-level core
of Simulation.gradient, simulations.py:944-1095)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eb():
    import emg3d_b200
    from emg3d_b200 import _lib
    _lib.init()
    return emg3d_b200


def _setup(eb, n=16, mapping='Conductivity', aniso=False):
    rng = np.random.default_rng(2)
    h = [np.full(n, 50.0)] * 3
    grid = eb.TensorMesh(h, (-n * 25.0, -n * 25.0, -n * 25.0))
    sigma = 10 ** rng.uniform(-0.3, 0.3, grid.shape_cells)
    to_prop = {'Conductivity': lambda s: s, 'LgResistivity': lambda s: -np.log10(s)}[mapping]
    kw = dict(property_z=to_prop(1.5 * sigma)) if aniso else {}
    model = eb.Model(grid, to_prop(sigma), mapping=mapping, **kw)
    return grid, model


def test_batch_solve_dict_task(eb):
    """Both input formats; a model on a coarser grid is volume-averaged to the computational grid."""
    from emg3d_b200 import batch
    grid, model = _setup(eb)
    opts = dict(sslsolver=False, semicoarsening=False, linerelaxation=False, tol=1e-8, verb=-1)
    sfield = eb.get_source_field(grid, (10., 5., 0., 20., 10.), 1.0)
    e1, info1 = batch.solve({'model': model, 'sfield': sfield, 'efield': None, 'solver_opts': opts})
    e2, info2 = batch.solve({'model': model, 'grid': grid, 'source': (10., 5., 0., 20., 10.), 'frequency': 1.0,
                             'efield': None, 'solver_opts': opts})
    assert info1['exit'] == 0 and rel_err(e2.field, e1.field) == 0
    assert rel_err(e1.field, eb.solve(model, sfield, **opts).field) == 0
    # model on another grid
    hc = [np.full(8, 100.0)] * 3
    gc = eb.TensorMesh(hc, grid.origin)
    mc = eb.Model(gc, 0.5, mapping='Conductivity')
    e3, _ = batch.solve({'model': mc, 'sfield': sfield, 'efield': None, 'solver_opts': opts})
    e4 = eb.solve(eb.Model(grid, 0.5, mapping='Conductivity'), sfield, **opts)
    assert rel_err(e3.field, e4.field) < 1e-12


@pytest.mark.parametrize('mapping,aniso', [('Conductivity', False), ('LgResistivity', True)])
def test_gradient_against_finite_differences(eb, mapping, aniso):
    """The adjoint-state gradient against central finite differences of the misfit along a random
    model perturbation.  In general receiver sampling (interpolation along all three axes) and the
    adjoint source (a dipole spread bilinearly onto the four parallel edges of its cell) are not
    exact transposes of each other -- in the reference as here (its own test accepts 1.5 % in
    hand-picked cells, tests/test_simulations.py:822-876).  With linear sampling and x-directed
    receivers at the x-centres of their cells they are, and the gradient must match the finite
    differences to the accuracy of the difference quotient."""
    from emg3d_b200 import batch
    grid, model = _setup(eb, mapping=mapping, aniso=aniso)
    opts = dict(sslsolver=False, semicoarsening=False, linerelaxation=False, tol=1e-11, maxit=60, verb=-1,
                receiver_method='linear')
    sources = [(-100., 20., 10., 0., 0.), (120., -30., -20., 90., 0.)]
    rx = np.array([-225., -75., 75., 225.])
    receivers = (rx, 35.0, 15.0, 0.0, 0.0)
    rng = np.random.default_rng(5)
    truth = eb.Model(grid, model.property_x * 1.3, property_z=None if not aniso else model.property_z * 1.3,
                     mapping=mapping) if mapping == 'Conductivity' else \
        eb.Model(grid, model.property_x + 0.1, property_z=model.property_z + 0.1, mapping=mapping)
    _, _, observed = batch.gradient(truth, sources, 1.0, receivers, np.zeros((2, 4), complex), devices=[0], **opts)
    misfit, grad, syn = batch.gradient(model, sources, 1.0, receivers, observed, devices=[0], **opts)
    assert grad.shape == ((2 if aniso else 1), *grid.shape_cells) and misfit > 0
    # directional derivative along a smooth random perturbation of the inner cells
    d = np.zeros(grad.shape)
    d[:, 4:12, 4:12, 4:12] = rng.uniform(0.5, 1.0, (grad.shape[0], 8, 8, 8))
    eps = 1e-4 * np.abs(model.property_x).mean()

    def phi(sign):
        px = model.property_x + sign * eps * d[0]
        pz = model.property_z + sign * eps * d[1] if aniso else None
        m = eb.Model(grid, px, property_z=pz, mapping=mapping)
        return batch.gradient(m, sources, 1.0, receivers, observed, devices=[0], **opts)[0]

    fd = (phi(+1) - phi(-1)) / (2 * eps)
    ad = float(np.sum(grad * d))
    print(f"{mapping} aniso={aniso}: adjoint {ad:.6e}  finite differences {fd:.6e}  ratio {ad / fd:.4f}")
    assert abs(ad - fd) < 1e-3 * abs(fd)


def test_solve_many_on_every_visible_gpu(eb, golden):
    """solve_many with the default device list (all GPUs of the box): every device is used when
    there are at least as many sources, results equal single solves."""
    import ctypes
    from emg3d_b200 import _lib, batch
    n = ctypes.c_int(0)
    _lib.check(_lib.load().emg3d_b200_device_count(ctypes.byref(n)))
    grid, model = _setup(eb)
    opts = dict(sslsolver=False, semicoarsening=False, linerelaxation=False, tol=1e-8, verb=-1)
    srcs = [eb.get_source_field(grid, (20. * k - 60., 5., 0., 10. * k, 0.), 1.0) for k in range(2 * n.value)]
    many = eb.solve_many(model, srcs, **opts)
    for s, (e, _) in zip(srcs, many):
        assert rel_err(e.field, eb.solve(model, s, **opts).field) == 0
    devs = batch.process_map(batch._echo_device, list(range(4 * n.value)), devices=range(n.value))
    assert {d for _, d, _ in devs} == set(range(n.value))

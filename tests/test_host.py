"""CPU-only tests: host-side logic, golden host inputs, and the C-ABI surface.

No CUDA kernel is launched here (the build container has no GPU); what is
checked is everything around the kernels: the library exports every symbol the
header declares, the input builders agree with the reference, the parameter /
termination logic follows emg3d/solver.py, and the product fails loudly when no
device is present instead of falling back to the CPU.
"""
import os
import re

import numpy as np
import pytest

from conftest import REPO, rel_err


# --------------------------------------------------------------------------- #
# C ABI
# --------------------------------------------------------------------------- #

def _declared_symbols():
    with open(os.path.join(REPO, 'include', 'emg3d_b200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(emg3d_b200_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from emg3d_b200 import _lib
    if not os.path.exists(_lib.LIBPATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIBPATH)
    names = _declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/emg3d_b200.h but not exported"
    # and the Python binding declares a signature for each of them
    assert set(names) == set(_lib.EXPORTS)
    assert _lib.load().emg3d_b200_abi_version() == 1


def test_no_cpu_fallback():
    """Without a device every compute entry point raises; nothing runs on the host."""
    import ctypes
    import emg3d_b200 as eb
    from emg3d_b200 import _lib
    n = ctypes.c_int(0)
    rc = _lib.load().emg3d_b200_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    grid = eb.TensorMesh([np.ones(4) * 10] * 3, (0, 0, 0))
    model = eb.Model(grid, 1.0)
    sfield = eb.get_source_field(grid, (20, 20, 20, 0, 0), 1.0)
    with pytest.raises(eb.Emg3dB200Error):
        eb.solve(model, sfield)
    ex = np.zeros((4, 5, 5), order='F', dtype=complex)
    with pytest.raises(eb.Emg3dB200Error):
        eb.core.amat_x(ex, ex.reshape(5, 4, 5, order='F'), ex.reshape(5, 5, 4, order='F'),
                       ex, ex.reshape(5, 4, 5, order='F'), ex.reshape(5, 5, 4, order='F'),
                       np.ones((4, 4, 4), order='F', dtype=complex), np.ones((4, 4, 4), order='F', dtype=complex),
                       np.ones((4, 4, 4), order='F', dtype=complex), np.ones((4, 4, 4), order='F'),
                       np.ones(4), np.ones(4), np.ones(4))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; the package must never reference it."""
    pkg = os.path.join(REPO, 'emg3d_b200')
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(root, fn)) as f:
                    text = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), fn


# --------------------------------------------------------------------------- #
# host input builders against the reference (golden host.npz)
# --------------------------------------------------------------------------- #

def test_source_field_matches_reference(golden):
    import emg3d_b200 as eb
    gh = golden('host')
    grid = eb.TensorMesh([gh['hx'], gh['hy'], gh['hz']], gh['origin'])
    for k in range(int(gh['n_sources'])):
        for freq in (1.0, -2.5):
            sf = eb.get_source_field(grid, gh[f'src{k}'], freq)
            want = gh[f'src{k}_f{freq}']
            assert sf.field.dtype == want.dtype
            assert rel_err(sf.field, want) < 1e-12, (k, freq)


def test_source_assembly_beyond_electric_dipoles(golden):
    """Magnetic dipoles (square loops), finite dipoles, wires, complex strength, the frequency-
    independent vector, electric point sources and source OBJECTS with the reference's attributes
    (emg3d/fields.py:386-519, electrodes.py) against the reference's outputs."""
    import json
    import emg3d_b200 as eb
    gs = golden('sources')
    grid = eb.TensorMesh([gs['hx'], gs['hy'], gs['hz']], gs['origin'])

    def check(sf, want, what):
        assert sf.field.dtype == want.dtype, what
        assert rel_err(sf.field, want) < 1e-12, what

    for name in json.loads(str(gs['cases'])):
        kw = json.loads(str(gs[f'{name}_kwargs']))
        if isinstance(kw.get('strength'), list):
            kw['strength'] = complex(*kw['strength'])
        for freq in (1.0, -2.5, None):
            key = f'{name}_f{freq}'
            if key not in gs.files:
                continue
            sf = eb.get_source_field(grid, gs[f'{name}_source'], freq, **kw)
            idx, val, bg = sf.sparse
            dense = np.full(grid.n_edges, bg, dtype=val.dtype)
            dense[idx] = val
            assert np.array_equal(dense, sf.field)
            assert idx.size < 400                          # a handful of edges, not a dense array
            check(sf, gs[key], key)
    for name in json.loads(str(gs['objects'])):
        cls = type(str(gs[f'{name}_class']), (), {})       # same class name, same attributes
        obj = cls()
        obj.points, obj.coordinates = gs[f'{name}_points'], gs[f'{name}_coordinates']
        obj.strength = float(gs[f'{name}_strength'])
        for freq in (1.0, -2.5, None):
            check(eb.get_source_field(grid, obj, freq), gs[f'{name}_f{freq}'], (name, freq))
    # argument checks of the reference
    with pytest.raises(ValueError, match='Coordinates are wrong defined'):
        eb.get_source_field(grid, (1., 2., 3., 4.), 1.0)
    with pytest.raises(ValueError, match='The two electrodes are identical'):
        eb.get_source_field(grid, (1., 1., 2., 2., 3., 3.), 1.0)
    with pytest.raises(ValueError, match='outside grid'):
        eb.get_source_field(grid, (1e6, 0., 0., 0., 0.), 1.0)
    pt = type('TxElectricPoint', (), {})()
    pt.points, pt.coordinates, pt.strength = np.zeros((1, 3)), (1e6, 0., 0., 0., 0.), 1.0
    with pytest.raises(ValueError, match='outside grid'):
        eb.get_source_field(grid, pt, 1.0)
    mp = type('TxMagneticPoint', (), {})()
    mp.points, mp.coordinates, mp.strength = np.zeros((1, 3)), (0., 0., 0., 0., 0.), 1.0
    with pytest.raises(NotImplementedError, match='magnetic point'):
        eb.get_source_field(grid, mp, 1.0)


def test_source_field_of_bench_configs(golden):
    """The source vectors of the golden solves are reproduced by our builder."""
    import emg3d_b200 as eb
    from helpers import solve_case
    gs = golden('solves')
    for prefix in ('res_F_', 'lap_F_', 'config1_', 'config3_', 'config4_'):
        c = solve_case(gs, prefix)
        grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], c['origin'])
        sf = eb.get_source_field(grid, c['source'], c['frequency'])
        assert rel_err(sf.field, c['sfield']) < 1e-12, prefix


def test_volume_model_matches_reference(golden):
    import emg3d_b200 as eb
    gh = golden('host')
    grid = eb.TensorMesh([gh['hx'], gh['hy'], gh['hz']], gh['origin'])
    props = {k: gh['vm_' + k] for k in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r')}
    cases = {'iso': ['property_x'], 'vti': ['property_x', 'property_z'],
             'hti': ['property_x', 'property_y'],
             'tri': ['property_x', 'property_y', 'property_z'], 'full': list(props)}
    names = {'iso': 'isotropic', 'vti': 'VTI', 'hti': 'HTI', 'tri': 'triaxial', 'full': 'triaxial'}
    for case, keys in cases.items():
        model = eb.Model(grid, **{k: props[k] for k in keys})
        assert model.case == names[case]
        for freq in (0.7, -3.0):
            vm = eb.VolumeModel(model, eb.Field(grid, frequency=freq))
            for n in ('eta_x', 'eta_y', 'eta_z', 'zeta'):
                want = gh[f'vm_{case}_f{freq}_{n}']
                got = getattr(vm, n)
                assert got.dtype == want.dtype and got.shape == want.shape
                assert rel_err(got, want) < 1e-15, (case, freq, n)
            # aliasing contract (emg3d/models.py:698-712)
            assert (vm.eta_y is vm.eta_x) == (case in ('iso', 'vti'))
            assert (vm.eta_z is vm.eta_x) == (case in ('iso', 'hti'))


def test_field_container():
    import emg3d_b200 as eb
    grid = eb.TensorMesh([np.ones(3), np.ones(4), np.ones(5)], (0, 0, 0))
    f = eb.Field(grid, frequency=2.0)
    assert f.field.dtype == np.complex128 and f.field.size == grid.n_edges
    assert f.fx.shape == (3, 5, 6) and f.fy.shape == (4, 4, 6) and f.fz.shape == (4, 5, 5)
    f.fy[1, 2, 3] = 7
    assert f.field[grid.n_edges_x + 1 + 4 * (2 + 4 * 3)] == 7       # x fastest
    assert eb.Field(grid, frequency=-2.0).field.dtype == np.float64
    assert abs(f.sval - 2j * np.pi * 2.0) == 0 and eb.Field(grid, frequency=-2.0).sval == 2.0
    with pytest.raises(ValueError, match='`frequency` must be'):
        eb.Field(grid, frequency=0.0)
    h = eb.Field(grid, frequency=2.0, electric=False)              # magnetic: on the faces
    assert h.field.size == grid.n_faces == 4 * 4 * 5 + 3 * 5 * 5 + 3 * 4 * 6
    assert h.fx.shape == (4, 4, 5) and h.fy.shape == (3, 5, 5) and h.fz.shape == (3, 4, 6)
    assert 'magnetic' in repr(h) and not h.copy().electric and h != f
    g = f.copy()
    assert g == f
    g.field[0] = 1
    assert not (g == f)


# --------------------------------------------------------------------------- #
# transfer-operator tables against the reference (golden transfer.npz)
# --------------------------------------------------------------------------- #

def test_restrict_weights_and_interpolation_tables(golden):
    import emg3d_b200 as eb
    from emg3d_b200 import core, solver
    gt = golden('transfer')
    for k in range(int(gt['n_cases'])):
        p = f"t{k}_"
        grid = eb.TensorMesh([gt[p + 'hx'], gt[p + 'hy'], gt[p + 'hz']], gt[p + 'origin'])
        for sc in gt[p + 'sc_dirs']:
            q = p + f"sc{sc}_"
            fl = core.SC_FLAGS[int(sc)]
            ch = [np.diff(getattr(grid, 'nodes_' + n)[::2 if f else 1]) for n, f in zip('xyz', fl)]
            cgrid = eb.BaseMesh(ch, grid.origin)
            w = solver._get_restriction_weights(grid, cgrid, int(sc))
            for a, ax in enumerate('xyz'):
                np.testing.assert_allclose(np.array(w[a]), gt[q + f'w{ax}'], rtol=1e-13, atol=1e-15)
    # literal values of the reference's unit test (tests/test_core.py:459-478)
    edges = np.array([0., 500, 1200, 2000, 3000])
    width = (edges[1:] - edges[:-1])
    centr = edges[:-1] + width / 2
    c_edges, c_width = edges[::2], None
    c_width = c_edges[1:] - c_edges[:-1]
    c_centr = c_edges[:-1] + c_width / 2
    wl, w0, wr = core.restrict_weights(edges, centr, width, c_edges, c_centr, c_width)
    np.testing.assert_allclose(wl, [350 / 250, 250 / 600, 400 / 900])
    np.testing.assert_allclose(w0, [1, 1, 1])
    np.testing.assert_allclose(wr, [350 / 600, 500 / 900, 400 / 500])


def test_regular_grid_prolongator():
    """Against SciPy's interpolator, like tests/test_solver.py:780-840."""
    from scipy.interpolate import RegularGridInterpolator
    from emg3d_b200.solver import RegularGridProlongator
    rng = np.random.default_rng(4)
    x = np.cumsum(np.r_[0., rng.uniform(1, 3, 12)])
    y = np.cumsum(np.r_[-5., rng.uniform(1, 3, 8)])
    cx, cy = x[::2], y[::2]
    vals = rng.standard_normal((cx.size, cy.size)) + 1j * rng.standard_normal((cx.size, cy.size))
    fn = RegularGridProlongator(cx, cy, x, y)
    got = fn(vals).reshape((x.size, y.size), order='F')
    ref = RegularGridInterpolator((cx, cy), vals, bounds_error=False, fill_value=None)
    xx, yy = np.meshgrid(x, y, indexing='ij')
    want = ref(np.c_[xx.ravel(), yy.ravel()]).reshape(xx.shape)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)
    assert fn.size == x.size * y.size


def test_blocks_to_amat_known_pattern():
    """The reference's known-answer pattern 1..90 (tests/test_core.py:149-193)."""
    from emg3d_b200 import core
    n = 3
    amat = np.zeros(6 * (5 * n - 4))
    bvec = np.zeros(5 * n - 4)
    for im in range(n):
        middle = np.arange(25.) + 100 * (im + 1)
        left = np.arange(25.) + 1000 * (im + 1)
        rhs = np.arange(5.) + 10 * (im + 1)
        core.blocks_to_amat(amat, bvec, middle, left, rhs, im, n)
    # dense reconstruction: A(p, q) = amat[p + 5 q]
    size = 5 * n - 4
    dense = np.zeros((size, size))
    for q in range(size):
        for p in range(q, min(size, q + 6)):
            dense[p, q] = amat[p + 5 * q]
    assert dense[0, 0] == 100 and dense[4, 0] == 104 and dense[4, 4] == 124   # first middle
    assert dense[5, 5] == 200 and dense[9, 6] == 200 + 4 + 5                    # second middle
    assert dense[5, 1] == 2000 + 5 and dense[5, 4] == 2000 + 20                 # left, first row
    assert dense[6, 1] == 2000 + 6 and dense[9, 4] == 2000 + 24                 # left diagonal
    assert dense[10, 10] == 300 and dense[10, 6] == 3000 + 5                    # last block
    np.testing.assert_array_equal(bvec, [10, 11, 12, 13, 14, 20, 21, 22, 23, 24, 30])


# --------------------------------------------------------------------------- #
# parameters, direction tables, termination (emg3d/solver.py:1074-1664)
# --------------------------------------------------------------------------- #

def test_mgparameters():
    from emg3d_b200.solver import MGParameters
    var = MGParameters(verb=0, sslsolver=True, semicoarsening=True, linerelaxation=True,
                       shape_cells=(256, 256, 256))
    assert var.sslsolver == 'bicgstab' and var.cycmax == 2 and var.maxcycle == 3
    assert var.ssl_maxit == 50 and var.maxit == 3 and var._repr_maxit == '50 (3)'
    assert (var.sc_dir, var.lr_dir) == (1, 4) and next(var.sc_cycle) == 2 and next(var.lr_cycle) == 5
    assert list(var.clevel) == [7, 7, 7, 7]
    var = MGParameters(verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
                       shape_cells=(512, 512, 256), cycle='V')
    assert list(var.clevel) == [8, 8, 8, 8] and var.cycmax == 1 and var.maxit == 50
    assert var.sc_cycle is False and var.lr_cycle is False and (var.sc_dir, var.lr_dir) == (0, 0)
    var = MGParameters(verb=0, sslsolver=False, semicoarsening=1213, linerelaxation=456,
                       shape_cells=(8, 24, 4), clevel=1)
    assert list(var.raw_sc_cycle) == [1, 2, 1, 3] and var.maxcycle == 4
    assert list(var.clevel) == [1, 1, 1, 1]
    var = MGParameters(verb=0, sslsolver=False, semicoarsening=3, linerelaxation=7,
                       shape_cells=(12, 20, 28))
    assert list(var.clevel) == [2, 2, 2, 2] and var._repr_clevel['shape_cells'] == (3, 5, 7)
    assert var._repr_clevel['message'] == "  :: Grid not optimal for MG solver ::"
    assert "   Coarsest grid  :   3 x   5 x   7     => 105 cells\n" in repr(var)
    assert "   semicoarsening : True [3]  " in repr(var)
    # the reference's messages (tests/test_solver.py:685-737)
    for bad, msg in ((dict(semicoarsening=5), '`semicoarsening` must be one o'),
                     (dict(linerelaxation=-9), '`linerelaxation` must be one o'),
                     (dict(sslsolver='jacobi'), '`sslsolver` must be True'),
                     (dict(sslsolver=4), '`sslsolver` must be True'),
                     (dict(cycle='G'), '`cycle` must be one of'),
                     (dict(cycle=None, sslsolver=False), 'At least `cycle` or `sslsolve'),
                     (dict(shape_cells=(1, 2, 2)), 'Nr. of cells must be at least')):
        kw = dict(verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
                  shape_cells=(8, 8, 8))
        kw.update(bad)
        with pytest.raises(ValueError, match=msg):
            MGParameters(**kw)


def test_direction_tables():
    from emg3d_b200 import solver, meshes

    def g(nx, ny, nz):
        return meshes.BaseMesh([np.ones(nx), np.ones(ny), np.ones(nz)], (0, 0, 0))
    # tests/test_solver.py:843-900
    assert solver._current_sc_dir(0, g(8, 8, 8)) == 0
    assert solver._current_sc_dir(1, g(8, 8, 8)) == 1
    assert solver._current_sc_dir(0, g(2, 8, 8)) == 1
    assert solver._current_sc_dir(2, g(2, 8, 8)) == 6
    assert solver._current_sc_dir(3, g(2, 8, 8)) == 5
    assert solver._current_sc_dir(0, g(8, 3, 2)) == 4
    assert solver._current_sc_dir(0, g(2, 2, 2)) == 6
    table = {(2, 8, 8): {1: 0, 5: 3, 6: 2, 7: 4, 2: 2}, (8, 2, 8): {2: 0, 4: 3, 6: 1, 7: 5},
             (8, 8, 2): {3: 0, 4: 2, 5: 1, 7: 6}, (2, 2, 8): {7: 3, 6: 0, 4: 3},
             (8, 8, 8): {k: k for k in range(8)}}
    for shape, tab in table.items():
        for lr, want in tab.items():
            assert int(solver._current_lr_dir(lr, g(*shape))) == want, (shape, lr)


def test_terminate(capsys):
    from emg3d_b200 import solver

    class Var:
        def __init__(self, **kw):
            self.tol, self.l2_refe, self.maxit, self.sslsolver, self.verb = 1e-6, 1.0, 5, False, 3
            self.exit_message = ''
            self.__dict__.update(kw)

        def cprint(self, info, verbosity, **kw):
            if self.verb > verbosity:
                print(info)
    v = Var()
    assert solver._terminate(v, 1e-7, 1.0, 1) and v.exit_message == 'CONVERGED'
    v = Var()
    assert solver._terminate(v, 11.0, 1.0, 1) and v.exit_message == 'DIVERGED'
    v = Var()
    assert solver._terminate(v, np.nan, 1.0, 1) and v.exit_message == 'DIVERGED'
    v = Var()
    assert solver._terminate(v, 0.5, 0.4, 3) and v.exit_message == 'STAGNATED'
    v = Var()
    assert not solver._terminate(v, 0.5, 0.4, 2) and v.exit_message == ''
    v = Var()
    assert solver._terminate(v, 0.5, 0.6, 5) and v.exit_message == 'MAX. ITERATION REACHED, NOT CONVERGED'
    out, _ = capsys.readouterr()
    assert '   > CONVERGED' in out and '   > STAGNATED' in out
    v = Var(sslsolver='bicgstab')
    with pytest.raises(solver._ConvergenceError):
        solver._terminate(v, 11.0, 1.0, 1)
    v = Var(sslsolver='bicgstab')
    assert solver._terminate(v, 0.5, 0.6, 5) and v.exit_message == ''


def test_log_helpers(capsys):
    """Format of the per-cycle line and the one-liner (tests/test_solver.py:1037-1127)."""
    from emg3d_b200 import solver
    var = solver.MGParameters(verb=4, sslsolver=False, semicoarsening=False, linerelaxation=False,
                              shape_cells=(8, 8, 8))
    var.l2_refe = 1e-3
    var.level_all = [0, 1, 2, 1, 2, 1, 0]
    var.it = 1
    solver._print_cycle_info(var, 3.399e-5, 1e-3)
    out, _ = capsys.readouterr()
    assert "       h_\n      2h_ \\    /\n      4h_  \\/\\/ \n\n" in out
    assert "   3.399e-02  after   1 F-cycles   [3.399e-05, 0.034]   0 0" in out
    var.verb = 1
    var.exit_message = 'CONVERGED'
    solver._print_one_liner(var, 2e-9, True)
    out, _ = capsys.readouterr()
    assert out.startswith(":: emg3d :: 2.0e-06; 1; 0:00:0") and out.rstrip().endswith("; CONVERGED")


def test_bench_work_count():
    """W of SURVEY.md section 8(d)."""
    import bench
    assert bench.vcycle_work((256, 256, 256)) == 76695816
    assert bench.vcycle_work((128, 128, 128)) == 9586952
    assert bench.vcycle_work((32, 32, 32)) == 149768
    assert bench.vcycle_work((512, 512, 512)) == 613566728


def test_process_map_binds_one_device_per_worker():
    """batch.process_map: the reference's process-pool fan-out (_multiprocessing.py:33-65)
    with every worker bound to one GPU before its first task; order of results kept."""
    from emg3d_b200 import batch
    out = batch.process_map(batch._echo_device, list(range(6)), devices=[3, 5], payload='p')
    assert [x for x, _, _ in out] == list(range(6))
    assert {d for _, d, _ in out} <= {3, 5} and all(str(d) == env for _, d, env in out)
    with pytest.raises(ValueError, match='at least one GPU'):
        batch.process_map(batch._echo_device, [1], devices=[])


def test_bench_work_accounting_and_recorded_traffic():
    """bench.py: W = sum over levels of cells x sweeps for the plain V(2,2) cycle
    (76 695 816 at 256^3, 9 586 952 at 128^3, 149 768 at 32^3) and the
    DRAM traffic of the roofline kernel parsed from the committed ncu summary."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        'bench', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.vcycle_work((256, 256, 256)) == 76695816
    assert bench.vcycle_work((128, 128, 128)) == 9586952
    assert bench.vcycle_work((32, 32, 32)) == 149768
    t = bench.recorded_traffic('gs_point_tile_kernel')
    assert t is not None and 3.0e8 < t < 1.5e9            # algorithmic: 3.86e8 bytes per launch
    assert bench.recorded_traffic('no_such_kernel') is None


def test_workspace_digest_sees_every_element_and_trusts_frozen_arrays():
    """Workspace._digest (the cache key of device-resident coefficients): an in-place edit of ANY
    element changes it (a strided sample missed local edits); a read-only array is
    identified by its address, without the checksum pass."""
    from emg3d_b200.solver import Workspace
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.uniform(1, 2, (33, 17, 9)))
    d0 = Workspace._digest(a)
    assert Workspace._digest(a.copy(order='F')) == d0
    for idx in [(0, 0, 0), (32, 16, 8), (5, 3, 7), (17, 0, 4)]:
        b = a.copy(order='F')
        b[idx] *= 1.0 + 1e-15
        assert Workspace._digest(b) != d0, idx
    a.flags.writeable = False
    frozen = Workspace._digest(a)
    assert frozen != d0 and frozen == Workspace._digest(a)
    assert frozen[1] == a.__array_interface__['data'][0]

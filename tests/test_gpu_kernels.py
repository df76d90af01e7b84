"""Parity of the CUDA kernels (through the C ABI) with the reference / oracle.

All comparisons are norm-wise relative errors ``||x_gpu - x_ref|| / ||x_ref||``.
Bars (fp64 throughout):

* ``amat_x`` / residual, restriction, prolongation, cell sums: <= 1e-13
  (pure stencils, no solves);
* smoothers in ``order='lex'`` against the reference's own outputs (golden) and
  against the oracle on larger grids, after a call: point smoother <= 5e-12 (10 x the measured
  4e-13; every
  update is a 6x6 solve; the reference itself is only defined to ~1e-13 because
  numba runs with fastmath); line smoothers <= 2e-10: the CUDA kernels eliminate
  the line edges before the transverse edges (4x4 blocks, csrc/gs_line.cu) while
  the reference and the oracle factorise 5x5 blocks in natural order, and the line
  systems are ill-conditioned (the curl-curl part is singular on gradients and only
  the small eta term regularises it), so the two elimination orders differ by
  cond x eps: measured 5e-14 .. 7e-11 on these cases, worst in the Laplace domain;
  the reference's own ``Field.__eq__`` uses rtol = 1e-10 (fields.py:135);
* smoothers in ``order='color'`` against the oracle run in the same colour
  sequence: same bars.
"""
import numpy as np
import pytest

import oracle
from oracle import mg
from conftest import rel_err
from helpers import hfield_case, kernel_case, split_faces, split_field

pytestmark = pytest.mark.gpu

GS = ['gauss_seidel', 'gauss_seidel_x', 'gauss_seidel_y', 'gauss_seidel_z']
TOL_GS = [5e-12, 2e-10, 2e-10, 2e-10]          # point, x-, y-, z-lines (see above)


@pytest.fixture(scope='module')
def core():
    from emg3d_b200 import core, _lib
    _lib.init()
    return core


def _margs(c):
    return (c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'], c['hx'], c['hy'], c['hz'])


def random_case(rng, shape, cplx, aliased=False):
    """Random stretched grid, triaxial eta with real part (eps_r), mu_r != 1."""
    nx, ny, nz = shape
    h = [50 * 1.1 ** rng.uniform(-3, 3, nx), 60 * 1.2 ** rng.uniform(-3, 3, ny),
         40 * 1.15 ** rng.uniform(-3, 3, nz)]
    g = mg.Grid(h)
    res = 10 ** rng.uniform(-0.5, 1.5, shape)
    vm = mg.VolumeModel(g, res, None if aliased else 1.5 * res * rng.uniform(.5, 2, shape),
                        None if aliased else 3 * res, rng.uniform(1, 2, shape),
                        rng.uniform(1, 10, shape), 1.3 if cplx else -1.3)
    n = g.n_edges
    s = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    e = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    ex, ey, ez = g.split(e)
    ex[:, 0, :] = ex[:, -1, :] = 0
    ex[:, :, 0] = ex[:, :, -1] = 0
    ey[0, :, :] = ey[-1, :, :] = 0
    ey[:, :, 0] = ey[:, :, -1] = 0
    ez[0, :, :] = ez[-1, :, :] = 0
    ez[:, 0, :] = ez[:, -1, :] = 0
    d = dict(shape=shape, hx=h[0], hy=h[1], hz=h[2], eta_x=vm.eta_x, eta_y=vm.eta_y,
             eta_z=vm.eta_z, zeta=vm.zeta, s=s, e=e, grid=g, vm=vm)
    return d


def test_amat_x_golden(core, golden):
    gk = golden('kernels')
    for k in range(int(gk['n_cases'])):
        c = kernel_case(gk, k)
        r = c['s'].copy()
        core.amat_x(*split_field(c['shape'], r), *split_field(c['shape'], c['e']), *_margs(c))
        assert rel_err(r, c['r']) < 1e-13, k


def test_amat_x_random_boundaries(core):
    """Non-zero boundary values of e and s exercise the reference's boundary rules."""
    rng = np.random.default_rng(11)
    for shape, cplx in (((9, 7, 5), True), ((33, 6, 10), False), ((40, 37, 35), True),
                        ((70, 66, 65), True), ((65, 97, 50), False)):   # plane-streaming kernel
        c = random_case(rng, shape, cplx)
        e = rng.standard_normal(c['e'].size) + (1j * rng.standard_normal(c['e'].size) if cplx else 0)
        r1, r2 = c['s'].copy(), c['s'].copy()
        oracle.amat_x(*split_field(shape, r1), *split_field(shape, e), *_margs(c))
        core.amat_x(*split_field(shape, r2), *split_field(shape, e), *_margs(c))
        assert rel_err(r2, r1) < 1e-13


@pytest.mark.parametrize('ldir', [0, 1, 2, 3])
def test_gauss_seidel_lex_golden(core, golden, ldir):
    gk = golden('kernels')
    fn = getattr(core, GS[ldir])
    for k in range(int(gk['n_cases'])):
        c = kernel_case(gk, k)
        for nu in (1, 2):
            e = c['e'].copy()
            fn(*split_field(c['shape'], e), *split_field(c['shape'], c['s']), *_margs(c), nu,
               order='lex')
            assert rel_err(e, gk[c['prefix'] + f'gs{ldir}_nu{nu}']) < TOL_GS[ldir], (k, nu)


# shapes chosen so that both the single-block path (small grids) and the
# multi-launch path (> 512 interior nodes / > 1024 lines) are exercised
BIG = [((24, 20, 18), True), ((6, 40, 36), True), ((38, 5, 37), False), ((36, 35, 4), True),
       ((12, 11, 9), False)]


def _tile_variant():
    import ctypes
    from emg3d_b200 import _lib
    v = ctypes.c_int(0)
    _lib.check(_lib.load().emg3d_b200_point_tile_schedule(ctypes.byref(v)))
    return v.value


def _tile_shape():
    import ctypes
    from emg3d_b200 import _lib
    t = (ctypes.c_int * 3)()
    _lib.check(_lib.load().emg3d_b200_point_tile_shape(t))
    return tuple(t)


@pytest.mark.parametrize('ldir', [0, 1, 2, 3])
@pytest.mark.parametrize('order', ['lex', 'color'])
def test_gauss_seidel_vs_oracle(core, ldir, order):
    rng = np.random.default_rng(100 + ldir)
    fn, ofn = getattr(core, GS[ldir]), getattr(oracle, GS[ldir])
    cases = BIG + ([((70, 68, 66), True)] if ldir == 0 else [])   # tile-fused schedule
    for shape, cplx in cases:
        c = random_case(rng, shape, cplx, aliased=(shape[0] == 12))
        for nu in (1, 3):
            e_gpu, e_cpu = c['e'].copy(), c['e'].copy()
            fn(*split_field(shape, e_gpu), *split_field(shape, c['s']), *_margs(c), nu, order=order)
            if order == 'lex':
                ofn(*split_field(shape, e_cpu), *split_field(shape, c['s']), *_margs(c), nu)
            else:
                seq = oracle.color_sequence(ldir, shape, nu, tile_variant=_tile_variant(),
                                            tile=_tile_shape())
                oracle.gs_sequence(ldir, *split_field(shape, e_cpu), *split_field(shape, c['s']),
                                   *_margs(c), seq)
            assert rel_err(e_gpu, e_cpu) < TOL_GS[ldir], (shape, nu)


# long lines: the segment-parallel kernels (csrc/gs_line_seg.cu: one warp per line, 8 blocks per
# lane, jump-matrix scans) take over for 66 .. 257 cells along the line.  Shapes are given for
# x-lines and permuted for y / z; they cover 16 and 32 lanes per line, full and partial last
# segments, one block in the last segment, odd line counts (padding line of a two-line warp).
LONG = [((130, 7, 6), True), ((257, 4, 5), True), ((100, 6, 5), False), ((66, 5, 4), True),
        ((129, 3, 4), False), ((200, 4, 3), True)]


def _perm(shape, ldir):
    n, a, b = shape
    return {1: (n, a, b), 2: (a, n, b), 3: (a, b, n)}[ldir]


@pytest.mark.parametrize('ldir', [1, 2, 3])
def test_long_lines_segment_kernel_vs_oracle(core, ldir):
    """Multicolour order on long lines: segment-parallel kernel == oracle in the same colour
    sequence == the one-thread-per-line kernel, with PEC and with non-zero boundary data."""
    from emg3d_b200 import _lib
    rng = np.random.default_rng(300 + ldir)
    fn = getattr(core, GS[ldir])
    prev = _lib.line_seg_mask(7)
    try:
        for shape0, cplx in LONG:
            shape = _perm(shape0, ldir)
            c = random_case(rng, shape, cplx)
            for nonzero_bc in (False, True):
                e0 = c['e'].copy()
                if nonzero_bc:
                    e0 = rng.standard_normal(e0.size) + (1j * rng.standard_normal(e0.size) if cplx else 0)
                for nu in (1, 2):
                    e_seg, e_thr, e_cpu = e0.copy(), e0.copy(), e0.copy()
                    _lib.line_seg_mask(7)
                    fn(*split_field(shape, e_seg), *split_field(shape, c['s']), *_margs(c), nu, order='color')
                    _lib.line_seg_mask(0)
                    fn(*split_field(shape, e_thr), *split_field(shape, c['s']), *_margs(c), nu, order='color')
                    assert rel_err(e_seg, e_thr) < 1e-11, (shape, nonzero_bc, nu, 'seg vs thread kernel')
                    if nonzero_bc:      # the oracle sweeps assume PEC data on the side faces
                        continue
                    seq = oracle.color_sequence(ldir, shape, nu)
                    oracle.gs_sequence(ldir, *split_field(shape, e_cpu), *split_field(shape, c['s']),
                                       *_margs(c), seq)
                    assert rel_err(e_seg, e_cpu) < TOL_GS[ldir], (shape, nonzero_bc, nu, 'seg vs oracle')
    finally:
        _lib.line_seg_mask(prev)


@pytest.mark.parametrize('order', ['lex', 'color'])
def test_single_block_exactness(core, order):
    """On a grid with one block a sweep solves the system exactly (residual ~ 0)."""
    rng = np.random.default_rng(7)
    for ldir, shape in ((0, (2, 2, 2)), (1, (7, 2, 2)), (2, (2, 6, 2)), (3, (2, 2, 5))):
        c = random_case(rng, shape, True)
        e = np.zeros_like(c['e'])
        getattr(core, GS[ldir])(*split_field(shape, e), *split_field(shape, c['s']), *_margs(c), 1,
                                order=order)
        r = c['s'].copy()
        core.amat_x(*split_field(shape, r), *split_field(shape, e), *_margs(c))
        # interior residual only: boundary edges keep r = s
        rx, ry, rz = split_field(shape, r)
        inner = np.r_[rx[:, 1:-1, 1:-1].ravel(), ry[1:-1, :, 1:-1].ravel(), rz[1:-1, 1:-1, :].ravel()]
        assert np.linalg.norm(inner) < 1e-12 * np.linalg.norm(c['s'])


@pytest.mark.parametrize('order', ['lex', 'color'])
def test_single_block_exactness_nonzero_boundary(core, order):
    """Same with non-zero values on ALL boundary edges, including the planes at the
    ends of a line: the boundary planes of a z-slab are halo data of the neighbouring
    GPU, not PEC zeros, and every smoother must treat them as Dirichlet data."""
    rng = np.random.default_rng(17)
    for ldir, shape in ((0, (2, 2, 2)), (1, (7, 2, 2)), (2, (2, 6, 2)), (3, (2, 2, 5))):
        c = random_case(rng, shape, True)
        n = c['e'].size
        e = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        getattr(core, GS[ldir])(*split_field(shape, e), *split_field(shape, c['s']), *_margs(c), 1,
                                order=order)
        r = c['s'].copy()
        core.amat_x(*split_field(shape, r), *split_field(shape, e), *_margs(c))
        rx, ry, rz = split_field(shape, r)
        inner = np.r_[rx[:, 1:-1, 1:-1].ravel(), ry[1:-1, :, 1:-1].ravel(), rz[1:-1, 1:-1, :].ravel()]
        assert np.linalg.norm(inner) < 1e-11 * np.linalg.norm(c['s']), ldir


def test_line_equals_point_on_degenerate_grids(core):
    """With a single interior node along the line a line sweep is a point sweep
    (the reference's own test of the line smoothers, tests/test_core.py:88-139)."""
    rng = np.random.default_rng(8)
    for ldir, shape in ((1, (2, 8, 8)), (2, (8, 2, 8)), (3, (8, 8, 2))):
        c = random_case(rng, shape, True)
        e0, e1 = c['e'].copy(), c['e'].copy()
        core.gauss_seidel(*split_field(shape, e0), *split_field(shape, c['s']), *_margs(c), 2,
                          order='lex')
        getattr(core, GS[ldir])(*split_field(shape, e1), *split_field(shape, c['s']), *_margs(c), 2,
                                order='lex')
        assert rel_err(e1, e0) < 1e-10


def test_determinism(core):
    rng = np.random.default_rng(9)
    c = random_case(rng, (24, 20, 18), True)
    outs = []
    for _ in range(2):
        e = c['e'].copy()
        core.gauss_seidel(*split_field(c['shape'], e), *split_field(c['shape'], c['s']), *_margs(c),
                          2, order='color')
        core.gauss_seidel_y(*split_field(c['shape'], e), *split_field(c['shape'], c['s']), *_margs(c),
                            2, order='color')
        outs.append(e)
    assert np.array_equal(outs[0], outs[1])


def test_band_solve(core):
    rng = np.random.default_rng(5)
    for n, cplx in ((6, False), (6, True), (41, True), (1, True), (3, False), (636, True)):
        A = np.zeros((n, n), dtype=complex if cplx else float)
        for i in range(n):
            for j in range(max(0, i - 5), i + 1):
                A[i, j] = A[j, i] = rng.standard_normal() + (1j * rng.standard_normal() if cplx else 0)
            A[i, i] += 10
        b = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
        amat = np.zeros(6 * n, dtype=A.dtype)
        for j in range(n):
            for i in range(j, min(n, j + 6)):
                amat[i + 5 * j] = A[i, j]
        a_o, x_o = amat.copy(), b.copy()
        oracle.solve(a_o, x_o)
        x = b.copy()
        core.solve(amat, x)
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-10)
        np.testing.assert_allclose(x, x_o, rtol=1e-10)
        np.testing.assert_allclose(amat, a_o, rtol=1e-10, atol=1e-14)


def test_restrict_golden(core, golden):
    gt = golden('transfer')
    for k in range(int(gt['n_cases'])):
        p = f"t{k}_"
        shape = tuple(gt[p + n].size for n in ('hx', 'hy', 'hz'))
        for sc in gt[p + 'sc_dirs']:
            q = p + f"sc{sc}_"
            fl = core.SC_FLAGS[int(sc)]
            cshape = tuple(n // 2 if f else n for n, f in zip(shape, fl))
            cs = np.zeros_like(gt[q + 'cs'])
            w = [tuple(gt[q + f'w{a}']) for a in 'xyz']
            core.restrict(*split_field(cshape, cs), *split_field(shape, gt[p + 'r'].copy()), *w, sc)
            assert rel_err(cs, gt[q + 'cs']) < 1e-13, (k, sc)
    # argument checks of the host-array entry (coarse arrays of the wrong shape, bad sc_dir)
    fine = split_field(shape, gt[p + 'r'].copy())
    with pytest.raises(ValueError, match='expected'):
        core.restrict(*split_field(shape, np.zeros_like(gt[p + 'r'])), *fine, *w, 0)
    with pytest.raises(ValueError, match='sc_dir'):
        core.restrict(*split_field(cshape, cs), *fine, *w, 7)


def test_solver_wrappers_golden(golden):
    """restriction / prolongation / residual / smoothing wrappers on host containers."""
    import emg3d_b200 as eb
    from emg3d_b200 import solver
    gt = golden('transfer')
    for k in range(int(gt['n_cases'])):
        p = f"t{k}_"
        grid = eb.TensorMesh([gt[p + 'hx'], gt[p + 'hy'], gt[p + 'hz']], gt[p + 'origin'])
        dt = gt[p + 'r'].dtype
        freq = 1.3 if dt.kind == 'c' else -1.3

        class VM:
            pass
        vm = VM()
        vm.grid, vm.case = grid, 'triaxial'
        vm.eta_x, vm.eta_y, vm.eta_z, vm.zeta = (np.asfortranarray(gt[p + n]) for n in
                                                 ('eta_x', 'eta_y', 'eta_z', 'zeta'))
        sfield = eb.Field(grid, gt[p + 'e'].copy(), frequency=freq)
        rfield = eb.Field(grid, gt[p + 'r'].copy(), frequency=freq)
        for sc in gt[p + 'sc_dirs']:
            q = p + f"sc{sc}_"
            cm, cs, ce = solver.restriction(vm, sfield, rfield, int(sc))
            assert rel_err(cs.field, gt[q + 'cs']) < 1e-13
            assert np.all(ce.field == 0)
            for n in ('eta_x', 'eta_y', 'eta_z', 'zeta'):
                assert rel_err(getattr(cm, n), gt[q + 'c' + n]) < 1e-15, (k, sc, n)
            assert cm.grid.shape_cells == tuple(
                m // 2 if f else m for m, f in zip(grid.shape_cells, core_flags(sc)))
            ce.field[:] = gt[q + 'ce']
            e1 = sfield.copy()
            solver.prolongation(e1, ce, int(sc))
            assert rel_err(e1.field, gt[q + 'e_out']) < 1e-14, (k, sc)
            np.testing.assert_allclose(
                solver._restrict_model_parameters(vm.zeta, int(sc)), gt[q + 'czeta'], rtol=1e-15)


def core_flags(sc):
    from emg3d_b200 import core
    return core.SC_FLAGS[int(sc)]


def test_residual_and_smoothing_wrappers(golden):
    import emg3d_b200 as eb
    from emg3d_b200 import solver
    gk = golden('kernels')
    for k in range(int(gk['n_cases'])):
        c = kernel_case(gk, k)
        grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], (0, 0, 0))
        freq = 1.3 if c['s'].dtype.kind == 'c' else -1.3

        class VM:
            pass
        vm = VM()
        vm.grid, vm.case = grid, 'triaxial'
        vm.eta_x, vm.eta_y, vm.eta_z, vm.zeta = c['eta_x'], c['eta_y'], c['eta_z'], c['zeta']
        s = eb.Field(grid, c['s'].copy(), frequency=freq)
        e = eb.Field(grid, c['e'].copy(), frequency=freq)
        r = solver.residual(vm, s, e)
        assert rel_err(r.field, c['r']) < 1e-13
        nrm = solver.residual(vm, s, e, norm=True)
        assert abs(nrm - np.linalg.norm(c['r'])) < 1e-13 * np.linalg.norm(c['r'])
        # lr_dir dispatch incl. dropping 2-cell directions (solver.py:1534-1588)
        for lr_dir in range(8):
            e1 = e.copy()
            solver.smoothing(vm, s, e1, 2, lr_dir, order='lex')
            e2 = c['e'].copy()
            g = mg.Grid([c['hx'], c['hy'], c['hz']])
            ovm = mg.VolumeModel.from_arrays(g, c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'])
            mg.smoothing(ovm, c['s'], e2, 2, lr_dir)
            assert rel_err(e1.field, e2) < (5e-12 if lr_dir == 0 else 2e-10), (k, lr_dir)


def test_volume_model_on_device(golden):
    """Device-side eta/zeta (csrc/model.cu) against the reference's VolumeModel."""
    import emg3d_b200 as eb
    from emg3d_b200 import solver
    gh = golden('host')
    grid = eb.TensorMesh([gh['hx'], gh['hy'], gh['hz']], gh['origin'])
    props = {k: gh['vm_' + k] for k in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r')}
    cases = {'iso': ['property_x'], 'vti': ['property_x', 'property_z'],
             'hti': ['property_x', 'property_y'],
             'tri': ['property_x', 'property_y', 'property_z'], 'full': list(props)}
    for case, keys in cases.items():
        model = eb.Model(grid, **{k: props[k] for k in keys})
        for freq in (0.7, -3.0):
            lv = solver._Level.from_model(model, eb.Field(grid, frequency=freq))
            for k, n in enumerate(('eta_x', 'eta_y', 'eta_z')):
                want = gh[f'vm_{case}_f{freq}_{n}'].ravel('F')
                assert rel_err(lv.eta[k].download(), want) < 1e-15, (case, freq, n)
            assert rel_err(lv.zeta.download(), gh[f'vm_{case}_f{freq}_zeta'].ravel('F')) < 1e-15
            assert (lv.eta[1] is lv.eta[0]) == (case in ('iso', 'vti'))
            assert (lv.eta[2] is lv.eta[0]) == (case in ('iso', 'hti'))
    # other property maps
    for mapping, arr in (('Conductivity', 1 / props['property_x']),
                         ('LgResistivity', np.log10(props['property_x'])),
                         ('LnConductivity', -np.log(props['property_x']))):
        model = eb.Model(grid, property_x=arr, mapping=mapping)
        lv = solver._Level.from_model(model, eb.Field(grid, frequency=0.7))
        assert rel_err(lv.eta[0].download(), gh['vm_iso_f0.7_eta_x'].ravel('F')) < 1e-14


def test_workspace_reuse(golden):
    import emg3d_b200 as eb
    from helpers import solve_case
    c = solve_case(golden('solves'), 'config2_')
    grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], c['origin'])
    model = eb.Model(grid, **c['model'])
    sfield = eb.Field(grid, c['sfield'].copy(), frequency=c['frequency'])
    ws = eb.Workspace()
    e1 = eb.solve(model, sfield, order='lex', workspace=ws, **c['kwargs'])
    lv = ws.level(model, sfield)
    e2 = eb.solve(model, sfield, order='lex', workspace=ws, **c['kwargs'])
    assert ws.level(model, sfield) is lv                # hierarchy was reused
    assert np.array_equal(e1.field, e2.field)
    assert rel_err(e1.field, c['efield']) < 1e-8
    model.property_x[...] *= 2.0                         # in-place change -> rebuilt
    assert ws.level(model, sfield) is not lv


@pytest.mark.parametrize('cplx', [True, False])
def test_z_window_equals_subgrid(core, cplx):
    """Kernels on a z-window of a level (multi-GPU slabs) == kernels on the sub-grid.

    The smoothers and the residual run on cells [z0, z0 + nz) of the full arrays;
    the result must equal the same kernel on arrays sliced out on the host, and
    nothing outside the window may change.
    """
    from emg3d_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    shape, z0, nzw = (10, 7, 13), 4, 6
    c = random_case(rng, shape, cplx, aliased=True)
    dt = c['e'].dtype
    nx, ny, nz = shape
    n = c['e'].size
    e0 = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)   # non-zero halos
    handle = _lib.LevelHandle((c['hx'], c['hy'], c['hz']))
    d_eta = _lib.DeviceArray.from_host(np.asfortranarray(c['eta_x']).ravel('F').astype(dt))
    d_zeta = _lib.DeviceArray.from_host(np.asfortranarray(c['zeta']).ravel('F'))
    handle.set_model(cplx, d_eta, d_eta, d_eta, d_zeta)
    win = handle.window(z0, nzw)
    d_s = _lib.DeviceArray.from_host(c['s'])

    def sub(f):        # sliced copies (fx, fy, fz) of a full field for the window
        fx, fy, fz = split_field(shape, f)
        return [np.asfortranarray(fx[:, :, z0:z0 + nzw + 1]), np.asfortranarray(fy[:, :, z0:z0 + nzw + 1]),
                np.asfortranarray(fz[:, :, z0:z0 + nzw])]

    def scatter(full, parts):
        out = full.copy()
        fx, fy, fz = split_field(shape, out)
        fx[:, :, z0:z0 + nzw + 1], fy[:, :, z0:z0 + nzw + 1], fz[:, :, z0:z0 + nzw] = parts
        return out

    sl = np.s_[:, :, z0:z0 + nzw]
    margs = (np.asfortranarray(c['eta_x'][sl]),) * 3 + (np.asfortranarray(c['zeta'][sl]),
                                                         c['hx'], c['hy'], c['hz'][z0:z0 + nzw])
    margs = (margs[0], margs[0], margs[0]) + margs[3:]
    for ldir, name in enumerate(GS):
        for order in ('lex', 'color'):
            d_e = _lib.DeviceArray.from_host(e0)
            _lib.check(lib.emg3d_b200_gauss_seidel(win.ptr, d_e.ptr, d_s.ptr, 2, ldir,
                                                   core.order_id(order)))
            got = d_e.download()
            es = sub(e0)
            getattr(core, name)(*es, *sub(c['s']), *margs, 2, order=order)
            want = scatter(e0, es)
            assert rel_err(got, want) < 1e-12, (name, order)
    # residual: interior planes of the window agree with the sub-grid residual
    d_e = _lib.DeviceArray.from_host(e0)
    d_r = _lib.DeviceArray(n, dt)
    d_r.zero()
    _lib.check(lib.emg3d_b200_residual(win.ptr, d_s.ptr, d_e.ptr, d_r.ptr, None))
    got = d_r.download()
    rs = sub(c['s'])
    core.amat_x(*rs, *sub(e0), *margs)
    want = scatter(np.zeros(n, dtype=dt), rs)
    assert rel_err(got, want) < 1e-13
    win.free()


def test_sparse_upload_is_bit_identical():
    """emg3d_b200_h2d_sparse: same device bytes as a plain copy, for sparse sources
    (incl. -0.0 and NaN payloads, real and complex) and for dense arrays (fallback)."""
    from emg3d_b200 import _lib
    _lib.init()
    rng = np.random.default_rng(3)
    for dtype, n, bgv in ((np.complex128, 3_000_001, 0), (np.float64, 2_500_003, 0),
                          (np.complex128, 100_000, complex(0.0, -0.0)), (np.float64, 70_001, 3.25)):
        a = np.full(n, bgv, dtype=dtype)
        idx = rng.choice(n, size=min(37, n // 8), replace=False)
        a[idx] = rng.standard_normal(idx.size) + (1j * rng.standard_normal(idx.size)
                                                  if dtype is np.complex128 else 0)
        a[idx[0]] = -0.0
        a[idx[1]] = np.nan
        a[-1] = -2.5
        if n % 2:
            a[0] = 1.5
        d = _lib.DeviceArray(n, dtype)
        assert d.upload_sparse(a)
        assert d.download().tobytes() == a.tobytes()
        dense = rng.standard_normal(n).astype(dtype)
        assert not d.upload_sparse(dense)
        assert d.download().tobytes() == dense.tobytes()
    z = np.zeros(5000)
    d = _lib.DeviceArray(z.size, z.dtype)
    d.upload(np.ones(5000))
    assert d.upload_sparse(z) and not d.download().any()


def test_owned_planes_restrict_the_fused_norm():
    """emg3d_b200_level_set_owned: the norm of the residual kernel counts x/y-edges on
    the owned node planes and z-edges of the layers whose upper plane is owned (what a
    multi-GPU rank contributes to ||r||); plane1 = 0 restores the full norm."""
    from emg3d_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)
    for shape in ((12, 10, 9), (104, 100, 101)):          # simple and plane-streaming kernels
        c = random_case(rng, shape, True, aliased=True)
        nz = shape[2]
        dt = c['e'].dtype
        handle = _lib.LevelHandle((c['hx'], c['hy'], c['hz']))
        d_eta = _lib.DeviceArray.from_host(np.asfortranarray(c['eta_x']).ravel('F').astype(dt))
        d_zeta = _lib.DeviceArray.from_host(np.asfortranarray(c['zeta']).ravel('F'))
        handle.set_model(True, d_eta, d_eta, d_eta, d_zeta)
        d_s, d_e = _lib.DeviceArray.from_host(c['s']), _lib.DeviceArray.from_host(c['e'])
        d_r = _lib.DeviceArray(c['e'].size, dt)
        out = _lib.DeviceArray(2, np.float64)
        _lib.check(lib.emg3d_b200_residual(handle.ptr, d_s.ptr, d_e.ptr, d_r.ptr, out.ptr))
        full = out.download()[0]
        r = d_r.download()
        assert abs(full - np.vdot(r, r).real) <= 1e-12 * full
        rx, ry, rz = split_field(shape, r)
        p0, p1 = 3, nz - 2
        _lib.check(lib.emg3d_b200_level_set_owned(handle.ptr, p0, p1))
        _lib.check(lib.emg3d_b200_residual(handle.ptr, d_s.ptr, d_e.ptr, None, out.ptr))
        want = (np.sum(np.abs(rx[:, :, p0:p1])**2) + np.sum(np.abs(ry[:, :, p0:p1])**2)
                + np.sum(np.abs(rz[:, :, p0 - 1:p1 - 1])**2))
        assert abs(out.download()[0] - want) <= 1e-12 * want
        _lib.check(lib.emg3d_b200_level_set_owned(handle.ptr, 0, 0))
        _lib.check(lib.emg3d_b200_residual(handle.ptr, d_s.ptr, d_e.ptr, None, out.ptr))
        assert abs(out.download()[0] - full) <= 1e-14 * full


def test_peer_memory_api_needs_a_communicator():
    """The p2p entry points fail loudly (no silent fallback) without comm_init."""
    import ctypes
    from emg3d_b200 import _lib
    lib = _lib.init()
    on = ctypes.c_int(1)
    assert lib.emg3d_b200_p2p_init(ctypes.byref(on)) != 0
    slot = ctypes.c_int(0)
    assert lib.emg3d_b200_p2p_register(None, ctypes.byref(slot)) != 0
    st = ctypes.c_int(5)
    _lib.check(lib.emg3d_b200_p2p_status(ctypes.byref(st)))
    assert st.value == 0


def test_magnetic_field_golden(core, golden):
    """core.edge_curl_factor (host-array twin of fields._edge_curl_factor) and
    get_magnetic_field (device VolumeModel + curl kernel) against the reference's
    outputs, <= 1e-13; and against the oracle on a larger random grid."""
    import emg3d_b200 as eb
    gh = golden('hfield')
    for k in range(int(gh['n_cases'])):
        c = hfield_case(gh, k)
        shape = c['shape']
        h = np.full_like(c['h_k'], 7.0)                     # boundary faces must come back zero
        core.edge_curl_factor(*split_faces(shape, h), *split_field(shape, c['e']),
                              c['hx'], c['hy'], c['hz'], c['zeta_k'])
        assert rel_err(h, c['h_k']) < 1e-13
        grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], c['origin'])
        model = eb.Model(grid, c['property_x'], c['property_y'], c['property_z'], mu_r=c['mu_r'])
        ef = eb.Field(grid, c['e'], frequency=c['frequency'])
        hf = eb.get_magnetic_field(model, ef)
        assert not hf.electric and hf.fx.shape == (shape[0] + 1, shape[1], shape[2])
        assert rel_err(hf.field, c['h']) < 1e-13
    rng = np.random.default_rng(11)
    shape = (37, 20, 45)
    c = random_case(rng, shape, True)
    zeta = np.asfortranarray(c['zeta'] * (0.3 - 2j))
    n_faces = sum(int(np.prod(s)) for s in ((38, 20, 45), (37, 21, 45), (37, 20, 46)))
    h_gpu, h_cpu = np.zeros(n_faces, complex), np.zeros(n_faces, complex)
    core.edge_curl_factor(*split_faces(shape, h_gpu), *split_field(shape, c['e']),
                          c['hx'], c['hy'], c['hz'], zeta)
    oracle.edge_curl_factor(*split_faces(shape, h_cpu), *split_field(shape, c['e']),
                            c['hx'], c['hy'], c['hz'], zeta)
    assert rel_err(h_gpu, h_cpu) < 1e-13

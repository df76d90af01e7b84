"""Generate the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference and numba):

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden.py

The reference package hard-imports ``empymod`` and ``scooby`` (not installed,
not on the hot path); ``_refstubs/`` provides two-line stand-ins.  Outputs:

    kernels.npz   amat_x, gauss_seidel{,_x,_y,_z} on small odd-shaped stretched
                  grids, complex and real, triaxial, mu_r != 1, eps_r != 0
    transfer.npz  restrict / restrict_weights / prolongation / model restriction
                  for all seven semicoarsening patterns
    solves.npz    full solves: the reference's own regression data
                  (tests/data/regression.npz: res, reg_2, lap) re-exported as bare
                  arrays, plus small siblings of the five BASELINE.json configs
    host.npz      VolumeModel and source-field vectors (host-side inputs)
    hfield.npz    get_magnetic_field / _edge_curl_factor (the first "next" row of
                  SURVEY.md 8f that shares the stencil family of amat_x)

    maps.npz      volume averaging, edges -> cell averages, receiver sampling (cubic, linear),
                  grid-to-grid interpolation of fields and models

    gcrot.npz     GCROT(m,k) solves with the source scaled to norm one (see make_gcrot)
    cgs.npz       CGS solves
    sources.npz   source assembly: magnetic dipoles, wires, electric points, source objects

The fixtures travel to the GPU box; the reference does not.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, '_refstubs'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, REPO)

import emg3d  # noqa: E402
from emg3d import core, solver  # noqa: E402

from emg3d_b200 import recipes  # noqa: E402


def pec(f):
    f.fx[:, 0, :] = f.fx[:, -1, :] = 0.
    f.fx[:, :, 0] = f.fx[:, :, -1] = 0.
    f.fy[0, :, :] = f.fy[-1, :, :] = 0.
    f.fy[:, :, 0] = f.fy[:, :, -1] = 0.
    f.fz[0, :, :] = f.fz[-1, :, :] = 0.
    f.fz[:, 0, :] = f.fz[:, -1, :] = 0.


def random_case(rng, shape, cplx):
    nx, ny, nz = shape
    hx = 50 * 1.1 ** rng.uniform(-3, 3, nx)
    hy = 60 * 1.2 ** rng.uniform(-3, 3, ny)
    hz = 40 * 1.15 ** rng.uniform(-3, 3, nz)
    g = emg3d.TensorMesh([hx, hy, hz], (-hx.sum() / 2, -hy.sum() / 2, -hz.sum() / 2))
    rx = 10 ** rng.uniform(-0.5, 1.5, g.shape_cells)
    m = emg3d.Model(g, rx, 1.5 * rx * rng.uniform(.5, 2, g.shape_cells), 3 * rx,
                    mu_r=rng.uniform(1, 2, g.shape_cells),
                    epsilon_r=rng.uniform(1, 10, g.shape_cells))
    freq = 1.3 if cplx else -1.3
    sf = emg3d.Field(g, frequency=freq)
    ef = emg3d.Field(g, frequency=freq)
    n = sf.field.size
    sf.field[:] = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    ef.field[:] = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    pec(ef)
    return g, emg3d.models.VolumeModel(m, sf), sf, ef


def make_kernels():
    rng = np.random.default_rng(1)
    out = {}
    cases = [((6, 4, 8), True), ((2, 4, 6), True), ((5, 3, 2), False), ((7, 2, 3), True),
             ((8, 8, 8), False), ((8, 8, 8), True), ((4, 10, 6), True)]
    for k, (shape, cplx) in enumerate(cases):
        g, vm, sf, ef = random_case(rng, shape, cplx)
        p = f"k{k}_"
        out[p + 'hx'], out[p + 'hy'], out[p + 'hz'] = g.h
        out[p + 'eta_x'], out[p + 'eta_y'], out[p + 'eta_z'] = vm.eta_x, vm.eta_y, vm.eta_z
        out[p + 'zeta'] = vm.zeta
        out[p + 's'] = np.asarray(sf.field)
        out[p + 'e'] = np.asarray(ef.field)
        args = (vm.eta_x, vm.eta_y, vm.eta_z, vm.zeta, g.h[0], g.h[1], g.h[2])
        r = sf.copy()
        core.amat_x(r.fx, r.fy, r.fz, ef.fx, ef.fy, ef.fz, *args)
        out[p + 'r'] = np.asarray(r.field)
        for ldir, name in enumerate(['gauss_seidel', 'gauss_seidel_x', 'gauss_seidel_y',
                                     'gauss_seidel_z']):
            for nu in (1, 2):
                e1 = ef.copy()
                getattr(core, name)(e1.fx, e1.fy, e1.fz, sf.fx, sf.fy, sf.fz, *args, nu)
                out[p + f'gs{ldir}_nu{nu}'] = np.asarray(e1.field)
    out['n_cases'] = len(cases)
    np.savez_compressed(os.path.join(HERE, 'kernels.npz'), **out)


def make_transfer():
    rng = np.random.default_rng(2)
    out = {}
    cases = [((8, 4, 12), True), ((4, 6, 2), True), ((6, 2, 4), False), ((16, 8, 8), True)]
    sc_flags = {0: (1, 1, 1), 1: (0, 1, 1), 2: (1, 0, 1), 3: (1, 1, 0),
                4: (1, 0, 0), 5: (0, 1, 0), 6: (0, 0, 1)}
    for k, (shape, cplx) in enumerate(cases):
        g, vm, sf, ef = random_case(rng, shape, cplx)
        p = f"t{k}_"
        out[p + 'hx'], out[p + 'hy'], out[p + 'hz'] = g.h
        out[p + 'origin'] = g.origin
        out[p + 'eta_x'], out[p + 'eta_y'], out[p + 'eta_z'] = vm.eta_x, vm.eta_y, vm.eta_z
        out[p + 'zeta'] = vm.zeta
        out[p + 'r'] = np.asarray(ef.field)      # used as the fine "residual"
        out[p + 'e'] = np.asarray(sf.field)      # fine field to prolongate onto
        scs = []
        for sc, fl in sc_flags.items():
            if any(f and (n % 2 or n < 2) for f, n in zip(fl, shape)):
                continue
            scs.append(sc)
            cm, cs, ce = solver.restriction(vm, sf, ef, sc)
            q = p + f"sc{sc}_"
            out[q + 'cs'] = np.asarray(cs.field)
            out[q + 'ceta_x'], out[q + 'ceta_y'], out[q + 'ceta_z'] = cm.eta_x, cm.eta_y, cm.eta_z
            out[q + 'czeta'] = cm.zeta
            wx, wy, wz = solver._get_restriction_weights(g, cm.grid, sc)
            for a, w in zip('xyz', (wx, wy, wz)):
                out[q + f'w{a}'] = np.array(w)
            ce.field[:] = rng.standard_normal(ce.field.size)
            out[q + 'ce'] = np.asarray(ce.field)
            e1 = sf.copy()
            solver.prolongation(e1, ce, sc)
            out[q + 'e_out'] = np.asarray(e1.field)
        out[p + 'sc_dirs'] = np.array(scs)
    out['n_cases'] = len(cases)
    np.savez_compressed(os.path.join(HERE, 'transfer.npz'), **out)


def _store_solve(out, p, grid, model, sfield, kwargs, freq, source=None):
    """Run the reference and store inputs + result under prefix p."""
    out[p + 'hx'], out[p + 'hy'], out[p + 'hz'] = grid.h
    out[p + 'origin'] = np.asarray(grid.origin, dtype=float)
    for name in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r'):
        v = getattr(model, name)
        if v is not None:
            out[p + name] = np.asarray(v)
    out[p + 'frequency'] = float(freq)
    out[p + 'sfield'] = np.asarray(sfield.field)
    if source is not None:
        out[p + 'source'] = np.asarray(source, dtype=float)
    buf = io.StringIO()
    kw = dict(kwargs)
    with redirect_stdout(buf):
        efield, info = solver.solve(model, sfield, return_info=True, **kw)
    out[p + 'efield'] = np.asarray(efield.field)
    out[p + 'kwargs'] = json.dumps(kwargs)
    out[p + 'it_mg'] = info['it_mg']
    out[p + 'it_ssl'] = info['it_ssl']
    out[p + 'exit_message'] = info['exit_message']
    out[p + 'abs_error'] = info['abs_error']
    out[p + 'ref_error'] = info['ref_error']
    out[p + 'error_at_cycle'] = info['error_at_cycle']
    out[p + 'stdout'] = buf.getvalue()
    print(p, kwargs, '->', info['exit_message'], info['it_mg'], info['it_ssl'],
          f"{info['rel_error']:.3e}")


def make_solves():
    out = {}
    reg = emg3d.load('/root/reference/tests/data/regression.npz', verb=0)

    # --- reference regression data: 'res' (tests/test_solver.py:18-70)
    dat = reg['res']
    model = emg3d.Model(**dat['input_model'])
    src = dat['input_source']
    sfield = emg3d.get_source_field(**src)
    for key, kw in (('F', dict(plain=True, verb=4)), ('W', dict(plain=True, cycle='W')),
                    ('V', dict(plain=True, cycle='V')),
                    ('bic', dict(verb=4, sslsolver='bicgstab', plain=True))):
        p = f"res_{key}_"
        _store_solve(out, p, model.grid, model, sfield, kw, src['frequency'], src['source'])
        # the stored regression result must agree with what we just computed
        want = dat[{'F': 'Fresult', 'W': 'Wresult', 'V': 'Vresult', 'bic': 'bicresult'}[key]]
        np.testing.assert_allclose(want.field, out[p + 'efield'])
        out[p + 'regression'] = np.asarray(want.field)

    # --- 'reg_2' (tests/test_solver.py:152-176)
    dat = reg['reg_2']
    model, sfield, inp = dat['model'], dat['sfield'], dict(dat['inp'])
    for n in ['nu_init', 'nu_pre', 'nu_coarse', 'nu_post', 'clevel', 'maxit',
              'semicoarsening', 'linerelaxation', 'verb']:
        inp[n] = int(inp[n])
    inp['tol'] = float(inp['tol'])
    inp['sslsolver'] = False
    _store_solve(out, 'reg2_', model.grid, model, sfield, inp, sfield._frequency)
    np.testing.assert_allclose(dat['result'].field, out['reg2_efield'])
    out['reg2_regression'] = np.asarray(dat['result'].field)

    # --- 'lap' (tests/test_solver.py:227-247), Laplace domain = real arithmetic
    dat = reg['lap']
    model = emg3d.Model(**dat['input_model'])
    src = dat['input_source']
    sfield = emg3d.get_source_field(**src)
    _store_solve(out, 'lap_F_', model.grid, model, sfield, dict(plain=True),
                 src['frequency'], src['source'])
    np.testing.assert_allclose(dat['Fresult'].field, out['lap_F_efield'], atol=1e-14)
    out['lap_F_regression'] = np.asarray(dat['Fresult'].field)
    _store_solve(out, 'lap_bic_', model.grid, model, sfield,
                 dict(semicoarsening=False, linerelaxation=False),
                 src['frequency'], src['source'])
    np.testing.assert_allclose(dat['bicresult'].field, out['lap_bic_efield'], atol=1e-14)
    out['lap_bic_regression'] = np.asarray(dat['bicresult'].field)

    # --- small siblings of the BASELINE.json configurations
    for name, n in (('config1', 32), ('config2', 32), ('config3', 32), ('config4', 32),
                    ('config5', 32)):
        cfg = recipes.config(name, n)
        grid = emg3d.TensorMesh(cfg['h'], cfg['origin'])
        model = emg3d.Model(grid, **cfg['model'])
        sfield = emg3d.get_source_field(grid, cfg['source'], cfg['frequency'])
        _store_solve(out, f"{name}_", grid, model, sfield, cfg['solver'], cfg['frequency'],
                     cfg['source'])
    # tight-tolerance versions (for the multicolour ordering: agreement at convergence)
    for name in ('config2', 'config3'):
        cfg = recipes.config(name, 32)
        if name == 'config3':
            cfg['model'] = recipes.model_marine(cfg['h'], cfg['origin'], rho_air=1e4)
        grid = emg3d.TensorMesh(cfg['h'], cfg['origin'])
        model = emg3d.Model(grid, **cfg['model'])
        sfield = emg3d.get_source_field(grid, cfg['source'], cfg['frequency'])
        kw = dict(cfg['solver'], tol=1e-11, maxit=60)
        _store_solve(out, f"{name}_tight_", grid, model, sfield, kw, cfg['frequency'],
                     cfg['source'])
    np.savez_compressed(os.path.join(HERE, 'solves.npz'), **out)


def make_host():
    """VolumeModel and source vectors, to pin the host-side input builders."""
    out = {}
    rng = np.random.default_rng(3)
    hx = recipes.widths(8, 1.2, 30.)
    hy = recipes.widths(6, 1.1, 40.)
    hz = recipes.widths(10, 1.3, 20.)
    origin = (-hx.sum() / 2, -hy.sum() / 2 + 3., -hz.sum() / 2 - 7.)
    grid = emg3d.TensorMesh([hx, hy, hz], origin)
    out['hx'], out['hy'], out['hz'], out['origin'] = hx, hy, hz, np.array(origin)
    sources = [(0., 0., 0., 0., 0.), (12.3, -20.1, 5.5, 30., 20.), (-31., 14., -60., -50., 75.),
               (-40., 40., -10., 10., -20., 30.)]
    for k, src in enumerate(sources):
        for freq in (1.0, -2.5):
            sf = emg3d.get_source_field(grid, src, freq)
            out[f'src{k}_f{freq}'] = np.asarray(sf.field)
        out[f'src{k}'] = np.array(src)
    out['n_sources'] = len(sources)
    shape = grid.shape_cells
    props = dict(property_x=10 ** rng.uniform(-1, 2, shape), property_y=10 ** rng.uniform(-1, 2, shape),
                 property_z=10 ** rng.uniform(-1, 2, shape), mu_r=rng.uniform(1, 3, shape),
                 epsilon_r=rng.uniform(1, 20, shape))
    for k, v in props.items():
        out['vm_' + k] = v
    for case, keys in (('iso', ['property_x']), ('vti', ['property_x', 'property_z']),
                       ('hti', ['property_x', 'property_y']),
                       ('tri', ['property_x', 'property_y', 'property_z']),
                       ('full', list(props))):
        model = emg3d.Model(grid, **{k: props[k] for k in keys})
        for freq in (0.7, -3.0):
            sf = emg3d.Field(grid, frequency=freq)
            vm = emg3d.models.VolumeModel(model, sf)
            for n in ('eta_x', 'eta_y', 'eta_z', 'zeta'):
                out[f'vm_{case}_f{freq}_{n}'] = np.asarray(getattr(vm, n))
    np.savez_compressed(os.path.join(HERE, 'host.npz'), **out)


def make_hfield():
    """get_magnetic_field / _edge_curl_factor (emg3d/fields.py:617-659, 941-1009)."""
    from emg3d import fields
    rng = np.random.default_rng(4)
    out = {}
    cases = [((6, 4, 8), True), ((5, 3, 2), False), ((2, 2, 2), True), ((9, 7, 5), True)]
    for k, (shape, cplx) in enumerate(cases):
        nx, ny, nz = shape
        hx = 50 * 1.1 ** rng.uniform(-3, 3, nx)
        hy = 60 * 1.2 ** rng.uniform(-3, 3, ny)
        hz = 40 * 1.15 ** rng.uniform(-3, 3, nz)
        g = emg3d.TensorMesh([hx, hy, hz], (-hx.sum() / 2, -hy.sum() / 2, -hz.sum() / 2))
        rx = 10 ** rng.uniform(-0.5, 1.5, g.shape_cells)
        m = emg3d.Model(g, rx, 1.5 * rx, 3 * rx, mu_r=rng.uniform(1, 2, g.shape_cells))
        freq = 1.3 if cplx else -1.3
        ef = emg3d.Field(g, frequency=freq)
        n = ef.field.size
        ef.field[:] = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
        hf = fields.get_magnetic_field(m, ef)
        p = f"h{k}_"
        out[p + 'hx'], out[p + 'hy'], out[p + 'hz'] = g.h
        out[p + 'origin'] = np.asarray(g.origin, dtype=float)
        out[p + 'property_x'], out[p + 'property_y'], out[p + 'property_z'] = (
            m.property_x, m.property_y, m.property_z)
        out[p + 'mu_r'] = m.mu_r
        out[p + 'frequency'] = freq
        out[p + 'e'] = np.asarray(ef.field)
        out[p + 'h'] = np.asarray(hf.field)
        # the kernel alone, with an arbitrary (complex / real) factor array
        zeta = rng.uniform(1, 2, g.shape_cells) * ((1 + 0.5j) if cplx else 1.0)
        zeta = np.asfortranarray(zeta)
        hk = emg3d.Field(g, frequency=freq, electric=False)
        fields._edge_curl_factor(hk.fx, hk.fy, hk.fz, ef.fx, ef.fy, ef.fz, *g.h, zeta)
        out[p + 'zeta_k'] = zeta
        out[p + 'h_k'] = np.asarray(hk.field)
    out['n_cases'] = len(cases)
    np.savez_compressed(os.path.join(HERE, 'hfield.npz'), **out)


def make_maps():
    """Interpolation next to the solve: volume averaging, edges -> cell
    averages, receiver sampling, grid-to-grid interpolation of fields and models."""
    from emg3d import fields, maps
    rng = np.random.default_rng(11)
    out = {}

    def grid_of(n, scale, shift=0.0):
        h = [scale * 1.1 ** rng.uniform(-2, 2, m) for m in n]
        return emg3d.TensorMesh(h, tuple(-0.5 * hh.sum() + shift for hh in h))

    # --- interp_volume_average: output grid inside, across and beyond the input grid
    cases = [((7, 5, 6), (4, 6, 3), 1.3, 10.), ((6, 6, 6), (9, 8, 7), 0.8, -5.), ((5, 4, 3), (5, 4, 3), 1.0, 0.)]
    for k, (n_in, n_out, sc, shift) in enumerate(cases):
        gi, go = grid_of(n_in, 50.), grid_of(n_out, 50. * sc * np.array(n_in).mean() / np.array(n_out).mean(), shift)
        vals = np.asfortranarray(10 ** rng.uniform(-1, 2, gi.shape_cells))
        new = np.zeros(go.shape_cells, order='F')
        maps.interp_volume_average(gi.nodes_x, gi.nodes_y, gi.nodes_z, vals, go.nodes_x, go.nodes_y,
                                   go.nodes_z, new, go.cell_volumes.reshape(go.shape_cells, order='F'))
        p = f"va{k}_"
        out[p + 'h_in'] = np.concatenate(gi.h); out[p + 'n_in'] = np.array(n_in)
        out[p + 'o_in'] = np.array(gi.origin, float)
        out[p + 'h_out'] = np.concatenate(go.h); out[p + 'n_out'] = np.array(n_out)
        out[p + 'o_out'] = np.array(go.origin, float)
        out[p + 'values'], out[p + 'new'] = vals, new
        out[p + 'new_log'] = maps.interpolate(gi, vals, go, method='volume', log=True)
        for ax, (a, b) in enumerate(zip((gi.nodes_x, gi.nodes_y, gi.nodes_z), (go.nodes_x, go.nodes_y, go.nodes_z))):
            w, ii, io = maps._volume_average_weights(a, b)
            out[p + f'w{ax}'], out[p + f'ii{ax}'], out[p + f'io{ax}'] = w, ii, io
    out['n_va'] = len(cases)

    # --- interp_edges_to_vol_averages, real and complex
    for k, (n, cplx) in enumerate([((5, 4, 6), False), ((3, 7, 4), True), ((1, 2, 3), True)]):
        g = grid_of(n, 30.)
        f = emg3d.Field(g, frequency=1.0 if cplx else -1.0)
        f.field[:] = rng.standard_normal(f.field.size) + (1j * rng.standard_normal(f.field.size) if cplx else 0)
        vol = g.cell_volumes.reshape(g.shape_cells, order='F')
        o = [np.zeros(g.shape_cells, order='F', dtype=f.field.dtype) for _ in range(3)]
        maps.interp_edges_to_vol_averages(f.fx, f.fy, f.fz, vol, *o)
        p = f"ev{k}_"
        out[p + 'h'] = np.concatenate(g.h); out[p + 'n'] = np.array(n)
        out[p + 'field'] = np.asarray(f.field)
        out[p + 'ox'], out[p + 'oy'], out[p + 'oz'] = o
    out['n_ev'] = 3

    # --- receivers (cubic / linear) and grid-to-grid interpolation of a field and a model
    for k, (n, cplx) in enumerate([((12, 10, 9), True), ((8, 9, 7), False)]):
        g = grid_of(n, 40.)
        f = emg3d.Field(g, frequency=0.8 if cplx else -0.8)
        # a smooth field plus noise
        for comp, arr in enumerate((f.fx, f.fy, f.fz)):
            arr[...] = rng.standard_normal(arr.shape) + (1j * rng.standard_normal(arr.shape) if cplx else 0)
        nrec = 40
        lo = [g.nodes_x[0], g.nodes_y[0], g.nodes_z[0]]
        hi = [g.nodes_x[-1], g.nodes_y[-1], g.nodes_z[-1]]
        rec = [rng.uniform(l - 20, h + 20, nrec) for l, h in zip(lo, hi)]
        rec[0][:4] = [g.nodes_x[1], g.nodes_x[-2], g.nodes_x[2], 0.0]     # on the PEC limits / nodes
        rec += [rng.uniform(-180, 180, nrec), rng.uniform(-90, 90, nrec)]
        rec[3][:6] = [0, 90, 0, 0, 45, 0]; rec[4][:6] = [0, 0, 90, -90, 0, 0]
        p = f"rc{k}_"
        out[p + 'h'] = np.concatenate(g.h); out[p + 'n'] = np.array(n)
        out[p + 'origin'] = np.array(g.origin, float)
        out[p + 'field'] = np.asarray(f.field); out[p + 'frequency'] = f.frequency if cplx else -f.sval.real / (2 * np.pi) * -1
        out[p + 'freq_arg'] = 0.8 if cplx else -0.8
        out[p + 'rec'] = np.array(rec)
        out[p + 'cubic'] = np.asarray(fields.get_receiver(f, tuple(rec), 'cubic'))
        out[p + 'linear'] = np.asarray(fields.get_receiver(f, tuple(rec), 'linear'))
        g2 = grid_of(tuple(m - 2 for m in n), 55., 15.)
        out[p + 'h2'] = np.concatenate(g2.h); out[p + 'n2'] = np.array(g2.shape_cells)
        out[p + 'origin2'] = np.array(g2.origin, float)
        out[p + 'f2_cubic'] = np.asarray(f.interpolate_to_grid(g2).field)
        out[p + 'f2_linear'] = np.asarray(f.interpolate_to_grid(g2, method='linear').field)
        out[p + 'f2_cubic_extrap'] = np.asarray(f.interpolate_to_grid(g2, extrapolate=True).field)
        rx = 10 ** rng.uniform(-0.5, 1.5, g.shape_cells)
        m = emg3d.Model(g, rx, 2 * rx, 3 * rx, mapping='Resistivity')
        m2 = m.interpolate_to_grid(g2)
        out[p + 'prop'] = rx
        out[p + 'prop2_x'], out[p + 'prop2_z'] = m2.property_x, m2.property_z
    out['n_rc'] = 2
    np.savez_compressed(os.path.join(HERE, 'maps.npz'), **out)


def make_gcrot():
    """GCROT(m,k) solves.  The reference's multigrid preconditioner keeps the tolerance reference
    of the ORIGINAL source (solver.py:1285-1300 / _terminate) while GCROT hands it vectors of norm
    one, so it reports DIVERGED for a source of small norm ('res' as is: see solves.npz users);
    with the source scaled to norm one it converges.  Those are the cases stored here."""
    out = {}
    reg = emg3d.load('/root/reference/tests/data/regression.npz', verb=0)
    dat = reg['res']
    model = emg3d.Model(**dat['input_model'])
    src = dat['input_source']
    sfield = emg3d.get_source_field(**src)
    sfield.field /= np.linalg.norm(sfield.field)
    _store_solve(out, 'res_gcrot_', model.grid, model, sfield,
                 dict(plain=True, sslsolver='gcrotmk', verb=4), src['frequency'])
    _store_solve(out, 'res_gcrot_noprec_', model.grid, model, sfield,
                 dict(plain=True, sslsolver='gcrotmk', cycle=None, maxit=3), src['frequency'])
    cfg = recipes.config('config2', 32)
    grid = emg3d.TensorMesh(cfg['h'], cfg['origin'])
    model = emg3d.Model(grid, **cfg['model'])
    sfield = emg3d.get_source_field(grid, cfg['source'], cfg['frequency'])
    sfield.field /= np.linalg.norm(sfield.field)
    _store_solve(out, 'config2_gcrot_', grid, model, sfield,
                 dict(sslsolver='gcrotmk', semicoarsening=True, linerelaxation=True, cycle='V'),
                 cfg['frequency'])
    np.savez_compressed(os.path.join(HERE, 'gcrot.npz'), **out)


def make_sources():
    """Source assembly beyond electric dipoles (emg3d/fields.py:386-519): magnetic dipoles (square
    loops), wires, electric point sources, strengths, the frequency-independent vector."""
    out = {}
    hx = recipes.widths(8, 1.2, 30.)
    hy = recipes.widths(6, 1.1, 40.)
    hz = recipes.widths(10, 1.3, 20.)
    origin = (-hx.sum() / 2, -hy.sum() / 2 + 3., -hz.sum() / 2 - 7.)
    grid = emg3d.TensorMesh([hx, hy, hz], origin)
    out['hx'], out['hy'], out['hz'], out['origin'] = hx, hy, hz, np.array(origin)
    wire = np.array([[-40., -30., -50.], [-10., 12., -20.], [25., 12., 33.], [60., -44., 33.]])
    top = (grid.nodes_x[-1], grid.nodes_y[-1], grid.nodes_z[-1])
    cases = [
        # name, source as passed to the reference, kwargs
        ('mag5', (12.3, -20.1, 5.5, 30., 20.), dict(electric=False, strength=2.5, length=30.)),
        ('mag5z', (0., 0., 0., 0., 90.), dict(electric=False)),
        ('mag6', (-40., 40., -10., 10., -20., 30.), dict(electric=False, strength=0.5)),
        ('mag23', np.array([[-20., 5., -33.], [10., 5., -33.]]), dict(electric=False)),
        ('dip23', np.array([[-20., 5., -33.], [10., 25., -3.]]), dict(strength=3.0)),
        ('wire', wire, dict(strength=1.5)),
        ('cstr', (12.3, -20.1, 5.5, 30., 20.), dict(strength=2 - 3j, length=12.)),
    ]
    for name, src, kw in cases:
        out[f'{name}_source'] = np.asarray(src, dtype=float)
        out[f'{name}_kwargs'] = json.dumps({k: ([v.real, v.imag] if isinstance(v, complex) else v)
                                            for k, v in kw.items()})
        for freq in (1.0, -2.5, None):
            if freq is not None and freq < 0 and isinstance(kw.get('strength'), complex):
                continue                                   # (complex strength: frequency domain only)
            if freq is None and isinstance(kw.get('strength'), complex):
                continue
            sf = emg3d.get_source_field(grid, src, freq, **kw)
            out[f'{name}_f{freq}'] = np.asarray(sf.field)
    # source objects: points (trilinear adjoint), incl. the last cell / the grid corner
    objs = [
        ('pt', emg3d.TxElectricPoint((12.3, -20.1, 5.5, 30., 20.), strength=1.5)),
        ('pt_node', emg3d.TxElectricPoint((grid.nodes_x[3], grid.nodes_y[2], grid.nodes_z[4], -70., 15.))),
        ('pt_top', emg3d.TxElectricPoint((*top, 45., 45.), strength=2.0)),
        ('pt_low', emg3d.TxElectricPoint((grid.nodes_x[0], grid.nodes_y[0], grid.nodes_z[0], 10., -30.))),
        ('obj_wire', emg3d.TxElectricWire(wire, strength=4.0)),
        ('obj_mag', emg3d.TxMagneticDipole((5., 6., -7., 20., -40.), strength=3.0, length=50.)),
        ('obj_dip', emg3d.TxElectricDipole((-40., 40., -10., 10., -20., 30.), strength=0.25)),
    ]
    for name, obj in objs:
        out[f'{name}_class'] = type(obj).__name__
        out[f'{name}_points'] = np.asarray(obj.points, dtype=float)
        out[f'{name}_coordinates'] = np.asarray(obj.coordinates, dtype=float)
        out[f'{name}_strength'] = float(obj.strength)
        for freq in (1.0, -2.5, None):
            out[f'{name}_f{freq}'] = np.asarray(emg3d.get_source_field(grid, obj, freq).field)
    out['cases'] = json.dumps([c[0] for c in cases])
    out['objects'] = json.dumps([o[0] for o in objs])
    np.savez_compressed(os.path.join(HERE, 'sources.npz'), **out)


def make_cgs():
    """CGS solves (the third ``sslsolver`` of the reference, solver.py:763-765)."""
    out = {}
    reg = emg3d.load('/root/reference/tests/data/regression.npz', verb=0)
    dat = reg['res']
    model = emg3d.Model(**dat['input_model'])
    src = dat['input_source']
    sfield = emg3d.get_source_field(**src)
    _store_solve(out, 'res_cgs_', model.grid, model, sfield,
                 dict(plain=True, sslsolver='cgs'), src['frequency'])
    cfg = recipes.config('config2', 32)
    grid = emg3d.TensorMesh(cfg['h'], cfg['origin'])
    model = emg3d.Model(grid, **cfg['model'])
    sfield = emg3d.get_source_field(grid, cfg['source'], cfg['frequency'])
    _store_solve(out, 'config2_cgs_', grid, model, sfield,
                 dict(sslsolver='cgs', semicoarsening=True, linerelaxation=True, cycle='V'),
                 cfg['frequency'])
    np.savez_compressed(os.path.join(HERE, 'cgs.npz'), **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['kernels', 'transfer', 'solves', 'host', 'hfield', 'maps', 'gcrot', 'cgs', 'sources']
    for w in which:
        globals()['make_' + w]()
    for f in ('kernels', 'transfer', 'solves', 'host', 'hfield', 'maps', 'gcrot', 'cgs', 'sources'):
        fn = os.path.join(HERE, f + '.npz')
        if os.path.exists(fn):
            print(f, os.path.getsize(fn) // 1024, 'KiB')

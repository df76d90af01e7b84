"""Seeded synthetic inputs for the configurations named in BASELINE.json.

Plain arrays only (widths, origin, resistivities, source, frequency), so the same
numbers can be fed to this package, to the CPU oracle and -- in the build
container, when generating golden vectors -- to the reference itself.
Recipes follow SURVEY.md section 8(d) / Appendix C.
"""
import numpy as np

SEED = 20261017


def widths(n, alpha, core=50.):
    """n cells: n/2 uniform core cells, n/4 stretched padding cells per side."""
    npad = n // 4
    pad = core * alpha ** (np.arange(npad) + 1)
    return np.r_[pad[::-1], np.ones(n - 2 * npad) * core, pad]


def grid_arrays(nx, ny, nz, ax, ay, az, zc=-1000.):
    """Widths and origin of a stretched grid centred on (0, 0, zc)."""
    h = [widths(nx, ax), widths(ny, ay), widths(nz, az)]
    origin = (-h[0].sum() / 2, -h[1].sum() / 2, zc - h[2].sum() / 2)
    return h, origin


def _centers(h, origin):
    out = []
    for a in range(3):
        nodes = np.r_[0., h[a].cumsum()] + origin[a]
        out.append((nodes[1:] + nodes[:-1]) / 2)
    return out


def model_halfspace(h, origin):
    """Isotropic half-space: 1 Ohm.m below z = 0, 100 Ohm.m above (config 1)."""
    cx, cy, cz = _centers(h, origin)
    shape = (cx.size, cy.size, cz.size)
    Z = cz[None, None, :] * np.ones(shape)
    return {'property_x': np.where(Z > 0, 100., 1.)}


def model_triaxial(h, origin, rng):
    """Layered background with log-normal perturbation, triaxial (configs 2, 5)."""
    cx, cy, cz = _centers(h, origin)
    shape = (cx.size, cy.size, cz.size)
    Z = cz[None, None, :] * np.ones(shape)
    rho_h = np.where(Z > 0, 0.3, np.where(Z > -1500, 1., np.where(Z > -2500, 2., 5.)))
    rx = rho_h * np.exp(0.2 * rng.standard_normal(shape))
    return {'property_x': rx, 'property_y': 1.5 * rx, 'property_z': 3 * rx}


def model_marine(h, origin, rho_air=1e8):
    """Air / sea / VTI sediment with a thin resistor (configs 3, 4)."""
    cx, cy, cz = _centers(h, origin)
    shape = (cx.size, cy.size, cz.size)
    rh1 = np.where(cz > 0, rho_air, np.where(cz > -1000, 0.3, 1.0))
    rv1 = np.where(cz > 0, rho_air, np.where(cz > -1000, 0.3, 2.0))
    rh = np.empty(shape, order='F')
    rv = np.empty(shape, order='F')
    rh[:] = rh1[None, None, :]
    rv[:] = rv1[None, None, :]
    ix = np.flatnonzero(abs(cx) < 2500)
    iy = np.flatnonzero(abs(cy) < 2500)
    iz = np.flatnonzero((cz < -1900) & (cz > -2000))
    if ix.size and iy.size and iz.size:
        rh[np.ix_(ix, iy, iz)] = 100.
        rv[np.ix_(ix, iy, iz)] = 100.
    return {'property_x': rh, 'property_z': rv}


def config(name, n=None):
    """Inputs of one named configuration, optionally at a reduced size ``n``.

    Returns a dict with ``h`` (3 width arrays), ``origin``, ``model`` (kwargs of
    ``Model``), ``source`` (x, y, z, azimuth, elevation), ``frequency`` and
    ``solver`` (kwargs of ``solve``).
    """
    rng = np.random.default_rng(SEED)
    if name == 'config1':          # 32^3 uniform half-space, one V(2,2)-cycle
        n = n or 32
        h = [np.ones(n) * 50.] * 3
        origin = (-25. * n, -25. * n, -25. * n)
        return dict(h=h, origin=origin, model=model_halfspace(h, origin),
                    source=(0., 0., -100., 0., 0.), frequency=1.0,
                    solver=dict(plain=True, cycle='V', maxit=1))
    if name == 'config2':          # 128^3 stretched, triaxial, F-cycle
        n = n or 128
        alpha = {128: 1.04, 64: 1.06, 32: 1.10}.get(n, 1.10)
        h, origin = grid_arrays(n, n, n, alpha, alpha, alpha)
        return dict(h=h, origin=origin, model=model_triaxial(h, origin, rng),
                    source=(0., 0., -950., 0., 0.), frequency=1.0,
                    solver=dict(sslsolver=False, cycle='F'))
    if name == 'config3':          # 256^3 marine CSEM, V-cycle + BiCGSTAB
        n = n or 256
        alpha = {256: 1.03, 128: 1.04, 64: 1.06, 32: 1.10}.get(n, 1.10)
        h, origin = grid_arrays(n, n, n, alpha, alpha, alpha)
        return dict(h=h, origin=origin, model=model_marine(h, origin),
                    source=(0., 0., -950., 0., 0.), frequency=1.0,
                    solver=dict(cycle='V', sslsolver='bicgstab'))
    if name == 'config4':          # 512 x 512 x 256, F-cycle, sc + lr
        n = n or 512
        axy = {512: 1.02, 256: 1.03, 128: 1.04, 64: 1.06, 32: 1.10}.get(n, 1.10)
        az = {512: 1.03, 256: 1.04, 128: 1.06, 64: 1.10, 32: 1.15}.get(n, 1.15)
        h, origin = grid_arrays(n, n, n // 2, axy, axy, az)
        return dict(h=h, origin=origin, model=model_marine(h, origin),
                    source=(0., 0., -950., 0., 0.), frequency=1.0,
                    solver=dict(sslsolver=False, cycle='F', semicoarsening=True,
                                linerelaxation=True))
    if name == 'config5':          # 512^3 stretched, triaxial, W-cycle
        n = n or 512
        alpha = {512: 1.02, 256: 1.03, 128: 1.04, 64: 1.06, 32: 1.10}.get(n, 1.10)
        h, origin = grid_arrays(n, n, n, alpha, alpha, alpha)
        return dict(h=h, origin=origin, model=model_triaxial(h, origin, rng),
                    source=(0., 0., -950., 0., 0.), frequency=1.0,
                    solver=dict(cycle='W', sslsolver=False))
    raise ValueError(f"unknown configuration {name!r}")


def bench_shape(n_gpus, n=256):
    """Cells of the weak-scaling workload of bench.py (see :func:`bench_grid`)."""
    shape = [n, n, n]
    k, a = n_gpus, 0
    while k > 1:
        shape[a % 3] *= 2
        k //= 2
        a += 1
    return shape


def bench_grid(n_gpus, n=256):
    """Weak-scaling workload of bench.py: the marine model on n^3 cells per GPU.

    1 GPU: n^3 (BASELINE.json configs[2]); the cell count doubles with the GPU
    count along x, then y, then z: 2: 2n x n x n, 4: 2n x 2n x n (configs[3] shape),
    8: (2n)^3 (configs[4] shape).  Decomposed into z-slabs.
    """
    shape = bench_shape(n_gpus, n)
    alpha = [{512: 1.02, 256: 1.03, 128: 1.04, 64: 1.06, 32: 1.10}.get(m, 1.02 if m > 512 else 1.10)
             for m in shape]
    h, origin = grid_arrays(*shape, *alpha)
    return dict(h=h, origin=origin, model=model_marine(h, origin),
                source=(0., 0., -950., 0., 0.), frequency=1.0)

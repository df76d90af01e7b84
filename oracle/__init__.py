"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the emg3d multigrid hot path.

A plain-C restatement (``emg3d_oracle.c`` + ``kernels.inc``) of the reference's
numba kernels (emg3d/core.py) with the same call signatures, loaded via ctypes,
plus :mod:`oracle.mg`, a NumPy restatement of the driver (emg3d/solver.py).

Parity status: **pinned** -- checked against outputs of the reference itself
(imported in the build container by ``tests/golden/make_golden.py``) and against
the reference's own regression data (``tests/data/regression.npz``), see
``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  The product
(``emg3d_b200``) never does and has no CPU fallback.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, 'libemg3d_oracle.so')


_FASTPATH = os.path.join(_HERE, 'libemg3d_oracle_fast.so')


def build(force=False):
    """Compile the oracle with gcc (a few seconds): the strict build and the fast-math build
    used by :func:`noise_floor`."""
    stale = lambda path: not os.path.exists(path) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(path)
        for f in ('emg3d_oracle.c', 'kernels.inc', 'Makefile'))
    if force or stale(_LIBPATH) or stale(_FASTPATH):
        subprocess.run(['make', '-C', _HERE, '-s', '-B', 'all'], check=True)


_lib = None
_libs = {}


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = _libs['strict'] = ctypes.CDLL(_LIBPATH)
    return _lib


class variant:
    """Context manager: run the oracle kernels from another build of the same sources.

    ``'fast'`` = value-changing optimisations on (reassociation, FMA contraction), like the
    reference's ``fastmath=True`` numba kernels; ``'strict'`` = the default build.
    """

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        global _lib
        lib()
        if self.name not in _libs:
            _libs[self.name] = ctypes.CDLL({'fast': _FASTPATH, 'strict': _LIBPATH}[self.name])
        self._prev, _lib = _lib, _libs[self.name]
        return self

    def __exit__(self, *exc):
        global _lib
        _lib = self._prev


def noise_floor(run):
    """Rounding-noise floor of an oracle computation: ``run()`` must return an
    array computed with the oracle; it is evaluated with the strict and with the fast-math
    build and the relative L2 difference is returned together with the strict result."""
    strict = np.asarray(run())
    with variant('fast'):
        fast = np.asarray(run())
    return float(np.linalg.norm(fast - strict) / np.linalg.norm(strict)), strict


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f(a, dtype):
    """Check that `a` is an F-contiguous (or 1-D contiguous) array of dtype."""
    if a.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {a.dtype}")
    if not (a.flags.f_contiguous or a.flags.c_contiguous and a.ndim == 1):
        raise ValueError("array must be Fortran-contiguous")
    return a


def _suffix(dtype):
    return '_c' if np.dtype(dtype) == np.complex128 else '_r'


def _h(h):
    return np.ascontiguousarray(h, dtype=np.float64)


def amat_x(rx, ry, rz, ex, ey, ez, eta_x, eta_y, eta_z, zeta, hx, hy, hz):
    """r -= A e; signature of emg3d.core.amat_x (core.py:57-58)."""
    dt = ex.dtype
    hx, hy, hz = _h(hx), _h(hy), _h(hz)
    fn = getattr(lib(), 'orc_amat_x' + _suffix(dt))
    eta = [np.asfortranarray(a, dtype=dt) for a in (eta_x, eta_y, eta_z)]
    fn(_p(_f(rx, dt)), _p(_f(ry, dt)), _p(_f(rz, dt)),
       _p(_f(ex, dt)), _p(_f(ey, dt)), _p(_f(ez, dt)),
       _p(eta[0]), _p(eta[1]), _p(eta[2]),
       _p(_f(np.asfortranarray(zeta), np.float64)), _p(hx), _p(hy), _p(hz),
       ctypes.c_int(hx.size), ctypes.c_int(hy.size), ctypes.c_int(hz.size))


def _gs(ldir, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu):
    dt = ex.dtype
    hx, hy, hz = _h(hx), _h(hy), _h(hz)
    fn = getattr(lib(), 'orc_gauss_seidel' + _suffix(dt))
    eta = [np.asfortranarray(a, dtype=dt) for a in (eta_x, eta_y, eta_z)]
    fn(ctypes.c_int(ldir),
       _p(_f(ex, dt)), _p(_f(ey, dt)), _p(_f(ez, dt)),
       _p(_f(sx, dt)), _p(_f(sy, dt)), _p(_f(sz, dt)),
       _p(eta[0]), _p(eta[1]), _p(eta[2]),
       _p(_f(np.asfortranarray(zeta), np.float64)), _p(hx), _p(hy), _p(hz),
       ctypes.c_int(hx.size), ctypes.c_int(hy.size), ctypes.c_int(hz.size),
       ctypes.c_int(int(nu)))


def gauss_seidel(*args):
    """Point smoother; signature of emg3d.core.gauss_seidel (core.py:210-212)."""
    _gs(0, *args)


def gauss_seidel_x(*args):
    """x-line smoother; emg3d.core.gauss_seidel_x (core.py:506-508)."""
    _gs(1, *args)


def gauss_seidel_y(*args):
    """y-line smoother; emg3d.core.gauss_seidel_y (core.py:786-788)."""
    _gs(2, *args)


def gauss_seidel_z(*args):
    """z-line smoother; emg3d.core.gauss_seidel_z (core.py:1071-1073)."""
    _gs(3, *args)


def solve(amat, bvec):
    """Banded LDL^T solve in place; emg3d.core.solve (core.py:1481-1482)."""
    dt = bvec.dtype
    if amat.dtype != dt:
        raise TypeError("amat and bvec must have the same dtype")
    fn = getattr(lib(), 'orc_solve' + _suffix(dt))
    fn(_p(_f(amat, dt)), _p(_f(bvec, dt)), ctypes.c_int(bvec.size))


# sc_dir -> which axes are coarsened (solver.py:891-897)
SC_FLAGS = {0: (1, 1, 1), 1: (0, 1, 1), 2: (1, 0, 1), 3: (1, 1, 0),
            4: (1, 0, 0), 5: (0, 1, 0), 6: (0, 0, 1)}


def restrict(crx, cry, crz, rx, ry, rz, wx, wy, wz, sc_dir):
    """Full-weighting restriction; emg3d.core.restrict (core.py:1620-1621)."""
    dt = rx.dtype
    fn = getattr(lib(), 'orc_restrict' + _suffix(dt))
    nx, ny, nz = ry.shape[0] - 1, rx.shape[1] - 1, rx.shape[2] - 1
    flags = (ctypes.c_int * 3)(*SC_FLAGS[int(sc_dir)])
    w = [np.ascontiguousarray(a, dtype=np.float64) for t in (wx, wy, wz) for a in t]
    fn(_p(_f(crx, dt)), _p(_f(cry, dt)), _p(_f(crz, dt)),
       _p(_f(rx, dt)), _p(_f(ry, dt)), _p(_f(rz, dt)),
       *[_p(a) for a in w],
       ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(nz), flags)


def restrict_weights(nodes, cell_centers, h, cnodes, ccell_centers, ch):
    """1-D restriction weights; emg3d.core.restrict_weights (core.py:2004-2005)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64)
            for a in (nodes, cell_centers, h, cnodes, ccell_centers, ch)]
    n = arrs[3].size
    wl, w0, wr = np.empty(n), np.empty(n), np.empty(n)
    lib().orc_restrict_weights(*[_p(a) for a in arrs], ctypes.c_int(n),
                               _p(wl), _p(w0), _p(wr))
    return wl, w0, wr


def prolong(ex, ey, ez, cex, cey, cez, nodes, cnodes, sc_dir):
    """e += P ce on interior edges; emg3d.solver.prolongation (solver.py:947-1019).

    ``nodes``/``cnodes`` are (x, y, z) tuples of fine/coarse node coordinates.
    """
    dt = ex.dtype
    fn = getattr(lib(), 'orc_prolong' + _suffix(dt))
    nx, ny, nz = ey.shape[0] - 1, ex.shape[1] - 1, ex.shape[2] - 1
    flags = (ctypes.c_int * 3)(*SC_FLAGS[int(sc_dir)])
    xs = [np.ascontiguousarray(a, dtype=np.float64) for a in (*nodes, *cnodes)]
    fn(_p(_f(ex, dt)), _p(_f(ey, dt)), _p(_f(ez, dt)),
       _p(_f(cex, dt)), _p(_f(cey, dt)), _p(_f(cez, dt)),
       *[_p(a) for a in xs],
       ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(nz), flags)


def edge_curl_factor(mx, my, mz, ex, ey, ez, hx, hy, hz, zeta):
    """NumPy restatement of emg3d/fields.py:941-1009 `_edge_curl_factor`: curl of the
    edge field through every face that is not on the low boundary, times the sum of
    ``zeta`` over the two cells sharing the face, over (sum of the two widths across
    the face) x (face area); results into mx (nx+1, ny, nz), my, mz (other faces stay)."""
    hx, hy, hz = _h(hx), _h(hy), _h(hz)
    nx, ny, nz = hx.size, hy.size, hz.size
    X, Y, Z = hx[:, None, None], hy[None, :, None], hz[None, None, :]
    fx = (ez[:nx, 1:, :] - ez[:nx, :ny, :]) / Y - (ey[:nx, :, 1:] - ey[:nx, :, :nz]) / Z
    fy = (ex[:, :ny, 1:] - ex[:, :ny, :nz]) / Z - (ez[1:, :ny, :] - ez[:nx, :ny, :]) / X
    fz = (ey[1:, :, :nz] - ey[:nx, :, :nz]) / X - (ex[:, 1:, :nz] - ex[:, :ny, :nz]) / Y
    z = np.asarray(zeta)
    mx[1:nx] = (fx[1:] * (z[:-1] + z[1:]) / ((hx[:-1] + hx[1:])[:, None, None] * Y * Z))
    my[:, 1:ny] = (fy[:, 1:] * (z[:, :-1] + z[:, 1:]) / (X * (hy[:-1] + hy[1:])[None, :, None] * Z))
    mz[:, :, 1:nz] = (fz[:, :, 1:] * (z[:, :, :-1] + z[:, :, 1:])
                      / (X * Y * (hz[:-1] + hz[1:])[None, None, :]))


def gs_sequence(ldir, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, seq):
    """Relax the blocks listed in ``seq`` one after the other (test helper).

    ``seq``: int array of (ix, iy, iz) rows for ``ldir = 0`` or (t1, t2) rows
    (transverse node indices in the reference's slot order) for line smoothers.
    """
    dt = ex.dtype
    hx, hy, hz = _h(hx), _h(hy), _h(hz)
    fn = getattr(lib(), 'orc_gs_sequence' + _suffix(dt))
    eta = [np.asfortranarray(a, dtype=dt) for a in (eta_x, eta_y, eta_z)]
    seq = np.ascontiguousarray(seq, dtype=np.int32)
    fn(ctypes.c_int(ldir),
       _p(_f(ex, dt)), _p(_f(ey, dt)), _p(_f(ez, dt)),
       _p(_f(sx, dt)), _p(_f(sy, dt)), _p(_f(sz, dt)),
       _p(eta[0]), _p(eta[1]), _p(eta[2]),
       _p(_f(np.asfortranarray(zeta), np.float64)), _p(hx), _p(hy), _p(hz),
       ctypes.c_int(hx.size), ctypes.c_int(hy.size), ctypes.c_int(hz.size),
       _p(seq), ctypes.c_int(seq.shape[0]))


# schedule constants of the CUDA point smoother (csrc/kernels.h, csrc/gs_point.cu)
TILE = (64, 4, 4)            # default; tests pass the library's emg3d_b200_point_tile_shape
TILE_MIN_NODES = 300000


def color_sequence(ldir, shape, nu, tile_variant=0, tile=None):
    """Block sequence of ``nu`` multicolour sweeps as the CUDA kernels run them.

    Point smoother: 8 parity classes of (ix-1, iy-1, iz-1), class index
    ``px + 2 py + 4 pz``; line smoothers: 4 classes ``pp + 2 pq`` of the
    transverse node indices (``tile_variant`` / ``tile``: node order inside a tile and
    tile shape of the tile-fused point smoother, see emg3d_b200_point_tile_schedule /
    _shape in the C header).
    Odd sweeps run the classes in descending order,
    even sweeps ascending (the first sweep of the reference is the descending
    one, core.py:301, 311).
    """
    rows = []
    back = False
    tiled = ldir == 0 and (shape[0] - 1) * (shape[1] - 1) * (shape[2] - 1) > TILE_MIN_NODES
    for _ in range(nu):
        back = not back
        if tiled:
            # large grids (csrc/gs_point.cu, tile-fused schedule): tiles of
            # TILE nodes coloured by tile-index parity; per tile colour every
            # tile runs its 8 node colours; both orders reverse on odd sweeps
            nx, ny, nz = shape
            tx, ty, tz = tile or TILE
            ntile = [-(-(n - 1) // t) for n, t in zip(shape, (tx, ty, tz))]
            classes = list(range(7, -1, -1) if back else range(8))
            if tile_variant == 1:
                # y-marching tiles: per tile colour, 4 column colours (parity of ix,
                # iz); a column is relaxed node after node along y (descending on
                # odd sweeps); columns of one colour are independent
                for tc in classes:
                    for cc in (range(3, -1, -1) if back else range(4)):
                        px, pz = cc & 1, cc >> 1
                        for kz in range((tc >> 2) & 1, ntile[2], 2):
                            for ky in range((tc >> 1) & 1, ntile[1], 2):
                                for kx in range(tc & 1, ntile[0], 2):
                                    ys = list(range(1 + ky * ty, min(ny, 1 + (ky + 1) * ty)))
                                    if back:
                                        ys.reverse()
                                    for iz in range(1 + kz * tz + pz, min(nz, 1 + (kz + 1) * tz), 2):
                                        for ix in range(1 + kx * tx + px, min(nx, 1 + (kx + 1) * tx), 2):
                                            for iy in ys:
                                                rows.append((ix, iy, iz))
                continue
            for tc in classes:
                for c in classes:
                    px, py, pz = c & 1, (c >> 1) & 1, (c >> 2) & 1
                    for kz in range((tc >> 2) & 1, ntile[2], 2):
                        for ky in range((tc >> 1) & 1, ntile[1], 2):
                            for kx in range(tc & 1, ntile[0], 2):
                                for iz in range(1 + kz * tz + pz, min(nz, 1 + (kz + 1) * tz), 2):
                                    for iy in range(1 + ky * ty + py, min(ny, 1 + (ky + 1) * ty), 2):
                                        for ix in range(1 + kx * tx + px, min(nx, 1 + (kx + 1) * tx), 2):
                                            rows.append((ix, iy, iz))
        elif ldir == 0:
            nx, ny, nz = shape
            classes = range(7, -1, -1) if back else range(8)
            for c in classes:
                px, py, pz = c & 1, (c >> 1) & 1, (c >> 2) & 1
                for iz in range(1 + pz, nz, 2):
                    for iy in range(1 + py, ny, 2):
                        for ix in range(1 + px, nx, 2):
                            rows.append((ix, iy, iz))
        else:
            d = ldir - 1
            t1, t2 = (1 if d == 0 else 0), (1 if d == 2 else 2)
            classes = range(3, -1, -1) if back else range(4)
            for c in classes:
                pp, pq = c & 1, c >> 1
                for b in range(1 + pq, shape[t2], 2):
                    for a in range(1 + pp, shape[t1], 2):
                        rows.append((a, b))
    return np.array(rows, dtype=np.int32).reshape(-1, 3 if ldir == 0 else 2)

"""Array-level interface with the signatures of ``emg3d.core``.

These are the functions ``emg3d/solver.py`` calls; each
takes host NumPy arrays, runs the CUDA kernel through the C ABI and writes the
result back into the caller's arrays, so ``emg3d.solver`` can be pointed at this
module (``solver.core = emg3d_b200.core``) as a literal drop-in.  Inside
:mod:`emg3d_b200.solver` the same kernels are driven on device-resident data
instead.  There is no CPU fallback: without the CUDA library and a GPU every
call raises :class:`emg3d_b200.Emg3dB200Error`.

``ORDER`` selects the Gauss-Seidel ordering of the smoothers: ``'lex'`` is
sequentially equivalent to the reference's lexicographic sweeps, ``'color'`` is
the multicolour ordering used for throughput.
"""
import os
from ctypes import c_void_p

import numpy as np

from emg3d_b200 import _lib

__all__ = ['amat_x', 'gauss_seidel', 'gauss_seidel_x', 'gauss_seidel_y',
           'gauss_seidel_z', 'blocks_to_amat', 'solve', 'restrict',
           'restrict_weights', 'edge_curl_factor']

ORDER = os.environ.get('EMG3D_B200_ORDER', 'color')
_ORDERS = {'lex': _lib.ORDER_LEX, 'color': _lib.ORDER_COLOR}
# sc_dir -> which axes are coarsened (emg3d/solver.py:891-897)
SC_FLAGS = {0: (1, 1, 1), 1: (0, 1, 1), 2: (1, 0, 1), 3: (1, 1, 0),
            4: (1, 0, 0), 5: (0, 1, 0), 6: (0, 0, 1)}


def order_id(order=None):
    order = ORDER if order is None else order
    if order not in _ORDERS:
        raise ValueError(f"`order` must be 'lex' or 'color'; provided: {order!r}.")
    return _ORDERS[order]


def _farr(a, dtype, name, writable=False):
    """The reference passes Fortran-ordered views; require the same."""
    a = np.asarray(a)
    if a.dtype != dtype:
        if writable:
            raise TypeError(f"`{name}` must have dtype {dtype}, got {a.dtype}")
        a = a.astype(dtype)
    if not a.flags.f_contiguous:
        if writable:
            raise ValueError(f"`{name}` must be Fortran-contiguous")
        a = np.asfortranarray(a)
    return a


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _common(ex, eta_x, eta_y, eta_z, zeta, hx, hy, hz):
    dt = np.dtype(ex.dtype)
    if dt not in (np.dtype(np.complex128), np.dtype(np.float64)):
        raise TypeError(f"fields must be complex128 or float64, got {dt}")
    hx, hy, hz = (np.ascontiguousarray(h, dtype=np.float64) for h in (hx, hy, hz))
    ex_ = _farr(eta_x, dt, 'eta_x')
    ey_ = ex_ if eta_y is eta_x else _farr(eta_y, dt, 'eta_y')
    ez_ = ex_ if eta_z is eta_x else _farr(eta_z, dt, 'eta_z')
    zt = _farr(zeta, np.float64, 'zeta')
    return dt, (ex_, ey_, ez_, zt), (hx, hy, hz)


def amat_x(rx, ry, rz, ex, ey, ez, eta_x, eta_y, eta_z, zeta, hx, hy, hz):
    """``r -= A e`` in place (emg3d.core.amat_x, core.py:57-206)."""
    dt, (a, b, c, zt), (hx, hy, hz) = _common(ex, eta_x, eta_y, eta_z, zeta, hx, hy, hz)
    r = [_farr(v, dt, n, True) for v, n in ((rx, 'rx'), (ry, 'ry'), (rz, 'rz'))]
    e = [_farr(v, dt, n) for v, n in ((ex, 'ex'), (ey, 'ey'), (ez, 'ez'))]
    _lib.check(_lib.init().emg3d_b200_host_amat_x(
        int(dt.kind == 'c'), hx.size, hy.size, hz.size, _p(r[0]), _p(r[1]), _p(r[2]),
        _p(e[0]), _p(e[1]), _p(e[2]), _p(a), _p(b), _p(c), _p(zt), _p(hx), _p(hy), _p(hz)))


def edge_curl_factor(mx, my, mz, ex, ey, ez, hx, hy, hz, zeta):
    """Curl of an edge field on the faces, times a two-cell average of ``zeta`` over the
    dual-cell measure, into ``mx, my, mz`` (emg3d.fields._edge_curl_factor,
    fields.py:941-1009; not in ``emg3d.core`` but the same stencil family as amat_x).
    Faces on the grid boundary are set to zero (the reference leaves them untouched
    in a pre-zeroed field)."""
    dt = np.dtype(ex.dtype)
    if dt not in (np.dtype(np.complex128), np.dtype(np.float64)):
        raise TypeError(f"fields must be complex128 or float64, got {dt}")
    hx, hy, hz = (np.ascontiguousarray(h, dtype=np.float64) for h in (hx, hy, hz))
    m = [_farr(v, dt, n, True) for v, n in ((mx, 'mx'), (my, 'my'), (mz, 'mz'))]
    e = [_farr(v, dt, n) for v, n in ((ex, 'ex'), (ey, 'ey'), (ez, 'ez'))]
    zt = _farr(zeta, dt, 'zeta')
    _lib.check(_lib.init().emg3d_b200_host_edge_curl_factor(
        int(dt.kind == 'c'), hx.size, hy.size, hz.size, _p(m[0]), _p(m[1]), _p(m[2]),
        _p(e[0]), _p(e[1]), _p(e[2]), _p(hx), _p(hy), _p(hz), _p(zt)))


def _gs(ldir, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu,
        order=None):
    dt, (a, b, c, zt), (hx, hy, hz) = _common(ex, eta_x, eta_y, eta_z, zeta, hx, hy, hz)
    e = [_farr(v, dt, n, True) for v, n in ((ex, 'ex'), (ey, 'ey'), (ez, 'ez'))]
    s = [_farr(v, dt, n) for v, n in ((sx, 'sx'), (sy, 'sy'), (sz, 'sz'))]
    _lib.check(_lib.init().emg3d_b200_host_gauss_seidel(
        int(dt.kind == 'c'), ldir, order_id(order), hx.size, hy.size, hz.size,
        _p(e[0]), _p(e[1]), _p(e[2]), _p(s[0]), _p(s[1]), _p(s[2]),
        _p(a), _p(b), _p(c), _p(zt), _p(hx), _p(hy), _p(hz), int(nu)))


def gauss_seidel(ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu,
                 order=None):
    """Point-block smoother, ``nu`` sweeps in place (core.py:210-503)."""
    _gs(0, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu, order)


def gauss_seidel_x(ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu,
                   order=None):
    """x-line relaxation (core.py:506-783)."""
    _gs(1, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu, order)


def gauss_seidel_y(ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu,
                   order=None):
    """y-line relaxation (core.py:786-1068)."""
    _gs(2, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu, order)


def gauss_seidel_z(ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu,
                   order=None):
    """z-line relaxation (core.py:1071-1348)."""
    _gs(3, ex, ey, ez, sx, sy, sz, eta_x, eta_y, eta_z, zeta, hx, hy, hz, nu, order)


def restrict_weights(nodes, cell_centers, h, cnodes, ccell_centers, ch):
    """1-D restriction weights (core.py:2004-2076; Muld06 Eq. 9).

    O(n) host work, evaluated once per level and axis.  With ``d`` the
    dual-cell widths (half cells at both ends) the left/right weights are the
    distance between neighbouring fine and coarse cell centres times ``1/d``.
    """
    nodes, cell_centers, h, cnodes, ccell_centers, ch = (
        np.asarray(a, dtype=np.float64)
        for a in (nodes, cell_centers, h, cnodes, ccell_centers, ch))
    n = cnodes.size
    dual = np.empty(n + 1)
    dual[0], dual[-1] = h[0] / 2, h[-1] / 2
    dual[1:-1] = (h[0:2 * n - 3:2] + h[1:2 * n - 2:2]) / 2.
    dist_l = np.empty(n)
    dist_l[0] = (nodes[0] - h[0] / 2) - (cnodes[0] - ch[0] / 2)
    dist_l[1:] = cell_centers[1::2] - ccell_centers
    dist_r = np.empty(n)
    dist_r[-1] = (cnodes[-1] + ch[-1] / 2) - (nodes[-1] + h[-1] / 2)
    dist_r[:-1] = ccell_centers - cell_centers[0::2]
    wl = (1 / dual[:-1]) * dist_l
    wr = (1 / dual[1:]) * dist_r
    return wl, np.ones(n), wr


def interpolation_table(nodes, cnodes):
    """Per fine node: lower coarse node and fraction inside that coarse cell.

    The 1-D factors of the bilinear weights of the reference's
    ``RegularGridProlongator`` (emg3d/solver.py:1447-1473).
    """
    nodes, cnodes = np.asarray(nodes, float), np.asarray(cnodes, float)
    lo = np.clip(np.searchsorted(cnodes, nodes) - 1, 0, cnodes.size - 2)
    frac = (nodes - cnodes[lo]) / (cnodes[lo + 1] - cnodes[lo])
    return lo.astype(np.int32), frac


def restrict(crx, cry, crz, rx, ry, rz, wx, wy, wz, sc_dir):
    """Restrict the fine residual to the coarse grid (core.py:1620-2001), host arrays in and out
    through ``emg3d_b200_host_restrict``."""
    dt = np.dtype(rx.dtype)
    r = [_farr(v, dt, n) for v, n in ((rx, 'rx'), (ry, 'ry'), (rz, 'rz'))]
    cr = [_farr(v, dt, n, True) for v, n in ((crx, 'crx'), (cry, 'cry'), (crz, 'crz'))]
    fshape = (r[1].shape[0] - 1, r[0].shape[1] - 1, r[0].shape[2] - 1)
    if int(sc_dir) not in SC_FLAGS:
        raise ValueError(f"sc_dir must be 0 .. 6; provided: {sc_dir!r}.")
    cflag = SC_FLAGS[int(sc_dir)]
    cshape = tuple(n // 2 if f else n for n, f in zip(fshape, cflag))
    want = ((cshape[0], cshape[1] + 1, cshape[2] + 1), (cshape[0] + 1, cshape[1], cshape[2] + 1),
            (cshape[0] + 1, cshape[1] + 1, cshape[2]))
    for v, shp, name in zip(cr, want, ('crx', 'cry', 'crz')):
        if v.shape != shp:
            raise ValueError(f"{name}: shape {v.shape}, expected {shp} for sc_dir={sc_dir}.")
    keep, wp = [], (_lib.c_void_p * 9)()
    for a, w in enumerate((wx, wy, wz)):
        for k in range(3):
            if cflag[a]:
                arr = np.ascontiguousarray(w[k], dtype=np.float64)
                if arr.size != cshape[a] + 1:
                    raise ValueError(f"restriction weights of axis {a}: length {arr.size}, "
                                     f"expected {cshape[a] + 1}.")
                keep.append(arr)
                wp[3 * a + k] = arr.ctypes.data
            else:
                wp[3 * a + k] = None
    _lib.check(_lib.load().emg3d_b200_host_restrict(
        int(dt.kind == 'c'), *fshape, int(sc_dir), *[_p(v) for v in cr], *[_p(v) for v in r], wp))


def blocks_to_amat(amat, bvec, middle, left, rhs, im, nc):
    """Scatter one 5x5 block row into band storage (core.py:1351-1477).

    Host-side data movement kept for interface completeness; on the device the
    line systems are never assembled in band form (see csrc/gs_line.cu).
    ``amat[p + 5 q]`` holds A(p, q) for 0 <= p - q <= 5.
    """
    row0 = 5 * im
    nrows = 5 if im < nc - 1 else 1
    if im == nc - 1 and nc <= 1:
        raise ValueError("a line needs at least two cells")
    for k in range(nrows):
        bvec[row0 + k] = rhs[k]
        for m in range(k + 1):                       # within the block
            amat[(row0 + k) + 5 * (row0 + m)] = middle[k + 5 * m]
    if im > 0:
        col0 = row0 - 5
        for m in range(1, 5):                        # coupling to the previous block
            for k in range(min(m, nrows - 1) + 1):
                amat[(row0 + k) + 5 * (col0 + m)] = left[k + 5 * m]


def solve(amat, bvec):
    """Banded complex-symmetric LDL^T solve in place (core.py:1481-1616)."""
    dt = np.dtype(bvec.dtype)
    amat = _farr(amat, dt, 'amat', True)
    bvec = _farr(bvec, dt, 'bvec', True)
    _lib.check(_lib.init().emg3d_b200_host_solve(int(dt.kind == 'c'), int(bvec.size),
                                                 _p(amat), _p(bvec)))

"""CPU restatement (NumPy) of the interpolation kernels next to the solve -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this package;
the product (``emg3d_b200``) never does.

What is restated, and from where:

* ``volume_average_weights`` / ``interp_volume_average``: emg3d/maps.py:620-665 and 556-617
  (the numba kernels behind ``maps.interpolate(method='volume')``);
* ``interp_edges_to_vol_averages``: emg3d/maps.py:668-720 (used by ``Simulation.gradient``,
  emg3d/simulations.py:1040-1046);
* ``spline_filter3`` / ``map_coordinates3``: cubic-spline interpolation of
  ``scipy.ndimage.map_coordinates(order=3)`` (SciPy 1.18, ``ni_splines.c`` / ``ni_interpolation.c``),
  which the reference calls through ``maps.interp_spline_3d`` (emg3d/maps.py:500-553) for
  ``get_receiver`` and ``Field.interpolate_to_grid`` (emg3d/fields.py:303-346, 522-615): B-spline
  prefilter with mirror boundaries (pole sqrt(3) - 2, gain 6), then a 4 x 4 x 4 weighted sum with
  mirrored support; ``mode='constant'`` (points outside the data: ``cval``) and
  ``mode='nearest'`` (input padded by 12 edge samples, reflect boundaries).  SciPy is a
  third-party dependency of the reference (absent from /root/reference, present in this
  image): the restatement is pinned by running SciPy itself (tests/test_oracle_golden.py).
"""
import numpy as np

POLE = np.sqrt(3.0) - 2.0


def volume_average_weights(x_i, x_o):
    """Weights and index maps of one axis (maps.py:620-665)."""
    x_i, x_o = np.asarray(x_i, float), np.asarray(x_o, float)
    xs = np.unique(np.concatenate((x_i, x_o)))
    n1, n2 = len(x_i), len(x_o)
    center = 0.5 * (xs[:-1] + xs[1:])
    keep = (x_o[0] <= center) & (center <= x_o[-1])
    c = center[keep]
    w = (xs[1:] - xs[:-1])[keep]
    # while i < n - 1 and center >= x[i]: i += 1  ->  i = min(#{x <= center}, n - 1)
    i1 = np.minimum(np.searchsorted(x_i, c, side='right'), n1 - 1)
    i2 = np.minimum(np.searchsorted(x_o, c, side='right'), n2 - 1)
    return w, np.clip(i1 - 1, 0, n1 - 1).astype(np.int32), np.clip(i2 - 1, 0, n2 - 1).astype(np.int32)


def interp_volume_average(nodes, values, new_nodes, new_vol):
    """maps.py:556-617: returns the new values (the reference adds into a zero array)."""
    out = np.zeros(new_vol.shape)
    (wx, ix, ox), (wy, iy, oy), (wz, iz, oz) = (volume_average_weights(a, b) for a, b in zip(nodes, new_nodes))
    v = values[np.ix_(ix, iy, iz)] * wx[:, None, None] * wy[None, :, None] * wz[None, None, :]
    # sum the merged segments of each output cell, axis after axis
    for ax, (o, n) in enumerate(zip((ox, oy, oz), new_vol.shape)):
        acc = np.zeros(v.shape[:ax] + (n,) + v.shape[ax + 1:])
        np.add.at(acc, (slice(None),) * ax + (o,), v)
        v = acc
    out += v
    return out / new_vol


def interp_edges_to_vol_averages(ex, ey, ez, volumes):
    """maps.py:668-720: returns (ox, oy, oz)."""
    nx, ny, nz = volumes.shape

    def pair(n):       # cells an edge row j = 0 .. n adds to: max(0, j - 1) and min(n - 1, j)
        j = np.arange(n + 1)
        return np.maximum(j - 1, 0), np.minimum(j, n - 1)

    out = []
    for comp, e in enumerate((ex, ey, ez)):
        o = np.zeros(volumes.shape, dtype=e.dtype)
        axes = [a for a in range(3) if a != comp]
        pa, pb = pair(volumes.shape[axes[0]]), pair(volumes.shape[axes[1]])
        for sa in pa:
            for sb in pb:
                idx = [slice(None)] * 3
                idx[axes[0]], idx[axes[1]] = sa, sb
                ii = np.ix_(*[np.arange(volumes.shape[a]) if isinstance(idx[a], slice) else idx[a]
                              for a in range(3)])
                np.add.at(o, ii, volumes[ii] * e / 4)
        out.append(o)
    return out


# ---- cubic B-spline interpolation (scipy.ndimage.map_coordinates, order 3) ----------------------

def _filter_line(c, mode):
    """In-place prefilter of the lines along axis 0 of c (n, ...)."""
    n = c.shape[0]
    if n < 2:
        return
    z = POLE
    c *= (1.0 - z) * (1.0 - 1.0 / z)                     # gain (= 6)
    if mode == 'mirror':
        # causal start: sum of the mirrored sequence, closed form over one period
        zn1 = z ** (n - 1)
        c0 = c[0] + zn1 * c[n - 1]
        zi = z
        for i in range(1, n - 1):
            c0 = c0 + zi * (c[i] + zn1 * c[n - 1 - i])
            zi *= z
        c[0] = c0 / (1.0 - zn1 * zn1)
    else:                                                # 'reflect' (half-sample symmetric)
        zi = z
        zn = z ** n
        c0 = c[0] + zn * c[n - 1]
        for i in range(1, n):
            c0 = c0 + zi * (c[i] + zn * c[n - 1 - i])
            zi *= z
        c[0] = c[0] + c0 * z / (1.0 - zn * zn)
    for i in range(1, n):
        c[i] += z * c[i - 1]
    if mode == 'mirror':
        c[n - 1] = (z / (z * z - 1.0)) * (c[n - 1] + z * c[n - 2])
    else:
        c[n - 1] *= z / (z - 1.0)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])


def spline_filter3(data, mode='mirror'):
    """Cubic B-spline coefficients of a 3-D array (float or complex), all axes."""
    c = np.array(data, dtype=complex if np.iscomplexobj(data) else float, order='C')
    for ax in range(3):
        v = np.moveaxis(c, ax, 0)
        _filter_line(v, mode)
    return c


def _weights(x):
    w1 = (x * x * (x - 2.0) * 3.0 + 4.0) / 6.0
    zc = 1.0 - x
    w2 = (zc * zc * (zc - 2.0) * 3.0 + 4.0) / 6.0
    w0 = zc * zc * zc / 6.0
    return w0, w1, w2, 1.0 - w0 - w1 - w2


def map_coordinates3(data, coords, mode='constant', cval=0.0):
    """``scipy.ndimage.map_coordinates(data, coords, order=3, mode=mode, cval=cval)`` for a 3-D
    array; coords (3, npts) in index units."""
    data = np.asarray(data)
    coords = np.asarray(coords, float)
    if mode == 'nearest':
        npad = 12
        coef = spline_filter3(np.pad(data, npad, mode='edge'), 'reflect')
    elif mode == 'constant':
        npad = 0
        coef = spline_filter3(data, 'mirror')
    else:
        raise ValueError(mode)
    out = np.empty(coords.shape[1], dtype=coef.dtype)
    shape = coef.shape
    for p in range(coords.shape[1]):
        idx, wts, const = [], [], False
        for d in range(3):
            cc = coords[d, p]
            n = data.shape[d]
            if mode == 'constant':
                if cc < 0 or cc > n - 1:
                    const = True
                    break
            cc += npad
            if mode == 'nearest':                        # clamp to the PADDED array
                cc = min(max(cc, 0.0), shape[d] - 1.0)
            f = int(np.floor(cc))
            wts.append(_weights(cc - f))
            ii = np.arange(f - 1, f + 3)
            m = shape[d]
            if mode == 'constant':                       # mirror:  -k -> k,  m-1+k -> m-1-k
                ii = np.where(ii < 0, -ii, ii)
                ii = np.where(ii > m - 1, 2 * (m - 1) - ii, ii)
            else:                                        # (never leaves the padded array)
                ii = np.clip(ii, 0, m - 1)
            idx.append(ii)
        if const:
            out[p] = cval
            continue
        blk = coef[np.ix_(*idx)]
        out[p] = np.einsum('i,j,k,ijk->', np.array(wts[0]), np.array(wts[1]), np.array(wts[2]), blk)
    return out

"""Interpolation on either side of a solve: host-side mirror of the
functions of ``emg3d/maps.py`` that sit next to the hot path, backed by the CUDA kernels of
``csrc/interp.cu`` (through the C ABI, ``include/emg3d_b200.h``).

=========================================  ======================================================
reference (emg3d/maps.py)                  here
=========================================  ======================================================
``interpolate`` (232-369)                  :func:`interpolate`, same arguments; methods
                                           ``'volume'``, ``'cubic'``, ``'linear'``, ``'nearest'``
                                           on the device
``interp_spline_3d`` (500-553)             :func:`interp_spline_3d`
``interp_volume_average`` (556-617)        :func:`interp_volume_average`, same argument list
``_volume_average_weights`` (620-665)      :func:`_volume_average_weights` (O(n) per axis, NumPy)
``interp_edges_to_vol_averages`` (668-720) :func:`interp_edges_to_vol_averages`, same argument list
=========================================  ======================================================

``values`` may be host arrays (uploaded) or :class:`emg3d_b200._lib.DeviceArray` views of data
that already live on the device (the field of a solve: :func:`sample_points` is what
``solve(..., receivers=)`` and ``get_receiver`` run, so that only the responses cross PCIe).
The per-axis bookkeeping (merged-node weights; the index coordinates of the cubic spline, which
the reference obtains with ``scipy.interpolate.interp1d(kind='cubic')``) is O(points per axis)
host work; everything O(cells) or O(points) runs on the GPU.  No CPU fallback.
"""
import ctypes

import numpy as np

from emg3d_b200 import _lib

__all__ = ['interpolate', 'interp_spline_3d', 'interp_volume_average', 'interp_edges_to_vol_averages']


# ---------------------------------------------------------------------------------------------
# volume averaging
# ---------------------------------------------------------------------------------------------

def _volume_average_weights(x_i, x_o):
    """Weights and index maps of one axis for the volume averaging (maps.py:620-665).

    The nodes of both grids are merged; every merged interval whose centre lies inside the
    output grid contributes its length ``hs`` from input cell ``ix_i`` to output cell ``ix_o``
    (cells found by bisection instead of the reference's running counters; same result).
    """
    x_i, x_o = np.asarray(x_i, dtype=float), np.asarray(x_o, dtype=float)
    merged = np.union1d(x_i, x_o)
    mid = 0.5 * (merged[1:] + merged[:-1])
    inside = (mid >= x_o[0]) & (mid <= x_o[-1])
    mid, hs = mid[inside], np.diff(merged)[inside]
    cell_i = np.clip(np.searchsorted(x_i, mid, side='right'), 1, x_i.size - 1) - 1
    cell_o = np.clip(np.searchsorted(x_o, mid, side='right'), 1, x_o.size - 1) - 1
    return hs, cell_i.astype(np.int32), cell_o.astype(np.int32)


def _axis_tables(x_i, x_o):
    """Device tables of one axis: weights, input cells, segment offsets per output cell, widths."""
    hs, ci, co = _volume_average_weights(x_i, x_o)
    start = np.searchsorted(co, np.arange(len(x_o)), side='left').astype(np.int32)   # co is sorted
    return [_lib.DeviceArray.from_host(a) for a in
            (hs, ci, start, np.diff(np.asarray(x_o, dtype=float)))]


def _ptr_array(arrays):
    return (ctypes.c_void_p * len(arrays))(*[a.ptr for a in arrays])


def _volume_average_device(nodes, d_values, shape, new_nodes, d_out, log=False, add=False):
    tabs = [_axis_tables(a, b) for a, b in zip(nodes, new_nodes)]
    new_shape = [len(b) - 1 for b in new_nodes]
    lib = _lib.load()
    _lib.check(lib.emg3d_b200_volume_average(
        d_values.ptr, *[int(n) for n in shape], d_out.ptr, *new_shape,
        _ptr_array([t[0] for t in tabs]), _ptr_array([t[1] for t in tabs]),
        _ptr_array([t[2] for t in tabs]), _ptr_array([t[3] for t in tabs]), int(bool(log)), int(bool(add))))
    _lib.sync()                                          # (the tables are released on return)


def interp_volume_average(nodes_x, nodes_y, nodes_z, values, new_nodes_x, new_nodes_y, new_nodes_z,
                          new_values, new_vol):
    """Volume-averaging interpolation, argument list of maps.py:556-617: the result is ADDED to
    ``new_values`` and the sum divided by ``new_vol`` (in place)."""
    values = np.asarray(values, dtype=float)
    new_nodes = (new_nodes_x, new_nodes_y, new_nodes_z)
    expect = tuple(len(b) - 1 for b in new_nodes)
    if tuple(new_values.shape) != expect:
        raise ValueError(f"new_values must have shape {expect}; provided: {new_values.shape}.")
    d_val = _lib.DeviceArray.from_host(np.asfortranarray(values))
    d_out = _lib.DeviceArray.from_host(np.asfortranarray(new_values, dtype=float))
    _volume_average_device((nodes_x, nodes_y, nodes_z), d_val, values.shape, new_nodes, d_out, add=True)
    # (the kernel divides by the product of the new widths = new_vol of a tensor mesh; a caller's
    # own new_vol is honoured)
    res = d_out.download().reshape(expect, order='F')
    tensor_vol = (np.diff(new_nodes_x)[:, None, None] * np.diff(new_nodes_y)[None, :, None] *
                  np.diff(new_nodes_z)[None, None, :])
    new_values[...] = res * (tensor_vol / np.asarray(new_vol).reshape(expect, order='F'))


def interp_edges_to_vol_averages(ex, ey, ez, volumes, ox, oy, oz):
    """Edges to volume-weighted cell averages, argument list of maps.py:668-720 (results are added
    to ``ox, oy, oz``).  ``volumes`` must be those of a tensor mesh (products of widths)."""
    nx, ny, nz = volumes.shape
    vol = np.asarray(volumes, dtype=float)
    # widths up to a common factor: the kernel multiplies them back together
    hx = vol[:, 0, 0] / vol[0, 0, 0]
    hy = vol[0, :, 0] / vol[0, 0, 0]
    hz = vol[0, 0, :]
    dtype = np.result_type(ex, ey, ez)
    field = np.concatenate([np.asarray(a, dtype=dtype).ravel('F') for a in (ex, ey, ez)])
    d_f = _lib.DeviceArray.from_host(field)
    d_o = _lib.DeviceArray(3 * vol.size, dtype)
    d_h = [_lib.DeviceArray.from_host(np.ascontiguousarray(h)) for h in (hx, hy, hz)]
    _lib.check(_lib.load().emg3d_b200_edges_to_vol_averages(
        int(dtype.kind == 'c'), nx, ny, nz, d_f.ptr, d_h[0].ptr, d_h[1].ptr, d_h[2].ptr, d_o.ptr))
    res = d_o.download().reshape((nx, ny, nz, 3), order='F')
    for k, o in enumerate((ox, oy, oz)):
        o += res[..., k]


# ---------------------------------------------------------------------------------------------
# point evaluation (cubic spline, linear)
# ---------------------------------------------------------------------------------------------

def _index_coordinates(points, xi, method):
    """Coordinates in index units of the data along one axis.

    cubic: the reference maps them with a cubic spline through (points, 0 .. n-1) with
    extrapolation (maps.py:545-551); linear: cell index + normalised distance as
    RegularGridInterpolator computes them (searchsorted - 1 clipped to 0 .. n-2).
    """
    points, xi = np.asarray(points, dtype=float), np.asarray(xi, dtype=float)
    if method == 'cubic':
        from scipy.interpolate import interp1d
        return interp1d(points, np.arange(points.size), kind='cubic', bounds_error=False,
                        fill_value='extrapolate')(xi)
    i = np.clip(np.searchsorted(points, xi) - 1, 0, max(points.size - 2, 0))
    if points.size == 1:
        return np.zeros_like(xi)
    return i + (xi - points[i]) / (points[i + 1] - points[i])


class _Sampler:
    """Cubic / linear interpolation of ONE device-resident 3-D array at many point sets: the
    spline coefficients are computed once (a copy; the data are not touched)."""

    def __init__(self, d_values, shape, dtype, method, mode='constant', box=None):
        self.shape, self.dtype = tuple(int(n) for n in shape), np.dtype(dtype)
        self.cplx = int(self.dtype.kind == 'c')
        self.method, self.mode = method, mode
        lib = _lib.load()
        self.lo = (0, 0, 0)
        if box is not None and method == 'cubic' and mode == 'constant':
            # only the part of the array the points can see (see _sample_box): a compact copy
            lo, m = box
            if tuple(m) != self.shape:
                sub = _lib.DeviceArray(int(np.prod(m)), self.dtype, scratch=True)
                i3 = ctypes.c_int * 3
                _lib.check(lib.emg3d_b200_copy_box3(self.cplx, *self.shape, d_values.ptr, i3(*lo), i3(*m), sub.ptr))
                d_values, self.shape, self.lo = sub, tuple(int(v) for v in m), tuple(int(v) for v in lo)
        n0, n1, n2 = self.shape
        self.npad = 0
        if method == 'cubic':
            if mode == 'nearest':
                self.npad = 12
                m = [n + 2 * self.npad for n in self.shape]
                self.data = _lib.DeviceArray(int(np.prod(m)), self.dtype, scratch=True)
                _lib.check(lib.emg3d_b200_pad_edge3(self.cplx, n0, n1, n2, d_values.ptr, self.npad,
                                                    self.data.ptr))
                _lib.check(lib.emg3d_b200_spline_filter3(self.cplx, *m, self.data.ptr, 1))
            elif mode == 'constant':
                if tuple(self.shape) != tuple(int(n) for n in shape):
                    self.data = d_values                 # (the compact copy made above: filtered in place)
                else:
                    self.data = _lib.DeviceArray(int(np.prod(self.shape)), self.dtype, scratch=True)
                    _lib.check(lib.emg3d_b200_d2d(self.data.ptr, d_values.ptr, self.data.nbytes))
                _lib.check(lib.emg3d_b200_spline_filter3(self.cplx, n0, n1, n2, self.data.ptr, 0))
            else:
                raise ValueError(f"cubic interpolation: mode must be 'constant' or 'nearest'; provided: {mode!r}.")
        else:
            self.data = d_values

    def __call__(self, coords, d_out, tensor_shape=None, fill=0.0, scale=1.0, accumulate=False):
        """coords: three device arrays of index coordinates (per point, or per axis of a tensor
        grid of ``tensor_shape``); results into ``d_out``."""
        if tensor_shape is None:
            npts, tensor, m0, m1 = coords[0].size, 0, 0, 0
        else:
            npts, tensor, m0, m1 = int(np.prod(tensor_shape)), 1, int(tensor_shape[0]), int(tensor_shape[1])
        fill, scale = complex(fill), complex(scale)
        _lib.check(_lib.load().emg3d_b200_interp_points(
            self.cplx, 3 if self.method == 'cubic' else 1, *self.shape, self.data.ptr, self.npad,
            int(self.mode == 'nearest'), fill.real, fill.imag, coords[0].ptr, coords[1].ptr, coords[2].ptr,
            npts, tensor, m0, m1, scale.real, scale.imag, int(bool(accumulate)), d_out.ptr))


def _device_view(values):
    """(device array, shape, dtype, keep-alive) of host or device-resident 3-D data."""
    if isinstance(values, DeviceView):
        return values.array, values.shape, values.dtype
    values = np.asarray(values)
    if values.dtype.kind not in 'fc':
        values = values.astype(float)
    return _lib.DeviceArray.from_host(np.asfortranarray(values)), values.shape, values.dtype


class DeviceView:
    """A 3-D array (x fastest) inside a device buffer: ``array`` is a :class:`_lib.DeviceArray`
    (or a view object with ``ptr``), ``shape`` its extents."""

    def __init__(self, array, shape, dtype):
        self.array, self.shape, self.dtype = array, tuple(int(n) for n in shape), np.dtype(dtype)


class _PtrView:
    """Non-owning view into a device buffer (``base`` keeps the owner alive)."""

    def __init__(self, base, offset_elems, size, dtype):
        self.base, self.dtype, self.size = base, np.dtype(dtype), int(size)
        self.ptr = base.ptr + int(offset_elems) * self.dtype.itemsize
        self.nbytes = self.size * self.dtype.itemsize


def field_components(d_field, grid, dtype):
    """The three components of a device-resident field ``[fx | fy | fz]`` as :class:`DeviceView`."""
    out, off = [], 0
    for name in 'xyz':
        shape = getattr(grid, 'shape_edges_' + name)
        n = int(np.prod(shape))
        out.append(DeviceView(_PtrView(d_field, off, n, dtype), shape, dtype))
        off += n
    return out


def _points_from_grids(grid, shape, xi, method):
    """Locations of the data and of the requested values (maps.py:371-497): per axis the nodes or
    the cell centres of ``grid``, depending on where an array of ``shape`` lives."""
    shape = tuple(int(n) for n in shape)
    if method == 'volume':
        msg = "``method='volume'`` is only implemented for "
        if not hasattr(xi, 'nodes_x'):
            raise ValueError(msg + "TensorMesh instances as input for ``xi``.")
        if tuple(grid.shape_cells) != shape:
            raise ValueError(msg + f"cell-centered properties; required shape = {grid.shape_cells}.")
    else:
        known = [grid.shape_edges_x, grid.shape_faces_y, grid.shape_edges_z,
                 grid.shape_faces_x, grid.shape_edges_y, grid.shape_faces_z, tuple(grid.shape_cells)]
        if shape not in [tuple(k) for k in known]:
            raise ValueError("``values`` must be a 3D ndarray living on cell centers, "
                             "edges, or faces of the ``grid``.")
    electric = shape not in [tuple(grid.shape_faces_x), tuple(grid.shape_edges_y), tuple(grid.shape_faces_z)]
    is_grid = hasattr(xi, 'nodes_x')
    points, new_points = [], []
    for ax, c in enumerate('xyz'):
        full = [grid.shape_cells[ax], grid.shape_nodes[ax]][electric]
        on_primary = method == 'volume' or shape[ax] == full
        prop = (['cell_centers_', 'nodes_'] if on_primary else ['nodes_', 'cell_centers_'])[electric]
        points.append(getattr(grid, prop + c))
        if is_grid:
            new_points.append(getattr(xi, prop + c))
    if is_grid:
        shape_out = tuple(xi.shape_cells) if method == 'volume' else tuple(len(p) for p in new_points)
        return points, new_points, shape_out, True
    if isinstance(xi, tuple):
        arrs = np.broadcast_arrays(*[np.atleast_1d(np.asarray(a, dtype=float)) for a in xi])
        out_shape = arrs[0].shape
        new_points = [a.ravel() for a in arrs]
    else:
        xi = np.asarray(xi, dtype=float)
        if xi.shape[-1] != 3:
            raise ValueError("The requested sample points xi have dimension "
                             f"{xi.shape[-1]}, but this RegularGridInterpolator has dimension 3")
        out_shape = xi.shape[:-1]
        new_points = [xi[..., k].ravel() for k in range(3)]
    return points, new_points, out_shape, False


_SPLINE_MARGIN = 48      # 0.268^48 = 4e-28: the reach of the B-spline prefilter in double precision


def _sample_box(coords, shape):
    """Index box of an array of ``shape`` that holds everything cubic-spline values at ``coords``
    (index units, per axis) depend on to double precision: their bounding box grown by the reach
    of the prefilter's exponentially decaying impulse response, clipped to the array."""
    lo, m = [], []
    for c, n in zip(coords, shape):
        inside = c[(c >= 0) & (c <= n - 1)]
        if inside.size == 0:
            a, b = 0, min(n, 4)
        else:
            a = max(int(np.floor(inside.min())) - _SPLINE_MARGIN, 0)
            b = min(int(np.ceil(inside.max())) + _SPLINE_MARGIN + 1, n)
        lo.append(a)
        m.append(b - a)
    return tuple(lo), tuple(m)


def sample_points(values, points, new_points, method, mode='constant', fill=0.0, tensor=False,
                  d_out=None, scale=1.0, accumulate=False):
    """Device core of the cubic / linear interpolation: ``values`` (host array or
    :class:`DeviceView`) given at the tensor ``points`` are evaluated at ``new_points`` (three
    arrays: per point, or per axis when ``tensor``).  Returns the device array of results
    (``d_out`` if given: ``d_out = scale * value (+ d_out)``)."""
    d_val, shape, dtype = _device_view(values)
    host_coords = []
    for pts, new, n in zip(points, new_points, shape):
        c = np.asarray(_index_coordinates(pts, new, method), dtype=float)
        if method == 'linear' and fill is not None:          # outside the grid: NaN marks `fill`
            new = np.asarray(new, dtype=float)
            c = np.where((new < pts[0]) | (new > pts[-1]), np.nan, c)
        host_coords.append(c)
    box = None
    if method == 'cubic' and mode == 'constant':
        # a few receivers on a large grid: prefilter only what they can see; points outside the
        # array (-> `fill`) are marked by NaN before the coordinates are shifted into the box
        box = _sample_box(host_coords, shape)
        host_coords = [np.where((c < 0) | (c > n - 1), np.nan, c - lo)
                       for c, n, lo in zip(host_coords, shape, box[0])]
    sampler = _Sampler(d_val, shape, dtype, method, mode, box=box)
    coords = [_lib.DeviceArray.from_host(np.ascontiguousarray(c, dtype=float), scratch=True) for c in host_coords]
    tshape = tuple(len(p) for p in new_points) if tensor else None
    npts = int(np.prod(tshape)) if tensor else coords[0].size
    if d_out is None:
        d_out = _lib.DeviceArray(npts, dtype, scratch=True)
    sampler(coords, d_out, tshape, fill=0.0 if fill is None else fill, scale=scale, accumulate=accumulate)
    _lib.sync()
    return d_out


def interp_spline_3d(points, values, xi, **kwargs):
    """Cubic-spline interpolation in 3-D (maps.py:500-553): ``scipy.ndimage.map_coordinates`` of
    order 3 in the index space of ``points``; keywords ``mode`` ('constant' or 'nearest') and
    ``cval``.  ``xi``: coordinates of shape (npts, 3)."""
    order = kwargs.pop('order', 3)
    mode, cval = kwargs.pop('mode', 'constant'), kwargs.pop('cval', 0.0)
    if order != 3 or kwargs:
        raise NotImplementedError("interp_spline_3d: order 3 with keywords mode / cval only; "
                                  f"provided: order={order}, {sorted(kwargs)}.")
    xi = np.asarray(xi, dtype=float)
    d = sample_points(values, points, [xi[:, 0], xi[:, 1], xi[:, 2]], 'cubic', mode=mode, fill=cval)
    return d.download()


def interpolate(grid, values, xi, method='linear', extrapolate=True, log=False, **kwargs):
    """Interpolate values from one grid to another grid or to points: arguments and semantics of
    maps.py:232-369 (``method``: 'nearest', 'linear', 'volume', 'cubic')."""
    if isinstance(values, DeviceView):
        shape = values.shape
    else:
        values = np.asarray(values)
        shape = values.shape
        if log:
            values = np.log10(values)
    points, new_points, out_shape, is_grid = _points_from_grids(grid, shape, xi, method)

    if method == 'volume':
        d_val, _, _ = _device_view(values)
        d_out = _lib.DeviceArray(int(np.prod(out_shape)), float)
        # (log: the values were already transformed above; the kernel's own log path serves
        # device-resident data)
        _volume_average_device([grid.nodes_x, grid.nodes_y, grid.nodes_z], d_val, shape,
                               [xi.nodes_x, xi.nodes_y, xi.nodes_z], d_out,
                               log=log and isinstance(values, DeviceView))
        res = d_out.download()
    elif method == 'cubic':
        mode = kwargs.pop('mode', 'nearest' if extrapolate else 'constant')
        cval = kwargs.pop('cval', 0.0)
        if kwargs:
            raise NotImplementedError(f"cubic interpolation: unknown keywords {sorted(kwargs)}.")
        res = sample_points(values, points, new_points, 'cubic', mode=mode, fill=cval, tensor=is_grid).download()
    elif method == 'linear':
        bounds_error = kwargs.pop('bounds_error', False)
        fill = kwargs.pop('fill_value', None if extrapolate else 0.0)
        if kwargs:
            raise NotImplementedError(f"linear interpolation: unknown keywords {sorted(kwargs)}.")
        if bounds_error:
            for pts, new in zip(points, new_points):
                if np.any(np.asarray(new) < pts[0]) or np.any(np.asarray(new) > pts[-1]):
                    raise ValueError("One of the requested xi is out of bounds")
        res = sample_points(values, points, new_points, 'linear', fill=fill, tensor=is_grid).download()
    elif method == 'nearest':
        # nearest grid index per axis (O(points) bookkeeping, SciPy's rule: the upper neighbour from
        # a normalised distance > 0.5 on), then a gather on the device: the linear kernel evaluated
        # AT a grid index returns that sample
        fill = kwargs.pop('fill_value', None if extrapolate else 0.0)
        kwargs.pop('bounds_error', None)
        if kwargs:
            raise NotImplementedError(f"nearest interpolation: unknown keywords {sorted(kwargs)}.")
        coords = []
        for pts, new in zip(points, new_points):
            pts, new = np.asarray(pts, float), np.asarray(new, float)
            i = np.clip(np.searchsorted(pts, new) - 1, 0, max(pts.size - 2, 0))
            if pts.size > 1:
                i = i + ((new - pts[i]) / (pts[i + 1] - pts[i]) > 0.5)
            c = np.clip(i, 0, pts.size - 1).astype(float)
            if fill is not None:
                c = np.where((new < pts[0]) | (new > pts[-1]), np.nan, c)
            coords.append(c)
        d_val, vshape, dtype = _device_view(values)
        sampler = _Sampler(d_val, vshape, dtype, 'linear')
        d_c = [_lib.DeviceArray.from_host(np.ascontiguousarray(c), scratch=True) for c in coords]
        npts = int(np.prod(out_shape)) if is_grid else coords[0].size
        d_out = _lib.DeviceArray(npts, dtype, scratch=True)
        sampler(d_c, d_out, tuple(len(c) for c in coords) if is_grid else None,
                fill=0.0 if fill is None else fill)
        _lib.sync()
        res = d_out.download()
    else:
        raise ValueError(f"Method '{method}' is not defined")

    if log and not (method == 'volume' and isinstance(values, DeviceView)):
        res = 10 ** res
    return np.asarray(res).reshape(out_shape, order='F')

"""Field container and source-field assembly as consumed by the multigrid path.

Layout contract of the reference (emg3d/fields.py:40-301): one contiguous 1-D
array ``field = [fx | fy | fz]``; ``fx`` has shape ``(nx, ny+1, nz+1)``, ``fy``
``(nx+1, ny, nz+1)``, ``fz`` ``(nx+1, ny+1, nz)``, all Fortran-ordered (x
fastest); ``complex128`` for ``frequency > 0``, ``float64`` for
``frequency < 0`` (Laplace domain, ``s = -frequency``).  The same bytes are what
the CUDA library works on, so host<->device transfers are plain copies.
"""
import numpy as np
from scipy.constants import mu_0

__all__ = ['Field', 'SourceField', 'get_source_field', 'get_receiver', 'get_magnetic_field']


class Field:
    """x-, y-, z-directed edge fields in one array (emg3d/fields.py:40-136)."""

    def __init__(self, grid, data=None, frequency=None, dtype=None, electric=True):
        if frequency is not None:
            if frequency > 0:
                dtype = np.complex128
            elif frequency < 0:
                dtype = np.float64
            else:
                raise ValueError(
                    "`frequency` must be f>0 (frequency domain) or f<0 "
                    f"(Laplace domain). Provided: {frequency} Hz.")
        elif data is not None:
            dtype = np.asarray(data).dtype
        elif dtype is None:
            dtype = np.complex128
        self.grid = grid
        self._frequency = frequency
        self.electric = bool(electric)     # False: magnetic field, living on the faces
        n = grid.n_edges if self.electric else grid.n_faces
        if data is None:
            self._field = np.zeros(n, dtype=dtype)
        else:
            self._field = np.ascontiguousarray(np.asarray(data, dtype=dtype).ravel('F'))
            if self._field.size != n:
                raise ValueError(f"data must have {n} entries, got {self._field.size}")

    def __repr__(self):
        g = self.grid
        return (f"Field: {['magnetic', 'electric'][self.electric]}; "
                f"{g.shape_cells[0]} x {g.shape_cells[1]} x "
                f"{g.shape_cells[2]}; {self.field.size:,}")

    def __eq__(self, other):
        return (type(self).__name__ == type(other).__name__ and
                self.grid == other.grid and
                self._frequency == other._frequency and
                self.electric == getattr(other, 'electric', True) and
                np.allclose(self._field, other._field, atol=0, rtol=1e-10))

    def copy(self):
        return Field(self.grid, self._field.copy(), self._frequency, electric=self.electric)

    @property
    def field(self):
        return self._field

    @field.setter
    def field(self, value):
        self._field[:] = value

    def _view(self, comp):
        g = self.grid
        if self.electric:
            n = (g.n_edges_x, g.n_edges_y, g.n_edges_z)
            shp = (g.shape_edges_x, g.shape_edges_y, g.shape_edges_z)[comp]
        else:
            n = (g.n_faces_x, g.n_faces_y, g.n_faces_z)
            shp = (g.shape_faces_x, g.shape_faces_y, g.shape_faces_z)[comp]
        i0 = sum(n[:comp])
        return self._field[i0:i0 + n[comp]].reshape(shp, order='F')

    @property
    def fx(self):
        return self._view(0)

    @fx.setter
    def fx(self, v):
        self._view(0)[...] = v

    @property
    def fy(self):
        return self._view(1)

    @fy.setter
    def fy(self, v):
        self._view(1)[...] = v

    @property
    def fz(self):
        return self._view(2)

    @fz.setter
    def fz(self, v):
        self._view(2)[...] = v

    @property
    def dtype(self):
        return self._field.dtype

    @property
    def frequency(self):
        return None if self._frequency is None else abs(self._frequency)

    def interpolate_to_grid(self, grid, **interpolate_opts):
        """The field on another grid (emg3d/fields.py:303-346): every component is interpolated
        with :func:`emg3d_b200.maps.interpolate`; defaults ``method='cubic'``, ``log=False``,
        ``extrapolate=False``.  Returns ``self`` if the grids are identical."""
        from emg3d_b200 import maps
        if grid == self.grid:
            return self
        opts = {'method': 'cubic', 'extrapolate': False, 'log': False, **interpolate_opts,
                'grid': self.grid, 'xi': grid}
        data = np.concatenate([maps.interpolate(values=f, **opts).ravel('F')
                               for f in (self.fx, self.fy, self.fz)])
        return Field(grid, data, frequency=self._frequency, electric=self.electric)

    def get_receiver(self, receiver, method='cubic'):
        """Responses at receiver coordinates: :func:`get_receiver`."""
        return get_receiver(self, receiver, method)

    @property
    def sval(self):
        """Laplace parameter: 2 i pi f (f > 0) or -f (f < 0)."""
        if self._frequency is None:
            return None
        if self._frequency < 0:
            return np.array(-self._frequency)
        return np.array(2j * np.pi * self._frequency)

    @property
    def smu0(self):
        s = self.sval
        return None if s is None else s * mu_0


def _rotation(azimuth, elevation):
    a, e = np.deg2rad(azimuth), np.deg2rad(elevation)
    return np.array([np.cos(a) * np.cos(e), np.sin(a) * np.cos(e), np.sin(e)])


def _segment_vector(grid, p0, p1):
    """Distribute a straight current segment p0 -> p1 onto the edges.

    Per cell crossed: clip the segment to the cell, take the midpoint of the
    clipped piece and spread (piece length / total length) bilinearly onto the
    four parallel edges of each direction; finally scale each component by the
    segment's extent in that direction (emg3d/fields.py:792-938).  Returns the
    non-zero entries ``{flat index in [fx | fy | fz]: value}`` (a segment touches a
    handful of edges; no dense array is built).
    """
    nodes = [np.round(grid.nodes_x, 9), np.round(grid.nodes_y, 9),
             np.round(grid.nodes_z, 9)]
    p0, p1 = np.round(np.asarray(p0, float), 9), np.round(np.asarray(p1, float), 9)
    for a in range(3):
        if min(p0[a], p1[a]) < nodes[a][0] or max(p0[a], p1[a]) > nodes[a][-1]:
            raise ValueError(f"Provided source outside grid: {np.r_[[p0, p1]]}.")
    d = p1 - p0
    length = np.linalg.norm(d)
    if length < 1e-15:
        raise ValueError(f"Provided finite dipole has no length: {np.r_[[p0, p1]]}.")
    shapes = (grid.shape_edges_x, grid.shape_edges_y, grid.shape_edges_z)
    offs = (0, grid.n_edges_x, grid.n_edges_x + grid.n_edges_y)
    out = {}
    # cell index ranges touched by the segment
    rng = []
    for a in range(3):
        lo, hi = min(p0[a], p1[a]), max(p0[a], p1[a])
        i0 = max(0, int(np.searchsorted(nodes[a], lo, side='right')) - 1)
        i1 = max(0, int(np.searchsorted(nodes[a], hi, side='right')) - 1)
        rng.append(range(i0, min(i1 + 1, nodes[a].size - 1)))
    for iz in rng[2]:
        for iy in rng[1]:
            for ix in rng[0]:
                idx = (ix, iy, iz)
                t0, t1 = 0.0, 1.0          # parameter interval inside the cell
                for a in range(3):
                    if d[a] != 0:
                        ta = (nodes[a][idx[a]] - p0[a]) / d[a]
                        tb = (nodes[a][idx[a] + 1] - p0[a]) / d[a]
                        t0, t1 = max(t0, min(ta, tb)), min(t1, max(ta, tb))
                if t1 - t0 <= 0:
                    continue
                mid = p0 + 0.5 * (t0 + t1) * d
                frac = np.linalg.norm((t1 - t0) * d) / length
                r = [(mid[a] - nodes[a][idx[a]]) / grid.h[a][idx[a]] for a in range(3)]
                if min(min(r), min(1 - q for q in r)) < 0:
                    continue
                for c in range(3):
                    u, v = (c + 1) % 3, (c + 2) % 3
                    for du in (0, 1):
                        for dv in (0, 1):
                            j = list(idx)
                            j[u] += du
                            j[v] += dv
                            wu = r[u] if du else 1 - r[u]
                            wv = r[v] if dv else 1 - r[v]
                            flat = offs[c] + j[0] + shapes[c][0] * (j[1] + shapes[c][1] * j[2])
                            out[flat] = out.get(flat, 0.0) + wu * wv * frac
    for flat in out:
        c = 0 if flat < offs[1] else 1 if flat < offs[2] else 2
        out[flat] *= d[c]
    return out


class SourceField(Field):
    """Source field of dipoles / wires kept as what it is: a constant background and a handful
    of non-zero edges (``sparse`` = (indices, values)).  ``solve`` sends only those to the device
    (``emg3d_b200_fill_scatter``); the dense array of the Field contract is materialised on first
    access of ``field`` / ``fx`` ... -- after which the dense array is authoritative (a caller may
    have edited it) and is uploaded like any other field."""

    def __init__(self, grid, indices, values, background, frequency):
        self.grid, self._frequency, self.electric = grid, frequency, True
        self._dense = None
        self._sparse = (np.asarray(indices, dtype=np.int64), np.asarray(values))
        self._background = np.asarray(values).dtype.type(background)

    @property
    def sparse(self):
        """(indices, values, background), or None once the dense array was handed out."""
        return None if self._sparse is None else (*self._sparse, self._background)

    @property
    def dtype(self):
        return self._background.dtype if self._dense is None else self._dense.dtype

    @property
    def _field(self):
        if self._dense is None:
            idx, val = self._sparse
            self._dense = np.full(self.grid.n_edges, self._background, dtype=val.dtype)
            self._dense[idx] = val
            self._sparse = None
        return self._dense

    def copy(self):
        if self._sparse is not None:
            return SourceField(self.grid, self._sparse[0].copy(), self._sparse[1].copy(),
                               self._background, self._frequency)
        return Field(self.grid, self._dense.copy(), self._frequency)


def _square_loop(center, azimuth, elevation, area):
    """Closed square loop (5 points) of ``area`` m^2 perpendicular to a dipole: what stands in for
    a magnetic dipole (emg3d/electrodes.py:796-822; the loop's area takes the dipole's length)."""
    half_diag = np.sqrt(area / 2)
    hor = _rotation(azimuth + 90.0, 0.0) * half_diag
    ver = _rotation(azimuth, elevation + 90.0) * half_diag
    return np.asarray(center, dtype=float) + np.stack([hor, ver, -hor, -ver, hor])


def _dipole_angles(points):
    """(azimuth, elevation, length) of an electrode pair (emg3d/electrodes.py:758-793)."""
    dx, dy, dz = points[1] - points[0]
    return (np.degrees(np.arctan2(dy, dx)), np.degrees(np.arctan2(dz, np.hypot(dx, dy))),
            float(np.linalg.norm([dx, dy, dz])))


def _point_vector(grid, coordinates):
    """Point dipole ``(x, y, z, azimuth, elevation)`` by the adjoint of trilinear interpolation
    (emg3d/fields.py:662-745): per component the eight edges around the point get the trilinear
    weights (in the last cell of an axis: weight one), times the direction cosine.  Returns
    ``{flat index: value}``."""
    coo = np.asarray(coordinates, dtype=float)
    nodes = (grid.nodes_x, grid.nodes_y, grid.nodes_z)
    if any(coo[a] < nodes[a][0] or coo[a] > nodes[a][-1] for a in range(3)):
        raise ValueError(f"Provided source outside grid: {coordinates}.")
    centers = (grid.cell_centers_x, grid.cell_centers_y, grid.cell_centers_z)
    shapes = (grid.shape_edges_x, grid.shape_edges_y, grid.shape_edges_z)
    offs = (0, grid.n_edges_x, grid.n_edges_x + grid.n_edges_y)
    srcdir = _rotation(coo[3], coo[4])
    out = {}
    for c in range(3):
        axes = []                                            # per axis: [(index, weight), (index1, weight1)]
        for a in range(3):
            cc = centers[a] if a == c else nodes[a]
            n = shapes[c][a]
            i0 = max(0, int(np.searchsorted(cc, coo[a], side='right')) - 1)
            if i0 == n - 1:
                axes.append(((i0, 1.0), (i0, 1.0)))
            else:
                r = (coo[a] - cc[i0]) / (cc[i0 + 1] - cc[i0])
                axes.append(((i0, 1.0 - r), (i0 + 1, r)))
        for iz, wz in axes[2]:
            for iy, wy in axes[1]:
                for ix, wx in axes[0]:
                    flat = offs[c] + ix + shapes[c][0] * (iy + shapes[c][1] * iz)
                    out[flat] = wx * wy * wz * srcdir[c]     # (assigned, not added: as the reference)
    return out


def get_source_field(grid, source, frequency, strength=1.0, length=1.0, electric=True):
    """Source term ``-s mu_0 J_s`` on the edges (emg3d/fields.py:386-519).

    ``source``: ``(x, y, z, azimuth, elevation)`` (a dipole of ``length`` metres centred there;
    emg3d/electrodes.py:752-755), ``(x0, x1, y0, y1, z0, z1)`` or ``[[x0, y0, z0], [x1, y1, z1]]``
    (a finite dipole), an array of shape ``(n, 3)``, n > 2 (a wire) -- or a source object of the
    reference (``emg3d.TxElectricDipole``, ``TxMagneticDipole``, ``TxElectricWire``,
    ``TxElectricPoint``: anything with ``points`` / ``coordinates`` and ``strength``).
    ``electric=False``: a magnetic dipole, represented like in the reference by a square loop of
    electric wire perpendicular to it whose area is the dipole's length.  Dipoles and wires are
    distributed onto the edges by their length fraction per cell, point sources by the adjoint of
    trilinear interpolation.  ``frequency=None`` returns the real, frequency-independent vector.
    Magnetic POINT sources need ``discretize`` in the reference and are outside this path.
    Returns a :class:`SourceField`: same values as the reference's dense field, held sparsely.
    """
    point = None
    if hasattr(source, 'points') and hasattr(source, 'strength'):      # a Tx* object
        kind = type(source).__name__
        strength = source.strength
        if kind == 'TxMagneticPoint':
            raise NotImplementedError("magnetic point sources (they need `discretize` in the "
                                      "reference) are outside this path")
        if kind == 'TxElectricPoint':
            point = np.asarray(source.coordinates, dtype=float)
        pts = np.asarray(source.points, dtype=float).reshape(-1, 3)
    else:
        src = np.asarray(source, dtype=float).squeeze()
        if src.size == 5:
            if electric:
                half = _rotation(src[3], src[4]) * length / 2
                pts = np.array([src[:3] - half, src[:3] + half])
            else:
                pts = _square_loop(src[:3], src[3], src[4], length)
        elif src.size == 6:
            pts = src.reshape((2, 3), order='F') if src.ndim == 1 else src.reshape(2, 3)
            if np.allclose(pts[0], pts[1]):
                raise ValueError(
                    "The two electrodes are identical, use the format "
                    "(x, y, z, azimuth, elevation) instead. "
                    f"Provided coordinates: {src}.")
            if not electric:
                azimuth, elevation, dlen = _dipole_angles(pts)
                pts = _square_loop(pts.sum(0) / 2, azimuth, elevation, dlen)
        elif src.size > 6 and src.size % 3 == 0:
            pts = src.reshape(-1, 3)
        else:
            raise ValueError(
                "Coordinates are wrong defined. They must be defined either "
                "as a point, (x, y, z, azimuth, elevation), or as two points, "
                "(x1, x2, y1, y2, z1, z2) or [[x1, y1, z1], [x2, y2, z2]]. "
                f"Provided coordinates: {src}.")
    if point is not None:
        total = _point_vector(grid, point)
        pts = pts[:0]
    else:
        total = {}
    for a, b in zip(pts[:-1], pts[1:]):
        for flat, v in _segment_vector(grid, a, b).items():
            total[flat] = total.get(flat, 0.0) + v
    idx = np.array(sorted(total), dtype=np.int64)
    # the reference's sequence of operations on the dense array, applied to the non-zeros and to
    # one background zero (same bits as its dense result)
    if frequency is not None and frequency == 0:
        raise ValueError(
            "`frequency` must be f>0 (frequency domain) or f<0 "
            f"(Laplace domain). Provided: {frequency} Hz.")
    dtype = np.complex128 if frequency is not None and frequency > 0 else np.float64
    vals = np.array([total[i] for i in idx] + [0.0]).astype(dtype)
    vals *= strength
    if frequency is not None:
        sval = np.array(2j * np.pi * frequency) if frequency > 0 else np.array(-frequency)
        vals *= -(sval * mu_0)
    return SourceField(grid, idx, vals[:-1], vals[-1], frequency)


def receiver_coordinates(receiver):
    """``(x, y, z, azimuth, elevation)`` of a receiver argument (fields.py:560-583): an object with
    ``coordinates``, a list of such, or the tuple itself."""
    if hasattr(receiver, 'coordinates'):
        return receiver.coordinates
    if hasattr(tuple(receiver)[0], 'coordinates'):
        return tuple(np.array([r.coordinates for r in receiver], dtype=float).T)
    if len(receiver) != 5:
        raise ValueError(
            "`receiver` needs to be in the form "
            "(x, y, z, azimuth, elevation). "
            f"Length of provided `receiver`: {len(receiver)}.")
    return receiver


def get_receiver(field, receiver, method='cubic'):
    """Field (response) at receiver coordinates (emg3d/fields.py:522-615).

    ``receiver``: ``(x, y, z, azimuth, elevation)`` (scalars or arrays), an object with
    ``coordinates``, or a list of such.  The three components are interpolated (``'cubic'`` or
    ``'linear'``) and combined with the direction cosines; receivers outside the grid or inside
    the outermost cells (PEC boundary) are NaN.  ``field``: a :class:`Field` (uploaded) or a
    :class:`DeviceField` (the device-resident result of ``solve(..., return_field='device')``);
    the interpolation runs on the GPU (csrc/interp.cu), only the responses come back.
    """
    from emg3d_b200 import _lib, maps
    coordinates = receiver_coordinates(receiver)
    if method not in ('cubic', 'linear'):
        raise ValueError(f"get_receiver: method must be 'cubic' or 'linear'; provided: {method!r}.")
    grid = field.grid
    xyz = np.broadcast_arrays(*[np.atleast_1d(np.asarray(c, dtype=float)) for c in coordinates[:3]])
    shape = xyz[0].shape
    xyz = [c.ravel() for c in xyz]
    factors = [np.broadcast_to(f, shape).ravel() for f in _rotation(np.asarray(coordinates[3], dtype=float),
                                                                    np.asarray(coordinates[4], dtype=float))]
    dtype = np.dtype(field.dtype)
    if isinstance(field, DeviceField):
        comps = maps.field_components(field.array, grid, dtype)
    else:
        d_f = _lib.DeviceArray.from_host(np.asarray(field.field))
        comps = maps.field_components(d_f, grid, dtype)
    d_out = _lib.DeviceArray(xyz[0].size, dtype, scratch=True)
    d_out.zero()
    # per component: out += factor * interpolated value, on the device; one factor per receiver
    # is applied on the host afterwards when the factors differ between receivers
    uniform = all(np.ptp(f) == 0 for f in factors)
    parts = []
    for k, (comp, f) in enumerate(zip(comps, factors)):
        if not np.any(np.abs(f) > 1e-10):
            continue
        points, _, _, _ = maps._points_from_grids(grid, comp.shape, tuple(xyz), method)
        if uniform:
            maps.sample_points(comp, points, xyz, method, mode='constant', fill=np.nan, d_out=d_out,
                               scale=float(f[0]), accumulate=True)
        else:
            parts.append(f * maps.sample_points(comp, points, xyz, method, mode='constant',
                                                fill=np.nan).download())
    resp = d_out.download() if uniform else (np.sum(parts, axis=0) if parts else np.zeros(xyz[0].size, dtype))
    # PEC: receivers in the outermost cells are not to be trusted
    pec = np.zeros(xyz[0].size, dtype=bool)
    for c, name in zip(xyz, 'xyz'):
        nodes = getattr(grid, 'nodes_' + name)
        pec |= (c < nodes[1]) | (c > nodes[-2])
    resp = np.array(resp)
    resp[pec] = np.nan
    return resp.reshape(shape, order='F')


class DeviceField:
    """Result field left on the device (``solve(..., return_field='device')``): ``array`` is the
    device buffer in the Field layout; ``download()`` gives the host :class:`Field`."""

    def __init__(self, grid, array, dtype, frequency):
        self.grid, self.array, self.dtype, self._frequency = grid, array, np.dtype(dtype), frequency

    def get_receiver(self, receiver, method='cubic'):
        return get_receiver(self, receiver, method)

    def download(self):
        f = Field(self.grid, dtype=self.dtype, frequency=self._frequency)
        self.array.download(out=f.field)
        return f


def get_magnetic_field(model, efield):
    r"""Magnetic field on the faces from the electric field on the edges, by Faraday's
    law :math:`\nabla \times \mathbf{E} = \rm{i}\omega\mu\mathbf{H}`
    (emg3d.fields.get_magnetic_field, fields.py:617-659).

    ``zeta = V / mu_r`` is built on the device from the model (models.VolumeModel),
    the curl runs in one streaming kernel (csrc/hfield.cu); only the electric field
    goes up and the magnetic field comes down.
    """
    from emg3d_b200 import _lib, solver
    lv = solver._Level.from_model(model, efield)
    d_e = _lib.DeviceArray.from_host(np.asarray(efield.field))
    hfield = Field(efield.grid, frequency=efield._frequency, electric=False)
    d_h = _lib.DeviceArray(hfield.field.size, hfield.field.dtype)
    scale = 1.0 / complex(efield.smu0)
    _lib.check(_lib.load().emg3d_b200_magnetic_field(lv.handle.ptr, d_e.ptr, d_h.ptr,
                                                     scale.real, scale.imag))
    d_h.download(out=hfield.field)
    return hfield

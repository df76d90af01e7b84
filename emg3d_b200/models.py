"""Model containers as consumed by the multigrid path.

``VolumeModel`` is the coefficient input of the hot path (emg3d/models.py:627-717):
``eta_a = -s mu_0 V (sigma_a + s eps_0 eps_r)`` and ``zeta = V / mu_r`` per cell,
Fortran-ordered ``(nx, ny, nz)``; isotropic models alias ``eta_y = eta_z = eta_x``,
VTI aliases ``eta_y = eta_x``, HTI ``eta_z = eta_x``.  ``Model`` is the small part
of emg3d.models.Model (36-624) needed to build one.
"""
import numpy as np
from scipy.constants import epsilon_0

from emg3d_b200 import meshes

__all__ = ['Model', 'VolumeModel']


class _Map:
    """property <-> conductivity (emg3d/maps.py:52-227, forward/backward only)."""
    _fwd = {
        'Conductivity': (lambda c: c, lambda p: p),
        'Resistivity': (lambda c: 1.0 / c, lambda p: 1.0 / p),
        'LgConductivity': (np.log10, lambda p: 10.0 ** p),
        'LnConductivity': (np.log, np.exp),
        'LgResistivity': (lambda c: -np.log10(c), lambda p: 10.0 ** -p),
        'LnResistivity': (lambda c: -np.log(c), lambda p: np.exp(-p)),
    }

    def __init__(self, name):
        if name not in self._fwd:
            raise ValueError(f"Unknown mapping: {name!r}")
        self.name = name

    # looked up by name, so that a Model can be pickled (batch.solve_many)
    def forward(self, conductivity):
        return self._fwd[self.name][0](conductivity)

    def backward(self, mapped):
        return self._fwd[self.name][1](mapped)

    def derivative_chain(self, gradient, mapped):
        """Chain rule from conductivity to the mapping's space, in place (maps.py:100-228)."""
        if self.name == 'Conductivity':
            return
        sigma = self.backward(mapped)
        gradient *= {'LgConductivity': sigma * np.log(10), 'LnConductivity': sigma,
                     'Resistivity': -sigma ** 2, 'LgResistivity': -sigma * np.log(10),
                     'LnResistivity': -sigma}[self.name]


class Model:
    """Resistivity/conductivity model with triaxial anisotropy, mu_r, eps_r."""

    _properties = ['property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r']

    def __init__(self, grid, property_x=1., property_y=None, property_z=None,
                 mu_r=None, epsilon_r=None, mapping='Resistivity'):
        self.grid = grid
        self.shape = tuple(grid.shape_cells)
        self.map = mapping if hasattr(mapping, 'backward') else _Map(mapping)
        for name, val in zip(self._properties,
                             (property_x, property_y, property_z, mu_r, epsilon_r)):
            setattr(self, name, self._init(name, val))
        self.case = {(False, False): 'isotropic', (True, False): 'HTI',
                     (False, True): 'VTI', (True, True): 'triaxial'}[
                         (self.property_y is not None, self.property_z is not None)]

    def _init(self, name, val):
        if val is None:
            return None
        arr = np.asarray(val, dtype=np.float64)
        if arr.ndim == 0:
            arr = arr * np.ones(self.shape)
        elif arr.ndim == 1 and arr.size == int(np.prod(self.shape)):
            arr = arr.reshape(self.shape, order='F')
        if arr.shape != self.shape:
            raise ValueError(
                f"`{name}` must be {self.shape} or (), provided: {arr.shape}.")
        if not np.all(np.isfinite(arr)):
            raise ValueError(f"`{name}` must be finite.")
        return np.asfortranarray(arr)

    def __repr__(self):
        return (f"Model: {self.map.name}; {self.case}; "
                f"{self.shape[0]} x {self.shape[1]} x {self.shape[2]}")

    def interpolate_to_grid(self, grid, **interpolate_opts):
        """The model on another grid (emg3d/models.py:322-380): every property is interpolated with
        :func:`emg3d_b200.maps.interpolate`; defaults ``method='volume'``, ``extrapolate=True`` and
        ``log=True`` unless the mapping is logarithmic already.  Returns ``self`` if the grids
        are identical."""
        from emg3d_b200 import maps
        if grid == self.grid:
            return self
        opts = {'method': 'volume', 'extrapolate': True, 'log': not self.map.name.startswith('L'),
                **interpolate_opts, 'grid': self.grid, 'xi': grid}
        props = {name: maps.interpolate(values=getattr(self, name), **opts)
                 for name in self._properties if getattr(self, name) is not None}
        return Model(grid, mapping=self.map.name, **props)


class VolumeModel:
    """Volume-averaged eta_{x,y,z} and zeta for one Laplace parameter."""

    def __init__(self, model, sfield):
        self.case = model.case
        self.grid = meshes.BaseMesh(model.grid.h, model.grid.origin)
        shape = tuple(self.grid.shape_cells)
        vol = self.grid.cell_volumes.reshape(shape, order='F')
        sval, smu0 = sfield.sval, sfield.smu0
        eps = None if model.epsilon_r is None else sval * epsilon_0 * model.epsilon_r
        for ax, name in zip('xyz', ('property_x', 'property_y', 'property_z')):
            prop = getattr(model, name)
            if prop is None:
                eta = None
            else:
                cond = model.map.backward(prop)
                eta = -smu0 * vol * (cond if eps is None else cond + eps)
                eta = np.asfortranarray(eta)
            setattr(self, '_eta_' + ax, eta)
        zeta = vol.copy()
        if model.mu_r is not None:
            zeta /= model.mu_r
        self._zeta = np.asfortranarray(zeta)

    @property
    def eta_x(self):
        return self._eta_x

    @property
    def eta_y(self):
        return self._eta_y if self.case in ('HTI', 'triaxial') else self._eta_x

    @property
    def eta_z(self):
        return self._eta_z if self.case in ('VTI', 'triaxial') else self._eta_x

    @property
    def zeta(self):
        return self._zeta

"""Multigrid solver driver: the public API of ``emg3d.solver`` on a B200.

Same functions, arguments, return values, messages and exit codes as the
reference (``emg3d/solver.py``): :func:`solve`, :func:`solve_source`,
:func:`multigrid`, :func:`krylov`, :func:`smoothing`, :func:`restriction`,
:func:`prolongation`, :func:`residual`, :class:`MGParameters`,
:class:`RegularGridProlongator`.  What differs is where the work happens:

* every field, coefficient array and scratch vector lives in HBM for the whole
  solve; only scalars (norms, dot products) and the final field cross PCIe;
* the grid hierarchy (coarse grids, summed coefficients, restriction weights,
  interpolation tables, line factorisations) is built once per solve and cached,
  where the reference rebuilds it at every visit (solver.py:888-931, 975-1007);
* BiCGSTAB, CGS and GCROT(m,k) run on the device(s) with the recurrences of SciPy's
  implementations (only GCROT's small Hessenberg problem is solved on the host);
* ``order='color'`` (default) smooths in multicolour order, ``order='lex'``
  reproduces the reference's lexicographic Gauss-Seidel sweeps.

The public functions also accept the reference's host containers (anything with
the attributes of ``emg3d.Field`` / ``emg3d.models.VolumeModel``): data are then
uploaded, processed and written back in place.
"""
import itertools
import os
from dataclasses import dataclass
from datetime import datetime, timedelta
from time import perf_counter
from typing import Union

import numpy as np
import scipy.sparse.linalg as _ssl
from scipy.constants import epsilon_0

from emg3d_b200 import _lib, core, fields, meshes, models

__all__ = ['solve', 'solve_source', 'multigrid', 'krylov', 'smoothing',
           'restriction', 'prolongation', 'residual', 'MGParameters',
           'RegularGridProlongator', 'Workspace']

__version__ = '0.1.0'

# lr_dir -> line directions applied one after the other (solver.py:836-846)
_LR_DIRS = {0: (), 1: (1,), 2: (2,), 3: (3,), 4: (2, 3), 5: (1, 3), 6: (1, 2),
            7: (1, 2, 3)}
_LR_CODE = {v: k for k, v in _LR_DIRS.items()}


class Timer:
    """Wall-clock helper with the reference's string formats (utils.py:169-198)."""

    def __init__(self):
        self._t0 = perf_counter()

    @property
    def t0(self):
        return self._t0

    @property
    def now(self):
        return datetime.now().strftime("%H:%M:%S")

    @property
    def elapsed(self):
        return perf_counter() - self._t0

    @property
    def runtime(self):
        return str(timedelta(seconds=np.round(self.elapsed)))

    def __repr__(self):
        return f"Runtime : {self.runtime}"


# =========================================================================== #
# Device-side hierarchy
# =========================================================================== #

def _set_dtype(handle, dtype):
    """Tell a bare level (no coefficients) which dtype its fields have."""
    _lib.check(_lib.load().emg3d_b200_level_set_model(
        handle.ptr, int(np.dtype(dtype).kind == 'c'), None, None, None, None))


class _Level:
    """One grid of the hierarchy with its coefficients resident on the device."""

    def __init__(self, grid, dtype, case, eta, zeta):
        self.grid = grid
        self.dtype = np.dtype(dtype)
        self.cplx = self.dtype.kind == 'c'
        self.case = case
        self.eta = eta            # three DeviceArrays, aliased as in the reference
        self.zeta = zeta
        self.handle = _lib.LevelHandle(grid.h)
        self.handle.set_model(self.cplx, eta[0], eta[1], eta[2], zeta)
        self.children = {}        # c_sc_dir -> coarse _Level
        self.sc_to_parent = None
        self._res = None          # residual scratch
        self.s = None             # coarse source / field (owned by coarse levels)
        self.e = None

    @property
    def shape(self):
        return tuple(self.grid.shape_cells)

    @property
    def n_edges(self):
        return int(self.grid.n_edges)

    def new_field(self, zero=True):
        f = _lib.DeviceArray(self.n_edges, self.dtype)
        if zero:
            f.zero()
        return f

    def res_buffer(self):
        if self._res is None:
            self._res = _lib.DeviceArray(self.n_edges, self.dtype)
        return self._res

    @classmethod
    def from_volume_model(cls, vmodel, dtype):
        """Upload a (host) VolumeModel; keeps eta aliases (models.py:698-712)."""
        grid = vmodel.grid
        dtype = np.dtype(dtype)
        arrs, devs = [], []
        for a in (vmodel.eta_x, vmodel.eta_y, vmodel.eta_z):
            for k, b in enumerate(arrs):
                if a is b:
                    devs.append(devs[k])
                    break
            else:
                if np.dtype(a.dtype).kind == 'c' and dtype.kind != 'c':
                    raise ValueError("complex coefficients with a real-valued field")
                devs.append(_lib.DeviceArray.from_host(np.asarray(a, dtype=dtype)))
            arrs.append(a)
        zeta = _lib.DeviceArray.from_host(np.asarray(vmodel.zeta, dtype=np.float64))
        if not isinstance(grid, meshes.BaseMesh):
            grid = meshes.BaseMesh(grid.h, grid.origin)
        return cls(grid, dtype, getattr(vmodel, 'case', 'triaxial'), devs, zeta)

    _MAPS = {'Conductivity': 0, 'Resistivity': 1, 'LgConductivity': 2, 'LnConductivity': 3,
             'LgResistivity': 4, 'LnResistivity': 5}

    @classmethod
    def from_model(cls, model, sfield):
        """Build eta/zeta on the device from a Model (models.VolumeModel, 654-691).

        Only the real property arrays cross PCIe; falls back to the host
        ``VolumeModel`` for property maps the device kernel does not know.
        """
        map_name = getattr(model.map, 'name', None)
        dtype = _field_dtype(sfield)
        if map_name not in cls._MAPS:
            return cls.from_volume_model(models.VolumeModel(model, sfield), dtype)
        grid = meshes.BaseMesh(model.grid.h, model.grid.origin)
        handle = _lib.LevelHandle(grid.h)
        n = grid.n_cells

        def up(arr):
            return None if arr is None else _lib.DeviceArray.from_host(
                np.asarray(arr, dtype=np.float64))

        props = [up(model.property_x), up(model.property_y), up(model.property_z),
                 up(model.mu_r), up(model.epsilon_r)]
        eta_x = _lib.DeviceArray(n, dtype)
        eta_y = _lib.DeviceArray(n, dtype) if props[1] is not None else eta_x
        eta_z = _lib.DeviceArray(n, dtype) if props[2] is not None else eta_x
        zeta = _lib.DeviceArray(n, np.float64)
        c = complex(-sfield.smu0)
        se = complex(sfield.sval * epsilon_0)
        ptr = lambda a: None if a is None else a.ptr
        _lib.check(_lib.load().emg3d_b200_volume_model(
            handle.ptr, int(dtype.kind == 'c'), c.real, c.imag, se.real, se.imag,
            cls._MAPS[map_name], *[ptr(p) for p in props], eta_x.ptr, eta_y.ptr, eta_z.ptr,
            zeta.ptr))
        self = cls.__new__(cls)
        self.grid, self.dtype, self.cplx, self.case = grid, dtype, dtype.kind == 'c', model.case
        self.eta, self.zeta, self.handle = [eta_x, eta_y, eta_z], zeta, handle
        handle.set_model(self.cplx, eta_x, eta_y, eta_z, zeta)
        self.children, self._res, self.sc_to_parent = {}, None, None
        self.s = self.e = None
        return self

    def coarse(self, sc_dir):
        """Coarse level for the (current) semicoarsening pattern; cached."""
        sc_dir = int(sc_dir)
        if sc_dir in self.children:
            return self.children[sc_dir]
        g = self.grid
        cflag = core.SC_FLAGS[sc_dir]
        nodes = (g.nodes_x, g.nodes_y, g.nodes_z)
        centers = (g.cell_centers_x, g.cell_centers_y, g.cell_centers_z)
        ch = [np.diff(nodes[a][::2 if cflag[a] else 1]) for a in range(3)]
        cg = meshes.BaseMesh(ch, g.origin)
        cnodes = (cg.nodes_x, cg.nodes_y, cg.nodes_z)
        ccenters = (cg.cell_centers_x, cg.cell_centers_y, cg.cell_centers_z)
        weights, lo, frac = [None] * 9, [], []
        for a in range(3):
            if cflag[a]:
                weights[3 * a:3 * a + 3] = core.restrict_weights(
                    nodes[a], centers[a], g.h[a], cnodes[a], ccenters[a], cg.h[a])
            li, fr = core.interpolation_table(nodes[a], cnodes[a])
            lo.append(li)
            frac.append(fr)
        handle = _lib.LevelHandle(cg.h)
        handle.link(self.handle, cflag, weights, lo, frac)
        lib = _lib.load()

        def sum_cells(arr, cplx):
            out = _lib.DeviceArray(cg.n_cells, arr.dtype)
            _lib.check(lib.emg3d_b200_restrict_cells(handle.ptr, int(cplx), arr.ptr, out.ptr))
            return out

        ceta = []
        for k, a in enumerate(self.eta):
            for j in range(k):
                if a is self.eta[j]:
                    ceta.append(ceta[j])
                    break
            else:
                ceta.append(sum_cells(a, self.cplx))
        czeta = sum_cells(self.zeta, False)
        child = _Level.__new__(_Level)
        child.grid, child.dtype, child.cplx, child.case = cg, self.dtype, self.cplx, self.case
        child.eta, child.zeta, child.handle = ceta, czeta, handle
        handle.set_model(self.cplx, ceta[0], ceta[1], ceta[2], czeta)
        child.children, child._res = {}, None
        child.sc_to_parent = sc_dir
        child.s = child.new_field()
        child.e = child.new_field()
        self.children[sc_dir] = child
        return child


_DIGEST_POOL = None


def _digest_pool():
    global _DIGEST_POOL
    if _DIGEST_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _DIGEST_POOL = ThreadPoolExecutor(max_workers=8)
    return _DIGEST_POOL


class Workspace:
    """Device-resident state that can be reused by consecutive :func:`solve` calls.

    ``solve(model, sfield, workspace=ws)`` keeps the coefficient arrays, the
    coarse-grid hierarchy and the cached line factorisations of ``model`` on the
    GPU, keyed by the model object, the Laplace parameter of ``sfield`` and the
    dtype.  A checksum over every element of the property arrays and the cell widths
    guards against models that were modified in place (``model.property_x[a:b] = v``
    between two solves rebuilds the hierarchy).
    """

    def __init__(self, max_models=2, pinned_result=False):
        self.max_models = int(max_models)
        self._levels = {}
        # With ``pinned_result=True`` a solve without ``efield=`` returns a Field
        # whose data live in a page-locked buffer owned by this workspace (fast
        # device-to-host copy).  The buffer is reused: the returned field is only
        # valid until the next solve with this workspace -- copy it to keep it.
        self.pinned_result = bool(pinned_result)
        self._result = None
        self._buffers = {}

    @staticmethod
    def _digest(a):
        """Checksum over EVERY element of a property array (an in-place edit anywhere changes it):
        wrapping sum and xor of the 64-bit patterns, evaluated in parallel chunks (NumPy
        reductions release the GIL).  ~10 ms for a 256^3 array."""
        a = np.asarray(a)
        if not a.flags.writeable:
            # a read-only array cannot be edited in place: its identity is its fingerprint
            # (freeze a model with `arr.flags.writeable = False` to skip the ~10 ms checksum
            # per 256^3 array and solve)
            return a.shape, a.__array_interface__['data'][0], -1
        flat = a.reshape(-1, order='A')
        if flat.dtype.itemsize % 8 or not (flat.flags.c_contiguous or flat.flags.f_contiguous):
            flat = np.ascontiguousarray(flat, dtype=np.float64)
        bits = flat.view(np.uint64)
        nchunk = max(1, min(8, bits.size // (1 << 20)))

        def part(k):
            c = bits[k * bits.size // nchunk:(k + 1) * bits.size // nchunk]
            return int(c.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(c))

        if nchunk == 1:
            parts = [part(0)]
        else:
            parts = list(_digest_pool().map(part, range(nchunk)))
        total, mixed = 0, 0
        for k, (sm, xr) in enumerate(parts):
            total = (total + sm * (2 * k + 1)) & 0xFFFFFFFFFFFFFFFF
            mixed ^= xr
        return a.shape, total, mixed

    @classmethod
    def _fingerprint(cls, model):
        fp = [tuple(np.asarray(h, dtype=np.float64).tobytes() for h in model.grid.h)]
        for name in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r'):
            a = getattr(model, name, None)
            fp.append(None if a is None else cls._digest(a))
        fp.append(getattr(model.map, 'name', None))
        return tuple(fp)

    def level(self, model, sfield):
        key = (id(model), complex(sfield.sval), _field_dtype(sfield).str)
        fp = self._fingerprint(model)
        hit = self._levels.get(key)
        if hit is not None and hit[0] == fp:
            return hit[1]
        while len(self._levels) >= self.max_models:
            self._levels.pop(next(iter(self._levels)))
        lv = _Level.from_model(model, sfield)
        self._levels[key] = (fp, lv)
        return lv

    def device_buffer(self, name, size, dtype):
        """Cached device array (source / field of the finest grid) reused across solves."""
        buf = self._buffers.get(name)
        if buf is None or buf.size != size or buf.dtype != np.dtype(dtype):
            buf = self._buffers[name] = _lib.DeviceArray(size, dtype)
        return buf

    def result_buffer(self, size, dtype):
        if (self._result is None or self._result.size != size or
                self._result.dtype != np.dtype(dtype)):
            self._result = _lib.PinnedArray(size, dtype)
        return self._result.array

    def clear(self):
        self._levels.clear()
        self._buffers.clear()
        self._result = None


def _order(var):
    return core.order_id(getattr(var, 'order', None))


def _dev_residual(lv, s, e, norm=False, out=None):
    """r = s - A e on the device; returns the norm or the residual buffer."""
    lib = _lib.load()
    if norm:
        val = _lib.c_double(0.0)
        _lib.check(lib.emg3d_b200_residual_norm(lv.handle.ptr, s.ptr, e.ptr, None,
                                                _lib.byref(val)))
        return val.value
    r = lv.res_buffer() if out is None else out
    _lib.check(lib.emg3d_b200_residual(lv.handle.ptr, s.ptr, e.ptr, r.ptr, None))
    return r


def _dev_smoothing(lv, s, e, nu, lr_dir, order):
    lib = _lib.load()
    c_lr_dir = int(_current_lr_dir(lr_dir, lv.grid))
    dirs = _LR_DIRS[c_lr_dir] or (0,)
    for ldir in dirs:
        _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, e.ptr, s.ptr, int(nu),
                                               ldir, order))


def _dev_restriction(lv, res, sc_dir):
    """Coarse level with its source = R res and a zero field."""
    child = lv.coarse(sc_dir)
    _lib.check(_lib.load().emg3d_b200_restrict(child.handle.ptr, res.ptr, child.s.ptr))
    child.e.zero()
    return child


def _dev_prolongation(child, e_fine):
    _lib.check(_lib.load().emg3d_b200_prolong(child.handle.ptr, e_fine.ptr, child.e.ptr))


# =========================================================================== #
# Public API
# =========================================================================== #

def _field_dtype(f):
    """dtype of a field without touching its array (a SourceField stays sparse)."""
    dt = getattr(f, 'dtype', None)
    return np.dtype(dt) if dt is not None else np.dtype(np.asarray(f.field).dtype)


def solve(model, sfield, sslsolver=True, semicoarsening=True,
          linerelaxation=True, verb=0, **kwargs):
    r"""Solve the 3-D EM diffusion problem with multigrid on the GPU.

    Drop-in for ``emg3d.solver.solve`` (emg3d/solver.py:52-449): same
    parameters (``cycle, efield, tol, maxit, nu_init, nu_pre, nu_coarse,
    nu_post, clevel, return_info, log, plain, always_return``), same return
    convention, same ``info_dict`` keys, messages and exit codes.

    Additional keyword
    ------------------
    order : {'color', 'lex'}
        Gauss-Seidel ordering inside the smoothers; default from
        ``emg3d_b200.core.ORDER`` (environment ``EMG3D_B200_ORDER``, 'color').
    workspace : Workspace, optional
        Keeps the device-resident coefficients, grid hierarchy and line
        factorisations of ``model`` alive between calls (many sources or
        restarts on one model and frequency).
    receivers : tuple or receiver object(s), optional
        ``(x, y, z, azimuth, elevation)``: the responses at these receivers are sampled on the
        device right after the solve (:func:`emg3d_b200.fields.get_receiver`,
        ``receiver_method`` 'cubic' or 'linear') and returned after the field.
    return_field : {True, False, 'device'}, default True
        False: the field is not downloaded (with ``receivers``: only the responses cross PCIe);
        'device': a :class:`emg3d_b200.fields.DeviceField` is returned (valid until the next
        solve with the same workspace).  A provided ``efield`` is only updated when True.
    comm : parallel.NcclComm, optional
        ONE solve on the GPUs of ``comm`` (z-slabs, one process per GPU): a collective call,
        every rank passes the same arguments and receives the whole field.
    n_gpus : int, optional
        Same from a single process: the ranks are spawned (devices 0 .. n_gpus - 1).
    """
    # one solve on several GPUs (z-slab decomposition, emg3d_b200.parallel): `comm=` inside a
    # multi-process launch (collective call), or `n_gpus=` to spawn the ranks from here
    comm, n_gpus = kwargs.pop('comm', None), kwargs.pop('n_gpus', None)
    if comm is not None or (n_gpus is not None and int(n_gpus) > 1):
        from emg3d_b200 import parallel
        if comm is not None:
            return parallel.solve_distributed(model, sfield, comm, sslsolver=sslsolver,
                                              semicoarsening=semicoarsening,
                                              linerelaxation=linerelaxation, verb=verb, **kwargs)
        return parallel.solve_spawn(model, sfield, int(n_gpus), sslsolver=sslsolver,
                                    semicoarsening=semicoarsening, linerelaxation=linerelaxation,
                                    verb=verb, **kwargs)
    always_return = kwargs.pop('always_return', False)
    if kwargs.pop('plain', False):
        sslsolver = False if sslsolver is True else sslsolver
        semicoarsening = False if semicoarsening is True else semicoarsening
        linerelaxation = False if linerelaxation is True else linerelaxation
    efield = kwargs.pop('efield', None)
    order = kwargs.pop('order', None)
    workspace = kwargs.pop('workspace', None)
    receivers = kwargs.pop('receivers', None)
    receiver_method = kwargs.pop('receiver_method', 'cubic')
    return_field = kwargs.pop('return_field', True)
    if return_field not in (True, False, 'device'):
        raise ValueError(f"return_field must be True, False or 'device'; provided: {return_field!r}.")
    core.order_id(order)  # validate early

    var = MGParameters(
        sslsolver=sslsolver, semicoarsening=semicoarsening,
        linerelaxation=linerelaxation, shape_cells=model.shape, verb=verb,
        **kwargs)
    var.order = order

    var.cprint(f"\n:: emg3d START :: {var.time.now} :: "
               f"v{__version__}\n", 2)
    var.cprint(var, 2)

    if sfield.frequency is None:
        # (the reference evaluates the source norm first; a Field without
        # frequency fails either way)
        raise ValueError(
            "Source field is missing frequency information; Create "
            "it with `emg3d.fields.get_source_field`, or initiate it "
            "with `emg3d.fields.Field`, providing frequency information."
        )

    # Coefficients to the device: eta/zeta are computed there from the property
    # arrays (models.VolumeModel, emg3d/models.py:654-691); a workspace keeps
    # them (and the whole level hierarchy) alive across calls.
    dtype = _field_dtype(sfield)
    if workspace is not None:
        level = workspace.level(model, sfield)
    else:
        level = _Level.from_model(model, sfield)
    if workspace is not None:
        d_s = workspace.device_buffer('s', level.n_edges, dtype)
    else:
        d_s = _lib.DeviceArray(level.n_edges, dtype)
    # dipole / wire sources are zero on all but a few edges: only those cross PCIe
    sparse = getattr(sfield, 'sparse', None)
    if sparse is not None:                               # fields.SourceField: nothing dense exists
        d_s.fill_scatter(sparse[2], sparse[0], sparse[1])
        var.sparse_source = True
    else:
        var.sparse_source = d_s.upload_sparse(np.asarray(sfield.field))
    info = ""

    # Reference error for the tolerance: ||b||_2 (solver.py:312), on the device.
    var.l2_refe = _Vec(level.cplx, level.n_edges).norm(d_s)
    var.error_at_cycle[0] = var.l2_refe

    if efield is None:
        if workspace is not None and workspace.pinned_result:
            # result lands in a page-locked buffer owned by the workspace
            efield = fields.Field(model.grid, workspace.result_buffer(level.n_edges, dtype),
                                  frequency=sfield._frequency)
        else:
            efield = fields.Field(model.grid, dtype=dtype, frequency=sfield._frequency)
        if workspace is not None:
            d_e = workspace.device_buffer('e', level.n_edges, dtype)
            d_e.zero()
        else:
            d_e = level.new_field()
        var.do_return = True
        # zero start: the initial residual is the source itself, ||r|| = ||b||
        var.e_is_zero, var.s_norm = not var.sslsolver, var.l2_refe
    else:
        if dtype != efield.field.dtype:
            raise ValueError(
                "Source field and electric field must have the same "
                "dtype; complex (f-domain) or real (s-domain). Provided:"
                f"sfield: {dtype}; efield: {efield.field.dtype}."
            )
        if efield.frequency is None:
            efield._frequency = sfield._frequency
        d_e = _lib.DeviceArray.from_host(np.asarray(efield.field))
        _lib.check(_lib.load().emg3d_b200_pec_zero(level.handle.ptr, d_e.ptr))
        var.do_return = always_return
        var.user_start = True       # the caller's field is the start vector (see _krylov)
        var.l2 = _dev_residual(level, d_s, d_e, norm=True)
        if var.l2 < var.tol * var.l2_refe:
            var.sslsolver = None
            var.cycle = None
            var.exit_message = "CONVERGED"
            info = "   > NOTHING DONE (provided efield already good enough)\n"

    if var.l2_refe < 100 * np.finfo(float).tiny:
        var.l2_refe = np.nan
        var.sslsolver = None
        var.cycle = None
        var.exit_message = "CONVERGED"
        info = "   > RETURN ZERO E-FIELD (provided sfield is zero)\n"
        efield = fields.Field(model.grid, dtype=dtype, frequency=sfield._frequency)
        d_e = level.new_field()
        var.e_is_zero = False

    header = f"   [hh:mm:ss]  {'rel. error':<22}"
    if var.sslsolver:
        header += f"{'solver':<20}"
        if var.cycle:
            header += f"{'MG':<11} l s"
        var.cprint(header + "\n", 3)
    elif var.cycle:
        var.cprint(header + f"{'[abs. error, last/prev]':>29}   l s\n", 3)

    if var.sslsolver:
        _krylov(level, d_s, d_e, var)
    elif var.cycle:
        _multigrid(level, d_s, d_e, var)

    # Responses at the receivers, sampled where the field is (fields.get_receiver on the device).
    responses = None
    if receivers is not None:
        responses = fields.get_receiver(fields.DeviceField(model.grid, d_e, dtype, sfield._frequency),
                                        receivers, receiver_method)
    if return_field is True:
        # Bring the result back into the (possibly user-provided) host field.
        d_e.download(out=efield.field.view(np.ndarray))
    elif return_field == 'device':
        efield = fields.DeviceField(model.grid, d_e, dtype, sfield._frequency)
    else:
        efield = None

    return _finish(var, efield, info, responses)


def _finish(var, efield, info, responses=None):
    """Closing log lines, info dict and return convention of :func:`solve` (solver.py:407-449)."""
    exit_status = int(var.exit_message != 'CONVERGED')

    if var.verb in [1, 2]:
        _print_one_liner(var, var.l2, True)
    elif var.verb > 2:
        if var.sslsolver:
            info = f"   > Solver steps     : {var.ssl_it}\n"
            if var.cycle:
                info += f"   > MG prec. steps   : {var.it}\n"
        elif var.cycle:
            info = f"   > MG cycles        : {var.it}\n"
        info += f"   > Final rel. error : {var.l2/var.l2_refe:.3e}\n\n"
        info += f":: emg3d END   :: {var.time.now} :: "
        info += f"runtime = {var.time.runtime}\n"
        var.cprint(info, 2)
    elif var.verb == 0 and exit_status == 1:
        var.cprint(f"* WARNING :: {var.exit_message}", -1)

    info_dict = None
    if var.return_info:
        info_dict = {
            'exit': exit_status,
            'exit_message': var.exit_message,
            'abs_error': var.l2,
            'rel_error': var.l2 / var.l2_refe,
            'ref_error': var.l2_refe,
            'tol': var.tol,
            'it_mg': var.it,
            'it_ssl': var.ssl_it,
            'time': var.runtime_at_cycle[-1],
            'runtime_at_cycle': var.runtime_at_cycle,
            'error_at_cycle': var.error_at_cycle,
            'log': var.log_message,
        }

    # the reference's conventions (efield / (efield, info) / info / None), with the responses
    # of `receivers=` after the field
    out = []
    if var.do_return and efield is not None:
        out.append(efield)
    if responses is not None:
        out.append(responses)
    if var.return_info:
        out.append(info_dict)
    if not out:
        return None
    return out[0] if len(out) == 1 else tuple(out)


def solve_source(model, source, frequency, **kwargs):
    """``get_source_field`` followed by :func:`solve` (solver.py:452-467)."""
    sfield = fields.get_source_field(model.grid, source, frequency)
    return solve(model, sfield, **kwargs)


# --------------------------------------------------------------------------- #
# host <-> device adapters for the public sub-routines
# --------------------------------------------------------------------------- #

def _as_level(model, dtype):
    """Device level of a (Volume)model-like object; cached on the object."""
    if isinstance(model, _Level):
        return model
    if isinstance(model, _CoarseModel):          # device-born (restriction): nothing to guard
        return model._b200_level
    # cached on the object, guarded by a checksum over every coefficient (a VolumeModel whose
    # eta / zeta were edited in place gets a new device level)
    stamp = tuple(Workspace._digest(getattr(model, n)) for n in ('eta_x', 'eta_y', 'eta_z', 'zeta'))
    cached = getattr(model, '_b200_level', None)
    if cached is not None and cached[0] == stamp and cached[1].dtype == np.dtype(dtype):
        return cached[1]
    lv = _Level.from_volume_model(model, dtype)
    try:
        model._b200_level = (stamp, lv)
    except AttributeError:
        pass
    return lv


def _up(field):
    if isinstance(field, _lib.DeviceArray):
        return field
    return _lib.DeviceArray.from_host(np.asarray(field.field))


def _down(dev, field):
    if not isinstance(field, _lib.DeviceArray):
        dev.download(out=np.asarray(field.field).view(np.ndarray))


def multigrid(model, sfield, efield, var, **kwargs):
    """Multigrid V/W/F cycling (solver.py:471-649); ``efield`` updated in place."""
    lv = _as_level(model, _field_dtype(sfield)
                   if not isinstance(sfield, _lib.DeviceArray) else sfield.dtype)
    d_s, d_e = _up(sfield), _up(efield)
    _multigrid(lv, d_s, d_e, var, **kwargs)
    _down(d_e, efield)


def krylov(model, sfield, efield, var):
    """Krylov solver with optional MG preconditioner (solver.py:652-784)."""
    lv = _as_level(model, _field_dtype(sfield)
                   if not isinstance(sfield, _lib.DeviceArray) else sfield.dtype)
    d_s, d_e = _up(sfield), _up(efield)
    _krylov(lv, d_s, d_e, var)
    _down(d_e, efield)


def smoothing(model, sfield, efield, nu, lr_dir, order=None):
    """``nu`` Gauss-Seidel sweeps with line relaxation ``lr_dir`` (solver.py:788-846)."""
    lv = _as_level(model, _field_dtype(sfield))
    d_s, d_e = _up(sfield), _up(efield)
    _dev_smoothing(lv, d_s, d_e, nu, lr_dir, core.order_id(order))
    _down(d_e, efield)


class _CoarseModel:
    """Coarse-grid model returned by :func:`restriction` (solver.py:909-926)."""

    def __init__(self, level):
        self._b200_level = level
        self.case = level.case
        self.grid = level.grid
        self._host = {}

    def _get(self, k):
        if k not in self._host:
            lv = self._b200_level
            src = lv.zeta if k == 3 else lv.eta[k]
            for j in range(k if k < 3 else 0):
                if lv.eta[j] is src:
                    self._host[k] = self._get(j)
                    break
            else:
                self._host[k] = src.download().reshape(lv.shape, order='F')
        return self._host[k]

    eta_x = property(lambda self: self._get(0))
    eta_y = property(lambda self: self._get(1))
    eta_z = property(lambda self: self._get(2))
    zeta = property(lambda self: self._get(3))


def restriction(model, sfield, residual, sc_dir):
    """Coarse grid, coarse model and restricted residual (solver.py:849-944)."""
    dtype = _field_dtype(sfield)
    lv = _as_level(model, dtype)
    child = _dev_restriction(lv, _up(residual), sc_dir)
    freq = getattr(sfield, '_frequency', None)
    csfield = fields.Field(child.grid, data=child.s.download(), frequency=freq)
    cefield = fields.Field(child.grid, dtype=dtype, frequency=freq)
    return _CoarseModel(child), csfield, cefield


def prolongation(efield, cefield, sc_dir):
    """``efield += P cefield`` on interior edges (solver.py:947-1019)."""
    g, cg = efield.grid, cefield.grid
    dtype = np.asarray(efield.field).dtype
    cflag = core.SC_FLAGS[int(sc_dir)]
    fine = _lib.LevelHandle(g.h)
    coarse = _lib.LevelHandle(cg.h)
    lo, frac = [], []
    for a, n in enumerate('xyz'):
        li, fr = core.interpolation_table(getattr(g, 'nodes_' + n), getattr(cg, 'nodes_' + n))
        lo.append(li)
        frac.append(fr)
    weights = [None] * 9
    for a in range(3):
        if cflag[a]:
            weights[3 * a:3 * a + 3] = [np.zeros(cg.shape_nodes[a])] * 3
    coarse.link(fine, cflag, weights, lo, frac)
    _set_dtype(coarse, dtype)
    d_e, d_c = _up(efield), _up(cefield)
    _lib.check(_lib.load().emg3d_b200_prolong(coarse.ptr, d_e.ptr, d_c.ptr))
    _down(d_e, efield)


def residual(model, sfield, efield, norm=False):
    """Residual field ``s - A e`` or its l2-norm (solver.py:1022-1070)."""
    dtype = _field_dtype(sfield)
    lv = _as_level(model, dtype)
    d_s, d_e = _up(sfield), _up(efield)
    if norm:
        return _dev_residual(lv, d_s, d_e, norm=True)
    r = _dev_residual(lv, d_s, d_e)
    return fields.Field(sfield.grid, data=r.download(),
                        frequency=getattr(sfield, '_frequency', None))


# =========================================================================== #
# Cycling on the device
# =========================================================================== #

def _multigrid(lv, s, e, var, level=0, new_cycmax=0):
    """Recursive multigrid cycle on device-resident data (solver.py:471-649)."""
    order = _order(var)
    it = 0
    coarsest = level == var.clevel[var.sc_dir]
    if coarsest:
        cycmax = 1
    elif new_cycmax == 0 or var.cycle != 'F':
        cycmax = var.cycmax
    else:
        cycmax = new_cycmax
    cyc = 0

    # The reference evaluates the residual norm on entry at every level but uses
    # it on level 0 only (and for verb > 4 printing); skip the unused ones.
    want_norm = level == 0 or var.verb > 4
    if level == 0 and getattr(var, 'e_is_zero', False):
        # The caller guarantees e = 0, so r = s and ||r|| = ||s|| (the reference runs
        # amat_x on the zero field here, solver.py:530): one vector norm, or nothing
        # at all when the caller knows ||s||, instead of a residual evaluation.
        known = getattr(var, 's_norm', None)
        l2_last = _Vec(lv.cplx, lv.n_edges).norm(s) if known is None else float(known)
        var.e_is_zero = False
    else:
        l2_last = _dev_residual(lv, s, e, norm=True) if want_norm else 0.0
    l2_stag = np.ones(var.maxcycle) * l2_last

    if var.first_cycle and var.verb > 3:
        var.level_all.append(level)

    if level == 0:
        var.cprint("     it cycmax               error", 4)
        var.cprint("      level [  dimension  ]            info\n", 4)
        if var.verb > 4:
            _print_gs_info(var, it, level, cycmax, lv.grid, l2_last, "initial error")

    if level == 0 and var.nu_init > 0:
        _dev_smoothing(lv, s, e, var.nu_init, var.lr_dir, order)
        if var.verb > 4:
            _print_gs_info(var, it, level, cycmax, lv.grid,
                           _dev_residual(lv, s, e, norm=True), "initial smoothing")

    while level == 0 or (level > 0 and it < cycmax):
        l2_prev = l2_last
        l2_stag[(it - 1) % var.maxcycle] = l2_last

        if level == var.clevel[var.sc_dir]:
            _dev_smoothing(lv, s, e, var.nu_coarse, var.lr_dir, order)
            if var.verb > 4:
                _print_gs_info(var, it, level, cycmax, lv.grid,
                               _dev_residual(lv, s, e, norm=True), "coarsest level")
        else:
            if var.nu_pre > 0:
                _dev_smoothing(lv, s, e, var.nu_pre, var.lr_dir, order)
                if var.verb > 4:
                    _print_gs_info(var, it, level, cycmax, lv.grid,
                                   _dev_residual(lv, s, e, norm=True), "pre-smoothing")

            sc_dir = _current_sc_dir(var.sc_dir, lv.grid)
            res = _dev_residual(lv, s, e)
            child = _dev_restriction(lv, res, sc_dir)
            _subcycle(child, var, level + 1, cycmax - cyc)
            _dev_prolongation(child, e)

            if var.first_cycle and var.verb > 3:
                var.level_all.append(level)

            if var.nu_post > 0:
                _dev_smoothing(lv, s, e, var.nu_post, var.lr_dir, order)
                if var.verb > 4:
                    _print_gs_info(var, it, level, cycmax, lv.grid,
                                   _dev_residual(lv, s, e, norm=True), "post-smoothing")

        it += 1
        if level == 0:
            var.it += 1

        if level > 0:
            cyc += 1
        else:
            l2_last = _dev_residual(lv, s, e, norm=True)
            _print_cycle_info(var, l2_last, l2_prev)
            if var.sc_cycle:
                var.sc_dir = next(var.sc_cycle)
            if var.lr_cycle:
                var.lr_dir = next(var.lr_cycle)
            if _terminate(var, l2_last, l2_stag[(it - 1) % var.maxcycle], it):
                break

    var.l2 = l2_last


# Coarse sub-cycles as CUDA graphs.  Below the finest level a visit is a fixed
# sequence of launches on level-owned buffers (no norms, no host decisions), and
# the coarse levels are launch-latency bound: W-cycles visit level l 2^l times.
# The first visit of a (level, settings) combination runs eagerly (it creates the
# child levels, cached diagonals and factorisations), the second is captured, and
# from then on the graph is replayed.  EMG3D_B200_GRAPHS=0 disables this.
GRAPHS = os.environ.get('EMG3D_B200_GRAPHS', '1') != '0'
# only levels at or below this many cells are worth a graph of their own
_GRAPH_MAX_CELLS = 160 ** 3


def _subcycle(child, var, level, new_cycmax):
    use = (GRAPHS and var.verb <= 3 and not getattr(var, '_capturing', False)
           and child.grid.n_cells <= _GRAPH_MAX_CELLS)
    if not use:
        return _multigrid(child, child.s, child.e, var, level=level, new_cycmax=new_cycmax)
    key = (level, int(new_cycmax), var.cycle, int(var.sc_dir), int(var.lr_dir), _order(var),
           int(var.nu_pre), int(var.nu_post), int(var.nu_coarse), tuple(var.clevel))
    graphs = child.__dict__.setdefault('_graphs', {})
    g = graphs.get(key)
    if g is None:                                   # first visit: eager, builds the caches
        graphs[key] = False
        return _multigrid(child, child.s, child.e, var, level=level, new_cycmax=new_cycmax)
    if g is False:                                  # second visit: capture
        var._capturing = True
        try:
            with _lib.Graph() as g:
                _multigrid(child, child.s, child.e, var, level=level, new_cycmax=new_cycmax)
        finally:
            var._capturing = False
        graphs[key] = g
    g.launch()


class _Vec:
    """Device vector algebra for the Krylov iteration."""

    def __init__(self, cplx, n):
        self.cplx, self.n = int(cplx), int(n)
        self.lib = _lib.load()

    def dot(self, x, y, conj=True):
        out = (_lib.c_double * 2)()
        _lib.check(self.lib.emg3d_b200_dot_host(self.cplx, self.n, x.ptr, y.ptr,
                                                int(conj), out))
        return complex(out[0], out[1]) if self.cplx else out[0]

    def norm(self, x):
        out = (_lib.c_double * 2)()
        _lib.check(self.lib.emg3d_b200_dot_host(self.cplx, self.n, x.ptr, x.ptr, 1, out))
        return float(np.sqrt(out[0]))

    def axpby(self, a, x, b, y):
        """y = a x + b y"""
        a, b = complex(a), complex(b)
        _lib.check(self.lib.emg3d_b200_axpby(self.cplx, self.n, a.real, a.imag, x.ptr,
                                             b.real, b.imag, y.ptr))


class _DeviceOps:
    """What the device-resident Krylov solvers need from a (single- or multi-GPU) backend:
    vector algebra, the operator, the multigrid preconditioner and the residual norm."""

    def __init__(self, lv):
        self.lv, self.vec = lv, _Vec(lv.cplx, lv.n_edges)
        self.lib = _lib.load()

    def new(self):
        return _lib.DeviceArray(self.lv.n_edges, self.lv.dtype)

    def norm(self, x):
        return self.vec.norm(x)

    def dot(self, x, y):
        return self.vec.dot(x, y)

    def axpby(self, a, x, b, y):
        self.vec.axpby(a, x, b, y)

    def matvec(self, src, dst):
        _lib.check(self.lib.emg3d_b200_apply(self.lv.handle.ptr, src.ptr, dst.ptr))

    def psolve(self, src, dst, var):
        if var.cycle:
            dst.zero()
            var.e_is_zero, var.s_norm = True, None      # preconditioner starts from zero
            _multigrid(self.lv, src, dst, var)
        else:
            dst.copy_from(src)

    def residual_norm(self, s, x):
        return _dev_residual(self.lv, s, x, norm=True)


def _krylov(lv, s, e, var, ops=None):
    """Krylov solvers with multigrid as preconditioner (solver.py:652-784).

    ``ops``: backend (default: this GPU, :class:`_DeviceOps`); the multi-GPU driver passes its own.
    """
    ops = _DeviceOps(lv) if ops is None else ops

    def record(x):
        var.ssl_it += 1
        var.runtime_at_cycle = np.r_[var.runtime_at_cycle, var.time.elapsed]
        var.l2 = ops.residual_norm(s, x)
        var.error_at_cycle = np.r_[var.error_at_cycle, var.l2]
        if var.verb > 3:
            log = f"   [{var.time.now}]   {var.l2/var.l2_refe:.3e} "
            log += f" after {var.ssl_it:3} {var.sslsolver}-cycles"
            if var.ssl_it == 1 and var.it == 0 and var.cycle is not None:
                log += "\n"
            var.cprint(log, 3)
        elif var.verb in [2, 3]:
            _print_one_liner(var, var.l2)

    # When the preconditioner diverges the reference never assigns the solver's iterate
    # (solver.py:763-768): a start field supplied by the caller keeps its (PEC-cleaned) input
    # and only the default zero start "returns zero".  The device iterate is updated in place,
    # so a supplied start field is kept aside.
    x0 = e.copy() if getattr(var, 'user_start', False) else None
    try:
        if var.sslsolver == 'bicgstab':
            i = _bicgstab(ops, s, e, var, record)
        elif var.sslsolver == 'cgs':
            i = _cgs(ops, s, e, var, record)
        elif lv is None or os.environ.get('EMG3D_B200_GCROT', 'device') != 'host':
            i = _gcrotmk(ops, s, e, var, record)
        else:
            i = _scipy_krylov(lv, s, e, var, record)
    except _ConvergenceError:
        i = -1
        if x0 is None:
            e.zero()
        else:
            e.copy_from(x0)
        var.exit_message += " (returned field is zero)"
    del x0

    if var.verb == 3:
        pre = 50 * " " + "\r"
    else:
        pre = "\n"
    pre += "   > "
    if i < 0:
        if var.exit_message == '':
            var.exit_message = f"Error in {var.sslsolver} ({i})"
        pre = "\n* ERROR   :: "
    elif i > 0:
        var.exit_message = "MAX. ITERATION REACHED, NOT CONVERGED"
    else:
        var.exit_message = "CONVERGED"
    var.cprint(pre + var.exit_message, 2)


_KRYLOV_DEBUG = bool(int(__import__('os').environ.get('EMG3D_B200_KRYLOV_DEBUG', '0')))


def _bicgstab(ops, b, x, var, callback):
    """Preconditioned BiCGSTAB on the device(s).

    Same recurrence, breakdown tests and stopping rule as
    ``scipy.sparse.linalg.bicgstab`` (SciPy 1.18, ``_isolve/iterative.py``),
    which the reference calls at solver.py:763-765 with ``rtol=tol``,
    ``atol=1e-30``, ``maxiter=ssl_maxit`` and a callback per iteration.
    Returns SciPy's ``info`` code.
    """
    bnrm2 = ops.norm(b)
    atol = max(1e-30, float(var.tol) * bnrm2)
    if bnrm2 == 0:
        x.copy_from(b)
        return 0
    rhotol = np.finfo(np.float64).eps ** 2
    omegatol = rhotol

    r, rtilde, p, v, s_, t, phat, shat = (ops.new() for _ in range(8))
    # r = b - A x (x may be a user-provided start field)
    if ops.norm(x) > 0:
        ops.matvec(x, r)
        ops.axpby(1.0, b, -1.0, r)
    else:
        r.copy_from(b)
    rtilde.copy_from(r)
    rho_prev = omega = alpha = None

    for iteration in range(var.ssl_maxit):
        if ops.norm(r) < atol:
            return 0
        rho = ops.dot(rtilde, r)
        if abs(rho) < rhotol:
            return -10
        if iteration > 0:
            if abs(omega) < omegatol:
                return -11
            beta = (rho / rho_prev) * (alpha / omega)
            ops.axpby(-omega, v, 1.0, p)      # p -= omega v
            ops.axpby(1.0, r, beta, p)        # p = beta p + r
        else:
            p.copy_from(r)
        ops.psolve(p, phat, var)
        ops.matvec(phat, v)
        rv = ops.dot(rtilde, v)
        if rv == 0:
            return -11
        alpha = rho / rv
        ops.axpby(-alpha, v, 1.0, r)          # r -= alpha v
        s_.copy_from(r)
        if ops.norm(s_) < atol:
            ops.axpby(alpha, phat, 1.0, x)
            return 0
        ops.psolve(s_, shat, var)
        ops.matvec(shat, t)
        omega = ops.dot(t, s_) / ops.dot(t, t)
        ops.axpby(alpha, phat, 1.0, x)
        ops.axpby(omega, shat, 1.0, x)
        ops.axpby(-omega, t, 1.0, r)
        rho_prev = rho
        if _KRYLOV_DEBUG:                     # recurrence residual against the true one
            tr = ops.new()
            ops.matvec(x, tr)
            ops.axpby(1.0, b, -1.0, tr)
            ops.axpby(-1.0, r, 1.0, tr)
            print(f"   bicgstab {iteration}: |r| {ops.norm(r):.3e} |b - A x - r| {ops.norm(tr):.3e} "
                  f"rho {rho:.3e} alpha {alpha:.3e} omega {omega:.3e}", flush=True)
        callback(x)
    return var.ssl_maxit


def _cgs(ops, b, x, var, callback):
    """Preconditioned CGS on the device(s): recurrence, breakdown tests, stopping rule and the
    recomputed true residual of ``scipy.sparse.linalg.cgs`` (SciPy 1.18), which the reference
    calls for ``sslsolver='cgs'`` (solver.py:763-765).  Returns SciPy's ``info`` code."""
    bnrm2 = ops.norm(b)
    atol = max(1e-30, float(var.tol) * bnrm2)
    if bnrm2 == 0:
        x.copy_from(b)
        return 0
    rhotol = np.finfo(np.float64).eps ** 2
    r, rtilde, p, u, q, phat, vhat, uq, uhat = (ops.new() for _ in range(9))

    def true_residual():                      # r = b - A x
        ops.matvec(x, r)
        ops.axpby(1.0, b, -1.0, r)

    if ops.norm(x) > 0:
        true_residual()
    else:
        r.copy_from(b)
    rtilde.copy_from(r)
    rho_prev = None
    for iteration in range(var.ssl_maxit):
        if ops.norm(r) < atol:
            return 0
        rho = ops.dot(rtilde, r)
        if abs(rho) < rhotol:
            return -10
        if iteration > 0:
            beta = rho / rho_prev
            u.copy_from(r)
            ops.axpby(beta, q, 1.0, u)        # u = r + beta q
            ops.axpby(1.0, q, beta, p)        # p = beta p + q
            ops.axpby(1.0, u, beta, p)        # p = beta (beta p + q) + u
        else:
            p.copy_from(r)
            u.copy_from(r)
        ops.psolve(p, phat, var)
        ops.matvec(phat, vhat)
        rv = ops.dot(rtilde, vhat)
        if rv == 0:
            return -11
        alpha = rho / rv
        q.copy_from(u)
        ops.axpby(-alpha, vhat, 1.0, q)       # q = u - alpha vhat
        uq.copy_from(u)
        ops.axpby(1.0, q, 1.0, uq)
        ops.psolve(uq, uhat, var)
        ops.axpby(alpha, uhat, 1.0, x)
        true_residual()
        rho_prev = rho
        callback(x)
    return var.ssl_maxit


def _gcrotmk(ops, b, x, var, callback, m=20, k=None):
    """Flexible GCROT(m,k) on the device(s), right-preconditioned by multigrid.

    The outer iteration, the inner FGMRES/Arnoldi process with its projection against the
    carried ``C`` vectors, the 'oldest' truncation, the stopping rule (true residual recomputed
    before convergence is declared) and the callback at the start of every outer iteration are
    those of ``scipy.sparse.linalg.gcrotmk`` (SciPy 1.18, ``_isolve/_gcrotmk.py``) as the
    reference calls it at solver.py:763-765 (``m=20``, ``k=m``, no recycled ``CU``, ``rtol=tol``,
    ``atol=1e-30``, ``maxiter=ssl_maxit``).  Vectors live on the device(s); only the small
    Hessenberg problem (QR update, least squares) is solved on the host.  Returns SciPy's
    ``info`` code (0 converged, else the number of outer iterations done).
    """
    from scipy.linalg import LinAlgError, lstsq, qr_insert
    k = m if k is None else k
    free = []                                 # work vectors are recycled between outer iterations

    def new():
        return free.pop() if free else ops.new()

    def scal(a, v):                           # v *= a  (0 * b + a * v: no aliased operands)
        ops.axpby(0.0, b, a, v)

    def axpy(a, v, w):                        # w += a v
        ops.axpby(a, v, 1.0, w)

    b_norm = ops.norm(b)
    if b_norm == 0:
        x.copy_from(b)
        return 0
    dtype = np.dtype(complex) if isinstance(ops.dot(b, b), complex) else np.dtype(float)
    beta_tol = max(1e-30, float(var.tol) * b_norm)
    eps = np.finfo(float).eps

    def true_residual(r):                     # r = b - A x
        ops.matvec(x, r)
        ops.axpby(1.0, b, -1.0, r)

    r = new()
    if ops.norm(x) > 0:
        true_residual(r)
    else:
        r.copy_from(b)

    def fgmres(v0, mm, atol, cs):
        """L A Z = C B + V H with H kept as Q R; returns Q, R, B, vs, zs, y."""
        vs, zs = [v0], []
        B = np.zeros((len(cs), mm), dtype=dtype)
        Q, R = np.ones((1, 1), dtype=dtype), np.zeros((1, 0), dtype=dtype)
        breakdown = False
        for j in range(mm):
            z, w = new(), new()
            ops.psolve(vs[-1], z, var)
            ops.matvec(z, w)
            w_norm = ops.norm(w)
            for i, c in enumerate(cs):        # (1 - C C^H) A z
                B[i, j] = alpha = ops.dot(c, w)
                axpy(-alpha, c, w)
            hcur = np.zeros(j + 2, dtype=dtype)
            for i, v in enumerate(vs):        # modified Gram-Schmidt against V
                hcur[i] = alpha = ops.dot(v, w)
                axpy(-alpha, v, w)
            h_last = ops.norm(w)
            hcur[j + 1] = h_last
            with np.errstate(over='ignore', divide='ignore'):
                alpha = np.float64(1.0) / np.float64(h_last)
            if np.isfinite(alpha):
                scal(float(alpha), w)
            if not (h_last > eps * w_norm):   # w in the span of the previous vectors, or nan
                breakdown = True
            vs.append(w)
            zs.append(z)
            Q2 = np.zeros((j + 2, j + 2), dtype=dtype, order='F')
            Q2[:j + 1, :j + 1] = Q
            Q2[j + 1, j + 1] = 1
            R2 = np.zeros((j + 2, j), dtype=dtype, order='F')
            R2[:j + 1, :] = R
            Q, R = qr_insert(Q2, R2, hcur, j, which='col', overwrite_qru=True, check_finite=False)
            if abs(Q[0, -1]) < atol or breakdown:
                break
        if not np.isfinite(R[j, j]):
            free.extend(vs + zs)
            raise LinAlgError()
        y = lstsq(R[:j + 1, :j + 1], Q[0, :j + 1].conj())[0]
        return Q, R, B[:, :j + 1], vs, zs, y

    CU = []
    for j_outer in range(var.ssl_maxit):
        callback(x)
        beta = ops.norm(r)
        if beta <= beta_tol and (j_outer > 0 or CU):
            true_residual(r)                  # recomputed: the recurrence residual drifts
            beta = ops.norm(r)
        if beta <= beta_tol:
            return 0
        ml = m + max(k - len(CU), 0)
        v0 = new()
        v0.copy_from(r)
        scal(1.0 / beta, v0)
        try:
            Q, R, B, vs, zs, y = fgmres(v0, ml, beta_tol / beta, [c for c, _ in CU])
        except LinAlgError:                   # over/underflow, nan from the operator
            return j_outer + 1
        y = y * beta
        # new outer pair:  ux = (Z - U B) y,  cx = V H y = A ux  (zs[0], vs[0] are reused)
        ux = zs[0]
        scal(y[0], ux)
        for z, yc in zip(zs[1:], y[1:]):
            axpy(yc, z, ux)
        for (_, u), byc in zip(CU, B.dot(y)):
            axpy(-byc, u, ux)
        with np.errstate(invalid='ignore'):
            hy = Q.dot(R.dot(y))
        cx = vs[0]
        scal(hy[0], cx)
        for v, hyc in zip(vs[1:], hy[1:]):
            axpy(hyc, v, cx)
        free.extend(vs[1:] + zs[1:])
        with np.errstate(over='ignore', divide='ignore'):
            alpha = np.float64(1.0) / np.float64(ops.norm(cx))
        if not np.isfinite(alpha):            # cannot update: skip this pair
            free.extend((cx, ux))
            continue
        scal(float(alpha), cx)
        scal(float(alpha), ux)
        gamma = ops.dot(cx, r)
        axpy(-gamma, cx, r)
        axpy(gamma, ux, x)
        while len(CU) >= k and CU:            # truncate='oldest'
            free.extend(CU.pop(0))
        CU.append((cx, ux))
    return var.ssl_maxit


def _scipy_krylov(lv, s, e, var, callback):
    """GCROT(m,k) driven by SciPy on the host: the GPU applies A and the preconditioner, full
    vectors cross PCIe per application.  Kept as a cross-check of the device-resident
    :func:`_gcrotmk` (the default; also what a multi-GPU solve runs): ``EMG3D_B200_GCROT=host``
    selects it (tests/test_gpu_solver_gcrot.py runs both against the reference)."""
    lib = _lib.load()
    n = lv.n_edges
    d_in, d_out = lv.new_field(False), lv.new_field(False)

    def amatvec(x):
        d_in.upload(np.ascontiguousarray(x, dtype=lv.dtype))
        _lib.check(lib.emg3d_b200_apply(lv.handle.ptr, d_in.ptr, d_out.ptr))
        return d_out.download()

    def mg_matvec(b):
        d_in.upload(np.ascontiguousarray(b, dtype=lv.dtype))
        d_out.zero()
        _multigrid(lv, d_in, d_out, var)
        return d_out.download()

    A = _ssl.LinearOperator((n, n), dtype=lv.dtype, matvec=amatvec)
    M = _ssl.LinearOperator((n, n), dtype=lv.dtype, matvec=mg_matvec) if var.cycle else None
    d_x = lv.new_field(False)

    def cb(x):
        d_x.upload(np.ascontiguousarray(x, dtype=lv.dtype))
        callback(d_x)

    x, i = getattr(_ssl, var.sslsolver)(
        A=A, b=s.download(), x0=e.download(), rtol=var.tol, maxiter=var.ssl_maxit,
        atol=1e-30, M=M, callback=cb)
    e.upload(np.ascontiguousarray(x, dtype=lv.dtype))
    return i


# =========================================================================== #
# Parameters
# =========================================================================== #

@dataclass
class MGParameters:
    """Settings and running state of one solve (solver.py:1074-1381)."""

    verb: int
    sslsolver: Union[str, bool]
    semicoarsening: Union[int, bool]
    linerelaxation: Union[int, bool]
    shape_cells: tuple
    cycle: Union[str, None] = 'F'
    tol: float = 1e-6
    maxit: int = 50
    nu_init: int = 0
    nu_pre: int = 2
    nu_coarse: int = 1
    nu_post: int = 2
    clevel: int = -1
    return_info: bool = False
    log: int = 0

    def __post_init__(self):
        self.level_all = list()
        self.first_cycle = True
        self.it = 0
        self.ssl_it = 0
        self.l2 = 1.0
        self.l2_refe = 1.0
        self.order = None
        self._max_level()
        self.exit_message = ''
        self.log_message = ''
        self.time = Timer()
        self.runtime_at_cycle = np.array([0.])
        self.error_at_cycle = np.array([0.])
        self.do_return = True
        self._semicoarsening()
        self._linerelaxation()
        self._solver_and_cycle()

    def __repr__(self):
        nx, ny, nz = self.shape_cells
        cl = self._repr_clevel
        return (
            f"   MG-cycle       : {self.cycle!r:17}"
            f"   sslsolver : {self.sslsolver!r}\n"
            f"   semicoarsening : {self._repr_sc_dir:17}"
            f"   tol       : {self.tol}\n"
            f"   linerelaxation : {self._repr_lr_dir:17}"
            f"   maxit     : {self._repr_maxit}\n"
            f"   nu_{{i,1,c,2}}   : {self.nu_init}, {self.nu_pre},"
            f" {self.nu_coarse}, {self.nu_post}       "
            f"   verb      : {self.verb}\n"
            f"   Original grid  : {nx:3} x {ny:3} x {nz:3}     =>"
            f" {nx*ny*nz:,} cells\n"
            f"   Coarsest grid  : {cl['shape_cells'][0]:3} x"
            f" {cl['shape_cells'][1]:3} x {cl['shape_cells'][2]:3}  "
            f"   => {cl['n_cells']:,} cells\n"
            f"   Coarsest level : {cl['clevel'][0]:3} ; {cl['clevel'][1]:3}"
            f" ;{cl['clevel'][2]:4}   {cl['message']}\n"
        )

    def cprint(self, info, verbosity, **kwargs):
        """Print and/or log ``info`` if ``verb > verbosity`` (solver.py:1181-1200)."""
        if self.verb > verbosity:
            if self.log != 0:
                self.log_message += str(info) + '\n'
            if self.log >= 0:
                print(info, **kwargs)

    def _max_level(self):
        """How often each axis can be halved (solver.py:1202-1270)."""
        user = np.inf if self.clevel < 0 else self.clevel
        halvings = np.zeros(3, dtype=np.int64)
        for a, n in enumerate(self.shape_cells):
            while n % 2 == 0 and n > 2:
                halvings[a] += 1
                n //= 2
            if -1 < self.clevel < halvings[a]:
                halvings[a] = self.clevel
        cx, cy, cz = (int(v) for v in halvings)
        # coarsest level for sc_dir = 0, 1, 2, 3
        self.clevel = np.array([max(cx, cy, cz), max(cy, cz), max(cx, cz), max(cx, cy)])
        coarsest = tuple(int(n / 2**c) for n, c in zip(self.shape_cells, halvings))
        self._repr_clevel = {'n_cells': int(np.prod(coarsest)), 'shape_cells': coarsest,
                             'clevel': halvings}
        too_big = any(c < user and m > 7 for c, m in zip(halvings, coarsest))
        too_few = any(halvings < min(user, 3))
        self._repr_clevel['message'] = (
            "  :: Grid not optimal for MG solver ::" if too_big or too_few else "")
        if np.any(np.array(self.shape_cells) < 2):
            raise ValueError(
                "Nr. of cells must be at least two in each direction "
                "Provided shape: ({self.shape_cells[0]}, "
                f"{self.shape_cells[1]}, {self.shape_cells[2]})."
            )

    @staticmethod
    def _direction_cycle(flag, default, upper):
        """(sequence, iterator-or-False) from a True/False/int/digits flag."""
        if flag is True:
            seq = np.array(default)
            return seq, itertools.cycle(seq)
        if flag in np.arange(upper + 1):
            return np.array([int(flag)]), False
        seq = np.array([int(c) for c in str(abs(flag))])
        return seq, itertools.cycle(seq)

    def _semicoarsening(self):
        seq, cyc = self._direction_cycle(self.semicoarsening, [1, 2, 3], 3)
        if np.any(seq < 0) or np.any(seq > 3):
            raise ValueError(
                "`semicoarsening` must be one of {False;True;0;1;2;3}. "
                "Or a combination of {0;1;2;3} to cycle, e.g. 1213. "
                f"Provided: {self.semicoarsening}."
            )
        self.sc_cycle = cyc
        self.sc_dir = next(cyc) if cyc else seq[0]
        self.semicoarsening = self.sc_dir != 0
        self._repr_sc_dir = f"{self.semicoarsening} {seq}"
        self.raw_sc_cycle = seq

    def _linerelaxation(self):
        seq, cyc = self._direction_cycle(self.linerelaxation, [4, 5, 6], 7)
        if np.any(seq < 0) or np.any(seq > 7):
            raise ValueError(
                "`linerelaxation` must be one of "
                "{False;True;0;1;2;3;4;5;6;7}. Or a combination of "
                "{1;2;3;4;5;6;7} to cycle, e.g. 1213. "
                f"Provided: {self.linerelaxation}."
            )
        self.lr_cycle = cyc
        self.lr_dir = next(cyc) if cyc else seq[0]
        self.linerelaxation = self.lr_dir != 0
        self._repr_lr_dir = f"{self.linerelaxation} {seq}"
        self.raw_lr_cycle = seq

    # Accepted Krylov wrappers / cycle types (solver.py:1341-1381).  The messages are pinned by
    # the reference's tests (tests/test_solver.py); the checks are table-driven here.
    _SSLSOLVERS = ('bicgstab', 'cgs', 'gcrotmk')
    _CYCMAX = {'V': 1, 'W': 2, 'F': 2, None: 1}

    def _solver_and_cycle(self):
        ssl = {True: 'bicgstab', False: False}.get(self.sslsolver, self.sslsolver) \
            if isinstance(self.sslsolver, bool) else self.sslsolver
        problems = (
            (ssl is not False and ssl not in self._SSLSOLVERS,
             f"`sslsolver` must be True, False, or one of {list(self._SSLSOLVERS)}. "
             f"Provided: {self.sslsolver!r}."),
            (self.cycle not in self._CYCMAX,
             "`cycle` must be one of {'F';'V';'W';None}. " f"Provided: {self.cycle}."),
            (not ssl and not self.cycle,
             "At least `cycle` or `sslsolver` is required. Provided"
             f"input: cycle={self.cycle}; sslsolver={ssl}."),
        )
        for bad, message in problems:
            if bad:
                raise ValueError(message)
        self.sslsolver = ssl
        self.cycmax = self._CYCMAX[self.cycle]
        # With a Krylov wrapper `maxit` bounds the Krylov iterations and each preconditioner
        # call runs one round of the sc / lr cycling patterns (solver.py:1370-1381).
        self.maxcycle = max(len(self.raw_sc_cycle), len(self.raw_lr_cycle))
        self._repr_maxit = f"{self.maxit}"
        self.ssl_maxit = self.maxit if ssl else 0
        if ssl and self.cycle is not None:
            self.maxit = self.maxcycle
            self._repr_maxit += f" ({self.maxit})"


class RegularGridProlongator:
    """Bilinear prolongation of 2-D slices from a coarse to a fine tensor grid.

    Interface of the reference's class (solver.py:1385-1478): initialise with
    coarse ``(cx, cy)`` and fine ``(x, y)`` coordinates; calling it with coarse
    values of shape ``(cx.size, cy.size)`` returns the fine values flattened in
    Fortran order.  (Inside the solver the same tables are applied by the CUDA
    prolongation kernel; this host class serves callers of the public name.)
    """

    def __init__(self, cx, cy, x, y):
        self._ix, self._tx = core.interpolation_table(x, cx)
        self._iy, self._ty = core.interpolation_table(y, cy)
        self.size = self._ix.size * self._iy.size

    def __call__(self, values):
        v = np.asarray(values)
        ix, iy = self._ix[:, None], self._iy[None, :]
        tx, ty = self._tx[:, None], self._ty[None, :]
        out = (v[ix, iy] * ((1 - tx) * (1 - ty)) + v[ix, iy + 1] * ((1 - tx) * ty) +
               v[ix + 1, iy] * (tx * (1 - ty)) + v[ix + 1, iy + 1] * (tx * ty))
        return out.ravel('F')


# =========================================================================== #
# Helpers
# =========================================================================== #

def _current_sc_dir(sc_dir, grid):
    """Semicoarsening pattern usable on this grid (solver.py:1482-1531)."""
    n = grid.shape_cells
    keep = [n[a] % 2 != 0 or n[a] < 3 or sc_dir == a + 1 for a in range(3)]
    table = {(False, False, False): 0, (True, False, False): 1,
             (False, True, False): 2, (False, False, True): 3,
             (False, True, True): 4, (True, False, True): 5,
             (True, True, False): 6, (True, True, True): 6}
    return table[tuple(bool(k) for k in keep)]


def _current_lr_dir(lr_dir, grid):
    """Drop line directions with only two cells (solver.py:1534-1588)."""
    n = grid.shape_cells
    dirs = tuple(d for d in _LR_DIRS[int(lr_dir)] if n[d - 1] != 2)
    return np.array(_LR_CODE[dirs])


class _ConvergenceError(Exception):
    """Raised inside a Krylov run when the preconditioner diverges/stagnates."""


def _terminate(var, l2_last, l2_stag, it):
    """Termination criteria of the multigrid iteration (solver.py:1591-1664).

    One ordered rule table: (condition, message, fatal).  The first rule that holds ends the
    iteration; a *fatal* outcome inside a Krylov run (multigrid as preconditioner) aborts the
    Krylov solver through `_ConvergenceError`; reaching `maxit` as preconditioner is the normal
    end of a preconditioner call and sets no message.  ``l2_refe`` is the global ||b||, also
    inside the preconditioner (solver.py:1622 with 714-719).
    """
    rules = (
        (l2_last < var.tol * var.l2_refe, "CONVERGED", False),
        (l2_last > 10 * var.l2_refe or not np.isfinite(l2_last), "DIVERGED", True),
        (it > 2 and l2_last >= l2_stag, "STAGNATED", True),
        (it == var.maxit, None if var.sslsolver else "MAX. ITERATION REACHED, NOT CONVERGED", False),
    )
    hit = next((rule for rule in rules if rule[0]), None)
    if hit is None:
        return False
    _, message, fatal = hit
    if message is not None:
        var.exit_message = message
    if var.sslsolver:
        if fatal:
            raise _ConvergenceError
        return True
    lead = {3: 50 * " " + "\r"}.get(var.verb, "\n" if var.verb < 5 else "")
    var.cprint(lead + "   > " + var.exit_message, 2)
    return True


def _restrict_model_parameters(param, sc_dir):
    """Coarse cell = sum of its fine cells, on the device (solver.py:1667-1718)."""
    param = np.asfortranarray(param)
    cflag = core.SC_FLAGS[int(sc_dir)]
    cshape = tuple(n // 2 if f else n for n, f in zip(param.shape, cflag))
    fine = _lib.LevelHandle([np.ones(n) for n in param.shape])
    coarse = _lib.LevelHandle([np.ones(n) for n in cshape])
    dummy = [(np.zeros(n + 1, np.int32), np.zeros(n + 1)) for n in param.shape]
    weights = [None] * 9
    for a in range(3):
        if cflag[a]:
            weights[3 * a:3 * a + 3] = [np.zeros(cshape[a] + 1)] * 3
    coarse.link(fine, cflag, weights, [d[0] for d in dummy], [d[1] for d in dummy])
    d_p = _lib.DeviceArray.from_host(param)
    d_c = _lib.DeviceArray(int(np.prod(cshape)), param.dtype)
    _lib.check(_lib.load().emg3d_b200_restrict_cells(
        coarse.ptr, int(param.dtype.kind == 'c'), d_p.ptr, d_c.ptr))
    return d_c.download().reshape(cshape, order='F')


def _get_restriction_weights(grid, cgrid, sc_dir):
    """Restriction weights per axis; dummies where not coarsened (solver.py:1721-1780)."""
    cflag = core.SC_FLAGS[int(sc_dir)]
    out = []
    for a, n in enumerate('xyz'):
        if cflag[a]:
            out.append(core.restrict_weights(
                getattr(grid, 'nodes_' + n), getattr(grid, 'cell_centers_' + n), grid.h[a],
                getattr(cgrid, 'nodes_' + n), getattr(cgrid, 'cell_centers_' + n), cgrid.h[a]))
        else:
            zeros = np.zeros(grid.shape_nodes[a], dtype=np.float64)
            out.append((zeros, np.ones(grid.shape_nodes[a], dtype=np.float64), zeros))
    return tuple(out)


# ---- log output (formats pinned by the reference's tests) -------------------

def _print_cycle_info(var, l2_last, l2_prev):
    """End-of-cycle log line and, once, the cycle diagram (solver.py:1788-1862)."""
    var.runtime_at_cycle = np.r_[var.runtime_at_cycle, var.time.elapsed]
    var.error_at_cycle = np.r_[var.error_at_cycle, l2_last]

    if var.verb in [2, 3]:
        _print_one_liner(var, l2_last)
    if var.verb < 4:
        return
    info = "\n" if var.verb > 4 else ""

    if var.first_cycle:
        lv = np.array(var.level_all, dtype=np.int64)
        depth = np.max(lv)
        step = ((lv[1:] + lv[:-1]) // 2 + 1) * (lv[1:] - lv[:-1])   # +down, -up
        shown = min(len(step), 70)
        rows = ["       h_\n"]
        for cl in range(depth):
            row = f"   {2**(cl+1):4}h_ "
            for v in range(shown):
                row += " " if abs(step[v]) != cl + 1 else "\\" if step[v] > 0 else "/"
            rows.append(row + ("\n" if cl < depth - 1 else ""))
        info += "".join(rows) + "\n\n"
        if len(step) > 70:
            info += "  (Cycle-QC restricted to first 70 steps of "
            info += f"{len(step)} steps.)\n"
        var.first_cycle = False

    info += f"   [{var.time.now}]   {l2_last/var.l2_refe:.3e}  "
    if var.sslsolver:
        info += f"after {19*' '} {var.it:3} {var.cycle}-cycles "
    else:
        info += f"after {var.it:3} {var.cycle}-cycles   "
        info += f"[{l2_last:.3e}, {l2_last/l2_prev:.3f}]"
    info += f"   {var.lr_dir} {var.sc_dir}"
    if var.verb > 4:
        info += "\n"
    var.cprint(info, 3)


def _print_gs_info(var, it, level, cycmax, grid, norm, add):
    """Log line after a smoothing step, verb > 4 (solver.py:1865-1892)."""
    n = grid.shape_cells
    info = f"     {it:2} {level} {cycmax} [{n[0]:3}, {n[1]:3}, {n[2]:3}]: {norm:.3e} "
    var.cprint(info + add, 4)


def _print_one_liner(var, l2_last, last=False):
    """Continuously updated one-line status (solver.py:1895-1919)."""
    info = f":: emg3d :: {l2_last/var.l2_refe:.1e}; "
    if var.sslsolver:
        info += f"{var.ssl_it}({var.it}); "
    else:
        info += f"{var.it}; "
    info += f"{var.time.runtime}"
    if last:
        var.cprint(info + f"; {var.exit_message}", -100)
    else:
        var.cprint(info, -100, end='\r')

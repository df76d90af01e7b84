"""Independent solves spread over the GPUs of a node -- the reference's own parallel
mode (emg3d/_multiprocessing.py:33-65 ``process_map`` with ``max_workers``, used by
``Simulation.compute`` for its (source, frequency) pairs, simulations.py:835-880).

One worker process per GPU (``spawn``; the library binds one device per process). A
worker receives the model once, keeps its coefficients, grid hierarchy and cached
factorisations on its GPU (:class:`emg3d_b200.solver.Workspace`) and solves the source
fields it is handed one after the other; fields travel by pickling, like in the
reference.  This is the simple multi-GPU mode; one solve on several GPUs is
:mod:`emg3d_b200.parallel`.
"""
import multiprocessing as mp
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

__all__ = ['process_map', 'solve', 'solve_many', 'gradient']

_STATE = {}


def _bind_device(queue, payload):
    """Worker initialiser: take one device id, remember the shared payload."""
    dev = queue.get()
    os.environ['EMG3D_B200_DEVICE'] = str(dev)
    _STATE.clear()
    _STATE['device'] = dev
    _STATE['payload'] = payload


def process_map(fn, *iterables, devices, payload=None):
    """``list(map(fn, *iterables))`` on one worker process per entry of ``devices``.

    Every worker is bound to its device before the first task (environment variable
    ``EMG3D_B200_DEVICE``, read by :func:`emg3d_b200._lib.init`); ``payload`` is sent
    to every worker once and is available to ``fn`` as ``batch.worker_payload()``.
    ``fn`` must be importable in the workers (a module-level function).
    """
    devices = list(devices)
    if not devices:
        raise ValueError("`devices` must name at least one GPU")
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    for d in devices:
        queue.put(d)
    with ProcessPoolExecutor(max_workers=len(devices), mp_context=ctx,
                             initializer=_bind_device, initargs=(queue, payload)) as ex:
        return list(ex.map(fn, *iterables))


def worker_payload():
    return _STATE.get('payload')


def worker_device():
    return _STATE.get('device')


def _echo_device(x):
    """Test helper: which device the worker that got task ``x`` is bound to."""
    return x, worker_device(), os.environ.get('EMG3D_B200_DEVICE')


def _solve_one(task):
    """Runs in a worker: one source field through solve() on the worker's GPU."""
    import emg3d_b200 as eb
    field, frequency, shm_name = task
    model, kwargs = worker_payload()
    ws = _STATE.get('workspace')
    if ws is None:
        ws = _STATE['workspace'] = eb.Workspace()
    if isinstance(field, tuple):                         # a sparse source: (indices, values, background)
        sfield = eb.SourceField(model.grid, *field, frequency)
    else:
        sfield = eb.Field(model.grid, field, frequency=frequency)
    out = eb.solve(model, sfield, workspace=ws, **kwargs)
    if isinstance(out, tuple):
        efield, info = out
    else:
        efield, info = out, None
    if shm_name is None:
        return np.asarray(efield.field), info, worker_device()
    # the result goes into the parent's shared-memory block: one host copy instead of pickling
    # 0.8 GB (at 256^3) through a pipe
    from multiprocessing import shared_memory
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        dst = np.ndarray(efield.field.shape, dtype=efield.field.dtype, buffer=shm.buf)
        dst[:] = efield.field
        del dst
    finally:
        shm.close()
    return None, info, worker_device()


def solve_many(model, sfields, devices=None, **kwargs):
    """Solve ``A e = s`` for every source field in ``sfields`` (same model and grid).

    ``devices``: GPU ids to use (default: all visible ones); other keyword arguments go
    to :func:`emg3d_b200.solve` (``efield=`` is not supported here).  Returns a list of
    ``(efield, info)`` in the order of ``sfields`` (``info`` is None unless
    ``return_info=True``).
    """
    from emg3d_b200 import _lib, fields
    if 'efield' in kwargs:
        raise ValueError("solve_many: `efield` is not supported; fields start from zero")
    sfields = list(sfields)
    if devices is None:
        import ctypes
        n = ctypes.c_int(0)
        _lib.check(_lib.load().emg3d_b200_device_count(ctypes.byref(n)))
        devices = list(range(max(n.value, 1)))
    devices = list(devices)[:max(len(sfields), 1)]
    # (dipole / wire sources travel as their few non-zero edges; the fields come back through
    # shared memory, one block per source, owned by the returned Field)
    import weakref
    from multiprocessing import shared_memory
    from emg3d_b200 import solver
    blocks = []
    for s in sfields:
        nbytes = int(s.grid.n_edges) * solver._field_dtype(s).itemsize
        blocks.append(shared_memory.SharedMemory(create=True, size=max(nbytes, 1)))
    tasks = [(s.sparse if getattr(s, 'sparse', None) is not None else np.asarray(s.field), s._frequency, b.name)
             for s, b in zip(sfields, blocks)]
    try:
        results = process_map(_solve_one, tasks, devices=devices, payload=(model, kwargs))
    except BaseException:
        for b in blocks:
            b.close()
            b.unlink()
        raise
    out = []
    for (_, info, _), s, b in zip(results, sfields, blocks):
        arr = np.ndarray(int(s.grid.n_edges), dtype=solver._field_dtype(s), buffer=b.buf)
        f = fields.Field(s.grid, dtype=arr.dtype, frequency=s._frequency)
        f._field = arr                      # the Field works on the shared block ...
        weakref.finalize(f, _release_block, b)          # ... and releases it when it goes away
        out.append((f, info))
    return out


def _release_block(block):
    """Unlink the shared block (the name disappears; the memory lives on while an array still maps
    it) and close our mapping if nobody holds a view of it any more."""
    for call in (block.unlink, block.close):
        try:
            call()
        except Exception:       # noqa: BLE001 -- views still exported / interpreter shutdown
            pass


def solve(inp):
    """``emg3d._multiprocessing.solve`` (emg3d/_multiprocessing.py:72-153) for dict input: the task a
    worker of the reference's process pool runs.  Keys ``model, sfield, efield, solver_opts`` (->
    :func:`emg3d_b200.solve`) or ``model, grid, source, frequency, efield, solver_opts`` (->
    :func:`emg3d_b200.solve_source`).  The model may live on another grid: it is interpolated to
    the computational grid first (volume averaging on the device, ``Model.interpolate_to_grid``).
    Returns ``(efield, info)``.  (The file-based variant of the reference belongs to its I/O layer.)
    """
    import emg3d_b200 as eb
    if not isinstance(inp, dict):
        raise NotImplementedError("batch.solve: dict input only (file-based computation is out of scope)")
    opts = dict(inp.get('solver_opts') or {})
    if 'sfield' in inp:
        grid = inp['sfield'].grid
        call, extra = eb.solve, {'sfield': inp['sfield']}
    else:
        grid = inp['grid']
        call, extra = eb.solve_source, {'source': inp['source'], 'frequency': inp['frequency']}
    model = inp['model'].interpolate_to_grid(grid)
    return call(model=model, efield=inp.get('efield'), return_info=True, always_return=True, **extra, **opts)


def adjoint_source_field(grid, receivers, strength, frequency, length=1.0):
    """Residual source field (emg3d/simulations.py:1234-1268): every receiver ``(x, y, z, azimuth,
    elevation)`` becomes an electric dipole of ``length`` metres with its complex ``strength``."""
    from emg3d_b200 import fields
    rec = np.array([np.broadcast_to(np.asarray(c, dtype=float), np.shape(strength)) for c in receivers]).T
    total = {}
    for coords, st in zip(rec, np.asarray(strength)):
        if np.isnan(st):
            continue
        i, v, _ = fields.get_source_field(grid, tuple(coords), frequency, strength=st, length=length).sparse
        for ii, vv in zip(i, v):
            total[int(ii)] = total.get(int(ii), 0.0) + vv
    keys = np.array(sorted(total), dtype=np.int64)
    vals = np.array([total[k] for k in keys], dtype=np.complex128 if frequency > 0 else np.float64)
    return fields.SourceField(grid, keys, vals, vals.dtype.type(0), frequency)


def _gradient_one(task):
    """Runs in a worker: forward solve, responses, residual source, back-propagation and the
    gradient contribution of ONE source -- the forward and the back-propagated field never leave
    the GPU (fields.DeviceField, csrc/interp.cu: gradient kernel)."""
    import emg3d_b200 as eb
    from emg3d_b200 import _lib
    source, observed, weights = task
    model, frequency, receivers, length, kwargs = worker_payload()
    for key in ('workspace', 'workspace_b'):
        if _STATE.get(key) is None:
            _STATE[key] = eb.Workspace()
    ws_f, ws_b = _STATE['workspace'], _STATE['workspace_b']
    grid = model.grid
    sfield = eb.get_source_field(grid, source, frequency)
    efield, synthetic = eb.solve(model, sfield, receivers=receivers, return_field='device', workspace=ws_f,
                                 **kwargs)
    residual = synthetic - observed
    misfit = 0.5 * float(np.nansum(weights * np.abs(residual) ** 2))
    smu0 = complex(sfield.smu0)
    strength = np.conj(residual * weights / -smu0)
    rfield = adjoint_source_field(grid, receivers, strength, frequency, length)
    # (a second workspace: its field buffer must not overwrite the forward field)
    bfield = eb.solve(model, rfield, return_field='device', workspace=ws_b, **kwargs)
    d_h = [_lib.DeviceArray.from_host(np.ascontiguousarray(h, dtype=float)) for h in grid.h]
    d_g = _lib.DeviceArray(3 * grid.n_cells, float)
    _lib.check(_lib.load().emg3d_b200_gradient_field(*grid.shape_cells, efield.array.ptr, bfield.array.ptr,
                                                     smu0.real, smu0.imag, d_h[0].ptr, d_h[1].ptr, d_h[2].ptr,
                                                     d_g.ptr))
    grad = d_g.download().reshape((*grid.shape_cells, 3), order='F')
    return synthetic, misfit, np.moveaxis(grad, -1, 0), worker_device()


def gradient(model, sources, frequency, receivers, observed, weights=None, devices=None, length=1.0,
             **kwargs):
    """Misfit and its adjoint-state gradient for a set of sources at one frequency, sources fanned
    out over the GPUs: the solver-level core of ``Simulation.gradient``
    (emg3d/simulations.py:944-1095; ``_bcompute`` 1193-1233, ``_get_rfield`` 1234-1268) on ONE
    grid: ``misfit = 1/2 sum w |syn - obs|^2``; per source the residuals become the sources of
    a back-propagated field ``b`` and ``Re(b s mu0 e)`` is mapped to volume-weighted cell averages
    (maps.interp_edges_to_vol_averages).  ``observed`` (n_sources, n_receivers) complex, NaN = no
    datum; ``weights`` default 1.  Returns ``(misfit, gradient, synthetic)`` with ``gradient`` of
    shape (1 | 2 | 3, nx, ny, nz) for isotropic | HTI, VTI | triaxial models, in the space of the
    model's mapping (derivative chain applied).  Surveys, data containers and source- /
    frequency-dependent grids of the reference's Simulation are outside this path."""
    from emg3d_b200 import _lib
    for name in ('epsilon_r', 'mu_r'):
        v = getattr(model, name)
        if v is not None and not np.allclose(v, 1.0):
            what = {'epsilon_r': 'el. permittivity', 'mu_r': 'magn. permeability'}[name]
            raise NotImplementedError(f"Gradient not implemented for {what}.")
    sources = list(sources)
    observed = np.asarray(observed)
    weights = np.ones(observed.shape) if weights is None else np.broadcast_to(weights, observed.shape)
    if devices is None:
        import ctypes
        n = ctypes.c_int(0)
        _lib.check(_lib.load().emg3d_b200_device_count(ctypes.byref(n)))
        devices = list(range(max(n.value, 1)))
    devices = list(devices)[:max(len(sources), 1)]
    tasks = [(src, observed[k], weights[k]) for k, src in enumerate(sources)]
    results = process_map(_gradient_one, tasks, devices=devices,
                          payload=(model, frequency, receivers, length, kwargs))
    grad = np.zeros((3, *model.shape), order='F')
    misfit = 0.0
    for _, mf, g, _ in results:
        grad += g
        misfit += mf
    synthetic = np.array([r[0] for r in results])
    # fold the directions the model does not distinguish, apply the mapping's chain rule
    # (simulations.py:1063-1090)
    keep = [0]
    if model.case in ('HTI', 'triaxial'):
        model.map.derivative_chain(grad[1], model.property_y)
        keep.append(1)
    else:
        grad[0] += grad[1]
    if model.case in ('VTI', 'triaxial'):
        model.map.derivative_chain(grad[2], model.property_z)
        keep.append(2)
    else:
        grad[0] += grad[2]
    model.map.derivative_chain(grad[0], model.property_x)
    return misfit, grad[keep], synthetic

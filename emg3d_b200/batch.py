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

__all__ = ['process_map', 'solve_many']

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
    field, frequency = task
    model, kwargs = worker_payload()
    ws = _STATE.get('workspace')
    if ws is None:
        ws = _STATE['workspace'] = eb.Workspace()
    sfield = eb.Field(model.grid, field, frequency=frequency)
    out = eb.solve(model, sfield, workspace=ws, **kwargs)
    if isinstance(out, tuple):
        efield, info = out
    else:
        efield, info = out, None
    return np.asarray(efield.field), info, worker_device()


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
    tasks = [(np.asarray(s.field), s._frequency) for s in sfields]
    results = process_map(_solve_one, tasks, devices=devices, payload=(model, kwargs))
    out = []
    for (arr, info, _), s in zip(results, sfields):
        out.append((fields.Field(s.grid, arr, frequency=s._frequency), info))
    return out

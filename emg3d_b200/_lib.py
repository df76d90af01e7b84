"""ctypes binding of ``libemg3d_b200.so`` (C ABI: include/emg3d_b200.h).

Thin layer: loads the in-tree shared library, declares every entry point,
turns non-zero status codes into :class:`Emg3dB200Error`, and provides small
RAII holders for device memory and grid levels.  No CPU fallback exists: if the
library or a CUDA device is missing, calls raise.
"""
import ctypes
import os
import subprocess
from ctypes import (POINTER, byref, c_char_p, c_double, c_float, c_int, c_longlong,
                    c_size_t, c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, 'libemg3d_b200.so')
CSRC = os.path.join(_HERE, 'csrc')

ORDER_LEX, ORDER_COLOR = 0, 1


class Emg3dB200Error(RuntimeError):
    """Raised for any failure reported through the C ABI."""


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a with nvcc (cross-compiles without GPU)."""
    cmd = ['make', '-C', CSRC, '-j8'] + (['-B'] if force else [])
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise Emg3dB200Error("building libemg3d_b200.so failed:\n" +
                             (res.stdout or '') + (res.stderr or ''))
    if not os.path.exists(LIBPATH):
        raise Emg3dB200Error("build finished but libemg3d_b200.so is missing")


_SIGNATURES = {
    # name: (restype, argtypes)
    'emg3d_b200_abi_version': (c_int, []),
    'emg3d_b200_last_error': (c_char_p, []),
    'emg3d_b200_device_count': (c_int, [POINTER(c_int)]),
    'emg3d_b200_init': (c_int, [c_int]),
    'emg3d_b200_device_name': (c_int, [c_char_p, c_int]),
    'emg3d_b200_mem_info': (c_int, [POINTER(c_size_t), POINTER(c_size_t)]),
    'emg3d_b200_sync': (c_int, []),
    'emg3d_b200_launch_count': (c_int, [POINTER(c_longlong)]),
    'emg3d_b200_malloc': (c_int, [POINTER(c_void_p), c_size_t]),
    'emg3d_b200_free': (c_int, [c_void_p]),
    'emg3d_b200_malloc_scratch': (c_int, [POINTER(c_void_p), c_size_t]),
    'emg3d_b200_free_scratch': (c_int, [c_void_p]),
    'emg3d_b200_memset': (c_int, [c_void_p, c_int, c_size_t]),
    'emg3d_b200_h2d': (c_int, [c_void_p, c_void_p, c_size_t]),
    'emg3d_b200_d2h': (c_int, [c_void_p, c_void_p, c_size_t]),
    'emg3d_b200_d2d': (c_int, [c_void_p, c_void_p, c_size_t]),
    'emg3d_b200_host_alloc': (c_int, [POINTER(c_void_p), c_size_t]),
    'emg3d_b200_host_free': (c_int, [c_void_p]),
    'emg3d_b200_h2d_sparse': (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_int)]),
    'emg3d_b200_event_create': (c_int, [POINTER(c_void_p)]),
    'emg3d_b200_event_record': (c_int, [c_void_p]),
    'emg3d_b200_event_elapsed_ms': (c_int, [c_void_p, c_void_p, POINTER(c_float)]),
    'emg3d_b200_event_destroy': (c_int, [c_void_p]),
    'emg3d_b200_graph_begin': (c_int, []),
    'emg3d_b200_graph_end': (c_int, [POINTER(c_void_p)]),
    'emg3d_b200_graph_launch': (c_int, [c_void_p]),
    'emg3d_b200_graph_destroy': (c_int, [c_void_p]),
    'emg3d_b200_level_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_level_destroy': (c_int, [c_void_p]),
    'emg3d_b200_level_set_model': (c_int, [c_void_p, c_int, c_void_p, c_void_p,
                                           c_void_p, c_void_p]),
    'emg3d_b200_level_window': (c_int, [POINTER(c_void_p), c_void_p, c_int, c_int]),
    'emg3d_b200_level_set_owned': (c_int, [c_void_p, c_int, c_int]),
    'emg3d_b200_level_set_zflip': (c_int, [c_void_p, c_int]),
    'emg3d_b200_level_line_chain': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_point_schedule_kind': (c_int, [c_void_p, POINTER(c_int)]),
    'emg3d_b200_level_factor_bytes': (c_int, [c_void_p, c_int, POINTER(c_size_t)]),
    'emg3d_b200_level_drop_factors': (c_int, [c_void_p]),
    'emg3d_b200_level_link': (c_int, [c_void_p, c_void_p, POINTER(c_int),
                                      POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_void_p)]),
    'emg3d_b200_amat_x': (c_int, [c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_apply': (c_int, [c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_point_tile_schedule': (c_int, [POINTER(c_int)]),
    'emg3d_b200_point_tile_shape': (c_int, [POINTER(c_int)]),
    'emg3d_b200_line_seg_mask': (c_int, [c_int, POINTER(c_int)]),
    'emg3d_b200_magnetic_field': (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double]),
    'emg3d_b200_host_edge_curl_factor': (c_int, [c_int, c_int, c_int, c_int] + [c_void_p] * 10),
    'emg3d_b200_residual': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_residual_norm': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                                         POINTER(c_double)]),
    'emg3d_b200_gauss_seidel': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
    'emg3d_b200_restrict': (c_int, [c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_prolong': (c_int, [c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_restrict_cells': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    'emg3d_b200_volume_model': (c_int, [c_void_p, c_int, c_double, c_double, c_double, c_double,
                                        c_int] + [c_void_p] * 9),
    'emg3d_b200_pec_zero': (c_int, [c_void_p, c_void_p]),
    'emg3d_b200_dot': (c_int, [c_int, c_longlong, c_void_p, c_void_p, c_int, c_void_p]),
    'emg3d_b200_dot_host': (c_int, [c_int, c_longlong, c_void_p, c_void_p, c_int,
                                    POINTER(c_double)]),
    'emg3d_b200_axpby': (c_int, [c_int, c_longlong, c_double, c_double, c_void_p,
                                 c_double, c_double, c_void_p]),
    'emg3d_b200_host_amat_x': (c_int, [c_int, c_int, c_int, c_int] + [c_void_p] * 13),
    'emg3d_b200_fill_scatter': (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p, c_size_t]),
    'emg3d_b200_volume_average': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    'emg3d_b200_edges_to_vol_averages': (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p]),
    'emg3d_b200_gradient_field': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_double,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_spline_filter3': (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int]),
    'emg3d_b200_copy_box3': (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'emg3d_b200_pad_edge3': (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'emg3d_b200_interp_points': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                         c_double, c_double, c_void_p, c_void_p, c_void_p,
                                         ctypes.c_longlong, c_int, c_int, c_int, c_double, c_double,
                                         c_int, c_void_p]),
    'emg3d_b200_host_gauss_seidel': (c_int, [c_int] * 6 + [c_void_p] * 13 + [c_int]),
    'emg3d_b200_host_solve': (c_int, [c_int, c_int, c_void_p, c_void_p]),
    'emg3d_b200_host_restrict': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
    'emg3d_b200_comm_unique_id': (c_int, [c_void_p]),
    'emg3d_b200_comm_init': (c_int, [c_void_p, c_int, c_int]),
    'emg3d_b200_comm_size': (c_int, [POINTER(c_int), POINTER(c_int)]),
    'emg3d_b200_comm_destroy': (c_int, []),
    'emg3d_b200_comm_sendrecv': (c_int, [c_int, POINTER(c_void_p), POINTER(c_size_t),
                                         POINTER(c_int), POINTER(c_int)]),
    'emg3d_b200_comm_allreduce_sum': (c_int, [c_void_p, c_int]),
    'emg3d_b200_p2p_init': (c_int, [POINTER(c_int)]),
    'emg3d_b200_p2p_register': (c_int, [c_void_p, POINTER(c_int)]),
    'emg3d_b200_p2p_release': (c_int, []),
    'emg3d_b200_p2p_exchange': (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'emg3d_b200_p2p_status': (c_int, [POINTER(c_int)]),
    'emg3d_b200_p2p_shutdown': (c_int, []),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None
_initialised = False


def load():
    """Load the shared library (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise Emg3dB200Error(
                f"{LIBPATH} not found: build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU "
                "fallback.")
        lib = ctypes.CDLL(LIBPATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        msg = load().emg3d_b200_last_error()
        raise Emg3dB200Error(f"[{status}] {msg.decode() if msg else 'unknown error'}")


def init(device=None):
    """Select the device (default: LOCAL_RANK or 0) and create the stream."""
    global _initialised
    lib = load()
    if not _initialised:
        if device is None:
            device = int(os.environ.get('EMG3D_B200_DEVICE',
                                        os.environ.get('LOCAL_RANK', '0')))
        n = c_int(0)
        check(lib.emg3d_b200_device_count(byref(n)))
        if n.value < 1:
            raise Emg3dB200Error("no CUDA device visible; emg3d_b200 has no CPU path")
        check(lib.emg3d_b200_init(int(device) % n.value))
        _initialised = True
    return lib


def device_name():
    buf = ctypes.create_string_buffer(256)
    check(init().emg3d_b200_device_name(buf, 256))
    return buf.value.decode()


def mem_info():
    f, t = c_size_t(0), c_size_t(0)
    check(init().emg3d_b200_mem_info(byref(f), byref(t)))
    return f.value, t.value


def sync():
    check(init().emg3d_b200_sync())


def launch_count():
    n = c_longlong(0)
    check(load().emg3d_b200_launch_count(byref(n)))
    return n.value


def line_seg_mask(mask=-1):
    """Set (mask >= 0) / query which line directions use the segment-parallel kernels
    (bit a = axis a; see emg3d_b200_line_seg_mask in the C header).  Returns the previous mask."""
    prev = ctypes.c_int(0)
    check(load().emg3d_b200_line_seg_mask(int(mask), byref(prev)))
    return prev.value


def _hptr(a):
    return a.ctypes.data_as(c_void_p)


class DeviceArray:
    """Owned device buffer with a NumPy-like dtype/size."""

    def __init__(self, size, dtype, scratch=False):
        """scratch: a short-lived work array from the stream-ordered pool (emg3d_b200_malloc_scratch)."""
        self.dtype = np.dtype(dtype)
        self.size = int(size)
        self.nbytes = self.size * self.dtype.itemsize
        self.scratch = bool(scratch)
        p = c_void_p(0)
        if self.scratch:
            check(init().emg3d_b200_malloc_scratch(byref(p), self.nbytes))
        else:
            check(init().emg3d_b200_malloc(byref(p), self.nbytes))
        self.ptr = p.value

    @classmethod
    def from_host(cls, arr, scratch=False):
        arr = np.ascontiguousarray(arr.ravel('F') if arr.ndim > 1 else arr)
        self = cls(arr.size, arr.dtype, scratch)
        self.upload(arr)
        return self

    def upload(self, arr):
        arr = np.ascontiguousarray(arr.ravel('F') if arr.ndim > 1 else arr, dtype=self.dtype)
        if arr.size != self.size:
            raise ValueError(f"size mismatch: {arr.size} != {self.size}")
        check(init().emg3d_b200_h2d(self.ptr, _hptr(arr), self.nbytes))

    def upload_sparse(self, arr):
        """Upload of a mostly-zero array (source fields): see emg3d_b200_h2d_sparse.
        Returns True if only the non-zeros crossed PCIe."""
        arr = np.ascontiguousarray(arr.ravel('F') if arr.ndim > 1 else arr, dtype=self.dtype)
        if arr.size != self.size:
            raise ValueError(f"size mismatch: {arr.size} != {self.size}")
        used = c_int(0)
        check(init().emg3d_b200_h2d_sparse(self.ptr, _hptr(arr), self.size, self.dtype.itemsize,
                                           byref(used)))
        return bool(used.value)

    def fill_scatter(self, fill, idx, val):
        """self[:] = fill; self[idx] = val (a sparse source field: emg3d_b200_fill_scatter)."""
        fill = np.array([fill], dtype=self.dtype)
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=self.dtype)
        check(init().emg3d_b200_fill_scatter(self.ptr, self.size, self.dtype.itemsize, _hptr(fill),
                                             _hptr(idx), _hptr(val), idx.size))

    def upload_ptr(self, hptr, nbytes=None):
        check(init().emg3d_b200_h2d(self.ptr, hptr, self.nbytes if nbytes is None else nbytes))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.size, dtype=self.dtype)
        check(init().emg3d_b200_d2h(_hptr(out), self.ptr, self.nbytes))
        return out

    def download_ptr(self, hptr, nbytes=None):
        check(init().emg3d_b200_d2h(hptr, self.ptr, self.nbytes if nbytes is None else nbytes))

    def zero(self):
        check(init().emg3d_b200_memset(self.ptr, 0, self.nbytes))

    def copy_from(self, other):
        check(init().emg3d_b200_d2d(self.ptr, other.ptr, self.nbytes))

    def copy(self):
        new = DeviceArray(self.size, self.dtype)
        new.copy_from(self)
        return new

    def free(self):
        if getattr(self, 'ptr', None):
            try:
                if getattr(self, 'scratch', False):
                    load().emg3d_b200_free_scratch(self.ptr)
                else:
                    load().emg3d_b200_free(self.ptr)
            except Exception:
                pass
            self.ptr = None

    def __del__(self):
        self.free()


def _host_free(ptr):
    try:
        load().emg3d_b200_host_free(ptr)
    except Exception:
        pass


class PinnedArray:
    """Page-locked host buffer exposed as a NumPy array.

    The memory belongs to the array: it is released when the LAST NumPy view of it is gone (a
    finalizer on the ctypes object every view keeps alive through its base chain), so a Field
    returned by ``solve(..., workspace=Workspace(pinned_result=True))`` stays valid after the
    workspace dropped or replaced the buffer."""

    def __init__(self, size, dtype):
        import weakref
        self.dtype = np.dtype(dtype)
        self.size = int(size)
        self.nbytes = self.size * self.dtype.itemsize
        p = c_void_p(0)
        check(init().emg3d_b200_host_alloc(byref(p), self.nbytes))
        self.ptr = p.value
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        weakref.finalize(buf, _host_free, self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=self.size)

    def free(self):
        """Drop this object's reference; views handed out keep the memory alive."""
        self.array = None
        self.ptr = None


class Event:
    def __init__(self):
        p = c_void_p(0)
        check(init().emg3d_b200_event_create(byref(p)))
        self.ptr = p.value

    def record(self):
        check(load().emg3d_b200_event_record(self.ptr))
        return self

    def elapsed_ms(self, stop):
        ms = c_float(0)
        check(load().emg3d_b200_event_elapsed_ms(self.ptr, stop.ptr, byref(ms)))
        return ms.value

    def __del__(self):
        if getattr(self, 'ptr', None):
            try:
                load().emg3d_b200_event_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = None


class Graph:
    """CUDA graph captured from the library stream."""

    def __init__(self):
        self.ptr = None

    def __enter__(self):
        check(init().emg3d_b200_graph_begin())
        return self

    def __exit__(self, et, ev, tb):
        p = c_void_p(0)
        status = load().emg3d_b200_graph_end(byref(p))
        if et is None:
            check(status)
            self.ptr = p.value
        return False

    def launch(self):
        check(load().emg3d_b200_graph_launch(self.ptr))

    def __del__(self):
        if getattr(self, 'ptr', None):
            try:
                load().emg3d_b200_graph_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = None


class LevelHandle:
    """Owner of one ``emg3d_b200_level``."""

    def __init__(self, h):
        hx, hy, hz = (np.ascontiguousarray(a, dtype=np.float64) for a in h)
        p = c_void_p(0)
        check(init().emg3d_b200_level_create(byref(p), hx.size, hy.size, hz.size,
                                             _hptr(hx), _hptr(hy), _hptr(hz)))
        self.ptr = p.value
        self.shape = (hx.size, hy.size, hz.size)

    def set_model(self, cplx, eta_x, eta_y, eta_z, zeta):
        check(load().emg3d_b200_level_set_model(self.ptr, int(cplx), eta_x.ptr, eta_y.ptr,
                                                eta_z.ptr, zeta.ptr))
        self._keep = (eta_x, eta_y, eta_z, zeta)   # keep the buffers alive

    def link(self, fine, cflag, weights, lo, frac):
        """weights: list of 9 arrays or None; lo/frac: 3 arrays each."""
        cf = (c_int * 3)(*[int(c) for c in cflag])
        keep = []
        wp = (c_void_p * 9)()
        for k in range(9):
            if weights[k] is None:
                wp[k] = None
            else:
                a = np.ascontiguousarray(weights[k], dtype=np.float64)
                keep.append(a)
                wp[k] = a.ctypes.data
        lp = (c_void_p * 3)()
        fp = (c_void_p * 3)()
        for a in range(3):
            li = np.ascontiguousarray(lo[a], dtype=np.int32)
            fr = np.ascontiguousarray(frac[a], dtype=np.float64)
            keep += [li, fr]
            lp[a], fp[a] = li.ctypes.data, fr.ctypes.data
        check(load().emg3d_b200_level_link(self.ptr, fine.ptr, cf, wp, lp, fp))

    def window(self, z0, nz):
        """View on cells [z0, z0 + nz) along z (model must be set); see the header."""
        p = c_void_p(0)
        check(load().emg3d_b200_level_window(byref(p), self.ptr, int(z0), int(nz)))
        w = LevelHandle.__new__(LevelHandle)
        w.ptr, w.shape = p.value, (self.shape[0], self.shape[1], int(nz))
        w._keep = self              # the parent owns the arrays the window points into
        return w

    def factor_bytes(self, ldir):
        n = c_size_t(0)
        check(load().emg3d_b200_level_factor_bytes(self.ptr, int(ldir), byref(n)))
        return n.value

    def drop_factors(self):
        check(load().emg3d_b200_level_drop_factors(self.ptr))

    def free(self):
        if getattr(self, 'ptr', None):
            try:
                load().emg3d_b200_level_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = None

    def __del__(self):
        self.free()

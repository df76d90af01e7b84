"""Where the time of one end-to-end solve() call goes (host wall clock, synchronised).

    python tools/e2e_profile.py [n]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, recipes, solver  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = recipes.config('config3', n)
grid = eb.TensorMesh(cfg['h'], cfg['origin'])
model = eb.Model(grid, **cfg['model'])
sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
_lib.init()
pin = _lib.PinnedArray(sfield.field.size, sfield.field.dtype)
pin.array[:] = sfield.field
h_s = eb.Field(grid, pin.array, frequency=cfg['frequency'])
ws = eb.Workspace(pinned_result=True)
kw = dict(plain=True, cycle='V', maxit=1, verb=-1, workspace=ws)
for _ in range(3):
    eb.solve(model, h_s, **kw)
_lib.sync()


def timed(label, fn, reps=5):
    _lib.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    _lib.sync()
    print(f"{label:34s} {1e3 * (time.perf_counter() - t0) / reps:8.2f} ms")
    return out


timed("solve() end to end", lambda: eb.solve(model, h_s, **kw))
lv = ws.level(model, h_s)
d_s = ws.device_buffer('s', lv.n_edges, h_s.field.dtype)
d_e = ws.device_buffer('e', lv.n_edges, h_s.field.dtype)
out = ws.result_buffer(lv.n_edges, h_s.field.dtype)
timed("  workspace.level (fingerprint)", lambda: ws.level(model, h_s))
timed("  upload source (pinned)", lambda: d_s.upload(np.asarray(h_s.field)))
timed("  upload source (pageable)", lambda: d_s.upload(np.asarray(sfield.field)))
timed("  upload_sparse (pageable)", lambda: d_s.upload_sparse(np.asarray(sfield.field)))
timed("  solve(), pageable source", lambda: eb.solve(model, sfield, **kw))
d_s.upload(np.asarray(h_s.field))
nrm = timed("  norm of source", lambda: solver._Vec(lv.cplx, lv.n_edges).norm(d_s))
timed("  zero field", d_e.zero)


def cyc():
    var = solver.MGParameters(verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
                              shape_cells=grid.shape_cells, cycle='V', maxit=1)
    var.order = None
    var.l2_refe = nrm
    d_e.zero()
    var.e_is_zero, var.s_norm = True, nrm
    solver._multigrid(lv, d_s, d_e, var)


timed("  V-cycle (device)", cyc)
timed("  MGParameters()", lambda: solver.MGParameters(
    verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
    shape_cells=grid.shape_cells, cycle='V', maxit=1))
timed("  download field (pinned)", lambda: d_e.download(out=out))
timed("  Field() around result", lambda: eb.Field(grid, out, frequency=1.0))

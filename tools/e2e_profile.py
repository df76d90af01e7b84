"""Host-side profile (cProfile) of the survey-style end-to-end call of bench.py at 256^3.

    python tools/e2e_profile.py [n]
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = recipes.config('config3', n)
grid = eb.TensorMesh(cfg['h'], cfg['origin'])
model = eb.Model(grid, **cfg['model'])
for k in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r'):
    if getattr(model, k) is not None:
        getattr(model, k).flags.writeable = False
src, freq = cfg['source'], cfg['frequency']
rec = (np.linspace(0.5 * (grid.nodes_x[0] + src[0]), 0.5 * (grid.nodes_x[-1] + src[0]), 101),
       float(src[1]), float(src[2]), 0.0, 0.0)
ws = eb.Workspace(pinned_result=True)
kw = dict(plain=True, cycle='V', maxit=1, verb=-1, workspace=ws, receivers=rec, return_field=False)
for _ in range(3):
    eb.solve(model, eb.get_source_field(grid, src, freq), **kw)
_lib.sync()
t0 = time.perf_counter()
for _ in range(5):
    eb.solve(model, eb.get_source_field(grid, src, freq), **kw)
_lib.sync()
print(f"survey-style step: {1e3 * (time.perf_counter() - t0) / 5:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    eb.solve(model, eb.get_source_field(grid, src, freq), **kw)
_lib.sync()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

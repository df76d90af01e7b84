"""cProfile of the first (cold) solve() call: model upload, device VolumeModel, hierarchy.

    python tools/cold_profile.py [n]
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = recipes.config('config3', n)
grid = eb.TensorMesh(cfg['h'], cfg['origin'])
model = eb.Model(grid, **cfg['model'])
sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
_lib.init()
small = recipes.config('config3', 16)                        # warm the library itself
g16 = eb.TensorMesh(small['h'], small['origin'])
eb.solve(eb.Model(g16, **small['model']), eb.get_source_field(g16, small['source'], 1.0),
         plain=True, cycle='V', maxit=1, verb=-1)
_lib.sync()
ws = eb.Workspace(pinned_result=True)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
eb.solve(model, sfield, plain=True, cycle='V', maxit=1, verb=-1, workspace=ws)
_lib.sync()
pr.disable()
print(f"cold solve(): {1e3 * (time.perf_counter() - t0):.1f} ms")
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

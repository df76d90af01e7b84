"""One solve of a named configuration (for ncu launch lists): python tools/one_solve.py config2:128 [maxit]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import recipes  # noqa: E402

name, n = sys.argv[1].split(':')
cfg = recipes.config(name, int(n))
grid = eb.TensorMesh(cfg['h'], cfg['origin'])
model = eb.Model(grid, **cfg['model'])
sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
kw = dict(cfg['solver'])
if len(sys.argv) > 2:
    kw['maxit'] = int(sys.argv[2])
e, info = eb.solve(model, sfield, return_info=True, **kw)
print(info['exit_message'], info['it_mg'], info['it_ssl'], info['rel_error'])

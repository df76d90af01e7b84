"""Run each hot kernel a few times on the finest grid of the bench workload.
Used under ncu (see profiles/README.md); prints nothing that is a bench value."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb
from emg3d_b200 import _lib, solver, recipes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else 'all'
cfg = recipes.config('config3', n)
grid = eb.TensorMesh(cfg['h'], cfg['origin'])
model = eb.Model(grid, **cfg['model'])
sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
vm = eb.VolumeModel(model, sfield)
lv = solver._Level.from_volume_model(vm, sfield.field.dtype)
d_s = _lib.DeviceArray.from_host(sfield.field)
d_e = lv.new_field()
lib = _lib.load()
order = _lib.ORDER_COLOR
if which in ('all', 'point'):
    _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, d_e.ptr, d_s.ptr, 1, 0, order))
if which in ('all', 'line'):
    for ldir in (1, 2, 3):
        _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, d_e.ptr, d_s.ptr, 1, ldir, order))
        lv.handle.drop_factors()
if which in ('all', 'residual'):
    r = lv.res_buffer()
    _lib.check(lib.emg3d_b200_residual(lv.handle.ptr, d_s.ptr, d_e.ptr, r.ptr, None))
    c = lv.coarse(0)
    _lib.check(lib.emg3d_b200_restrict(c.handle.ptr, r.ptr, c.s.ptr))
    _lib.check(lib.emg3d_b200_prolong(c.handle.ptr, d_e.ptr, c.e.ptr))
_lib.sync()
print("done")

"""Time the line smoothers (nu = 2 calls, CUDA events) on the finest grid of the bench workload,
one-thread-per-line kernels (mask 0) against the segment-parallel kernels (mask 7), and check that
both give the same field.  usage: time_lines.py [n=256] [masks=0,7] [ldirs=1,2,3]"""
import sys
import os
import json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb
from emg3d_b200 import _lib, solver, recipes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
masks = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else '0,7').split(',')]
ldirs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else '1,2,3').split(',')]
splits = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else '100').split(',')]
cfg = recipes.config('config3', n)
grid = eb.TensorMesh(cfg['h'], cfg['origin'])
model = eb.Model(grid, **cfg['model'])
sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
vm = eb.VolumeModel(model, sfield)
lv = solver._Level.from_volume_model(vm, sfield.field.dtype)
rng = np.random.default_rng(1)
s_host = np.asarray(sfield.field).copy()
d_s = _lib.DeviceArray.from_host(s_host)
lib = _lib.load()
order = _lib.ORDER_COLOR
cells = int(np.prod(grid.shape_cells))
peak = 6548.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
out = {}
fields = {}
for ldir in ldirs:
    for mask, split in [(m, sp) for m in masks for sp in (splits if m else [100])]:
        _lib.line_seg_mask(mask)
        lv.handle.drop_factors()
        d_e = lv.new_field()
        a, b = _lib.Event(), _lib.Event()
        a.record()
        _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, d_e.ptr, d_s.ptr, 2, ldir, order))
        b.record()
        t_first = a.elapsed_ms(b)
        fields[(ldir, mask, split)] = d_e.download()
        reps = 5
        a, b = _lib.Event(), _lib.Event()
        a.record()
        for _ in range(reps):
            _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, d_e.ptr, d_s.ptr, 2, ldir, order))
        b.record()
        t = a.elapsed_ms(b) / reps
        gbs = 184 * cells * 2 / (t * 1e-3) / 1e9
        out[f'ldir{ldir}_mask{mask}_split{split}'] = dict(ms_nu2=round(t, 4), first_call_ms=round(t_first, 3),
                                             GBs=round(gbs, 1), frac=round(gbs / peak, 4))
        print(f"n={n} ldir={ldir} mask={mask} split={split}: nu=2 call {t:.3f} ms (first call incl. factorisation "
              f"{t_first:.2f} ms), {gbs:.0f} GB/s algorithmic = {gbs / peak:.3f} of peak", flush=True)
        del d_e
    keys = [k for k in fields if k[0] == ldir]
    for k in keys[1:]:
        f0, f1 = fields[keys[0]], fields[k]
        print(f"   ldir={ldir}: |{k[1:]} - {keys[0][1:]}| / |.| = "
              f"{np.linalg.norm(f1 - f0) / np.linalg.norm(f0):.2e}", flush=True)
    lv.handle.drop_factors()
print(json.dumps(out))

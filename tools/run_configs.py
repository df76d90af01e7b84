"""Run the BASELINE.json configurations through emg3d_b200.solve and report
iterations, residuals and wall time (not a bench value; used for DESIGN.md).

    python tools/run_configs.py config2:128 config3:256 [--order color|lex] [--tol 1e-6]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, recipes  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    order = 'color'
    tol = None
    for a in sys.argv[1:]:
        if a.startswith('--order='):
            order = a.split('=')[1]
        if a.startswith('--tol='):
            tol = float(a.split('=')[1])
    for spec in args:
        name, n = spec.split(':')
        cfg = recipes.config(name, int(n))
        grid = eb.TensorMesh(cfg['h'], cfg['origin'])
        model = eb.Model(grid, **cfg['model'])
        sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
        kw = dict(cfg['solver'])
        if tol is not None:
            kw['tol'] = tol
        _lib.sync()
        free0, total = _lib.mem_info()
        t0 = time.perf_counter()
        ws = eb.Workspace()
        efield, info = eb.solve(model, sfield, return_info=True, order=order, workspace=ws, **kw)
        _lib.sync()
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        efield, info2 = eb.solve(model, sfield, return_info=True, order=order, workspace=ws, **kw)
        _lib.sync()
        dt2 = time.perf_counter() - t0
        free1, _ = _lib.mem_info()
        print(json.dumps({
            'config': name, 'shape': grid.shape_cells, 'order': order, 'solver': kw,
            'exit_message': info['exit_message'], 'it_mg': info['it_mg'], 'it_ssl': info['it_ssl'],
            'rel_error': info['rel_error'], 'wall_s': round(dt, 3), 'wall_s_warm': round(dt2, 3),
            'cycle_s': [round(float(b - a), 4) for a, b in zip([0] + list(info['runtime_at_cycle'][:-1]), info['runtime_at_cycle'])],
            'cycle_s_warm': [round(float(b - a), 4) for a, b in zip([0] + list(info2['runtime_at_cycle'][:-1]), info2['runtime_at_cycle'])],
            'error_at_cycle_rel': [float(f"{v:.3e}") for v in info['error_at_cycle'] / info['ref_error']],
            'efield_norm': float(np.linalg.norm(efield.field)),
            'device_free_GB_before': round(free0 / 1e9, 1)}), flush=True)


if __name__ == '__main__':
    main()

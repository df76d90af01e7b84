"""Line-smoother throughput against the number of lines in flight (same line length).

    python tools/line_scaling.py

Times gauss_seidel_x/_y/_z (nu = 2, colour order) on grids whose line length is 256
cells and whose number of lines grows; prints G cell-sweeps/s.  One thread relaxes one
line, so the lines per colour launch are the threads in flight.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, recipes, solver  # noqa: E402

lib = _lib.init()
for ldir, name in ((1, 'x'), (2, 'y'), (3, 'z')):
    for t in (256, 364, 512):
        shape = [t, t, t]
        shape[ldir - 1] = 256
        h, origin = recipes.grid_arrays(*shape, 1.03, 1.03, 1.03)
        grid = eb.TensorMesh(h, origin)
        model = eb.Model(grid, **recipes.model_marine(h, origin))
        sfield = eb.get_source_field(grid, (0., 0., -950., 0., 0.), 1.0)
        lv = solver._Level.from_model(model, sfield)
        d_s = _lib.DeviceArray.from_host(sfield.field)
        d_e = lv.new_field()
        for _ in range(2):
            _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, d_e.ptr, d_s.ptr, 2, ldir, 1))
        a, b = _lib.Event(), _lib.Event()
        a.record()
        for _ in range(3):
            _lib.check(lib.emg3d_b200_gauss_seidel(lv.handle.ptr, d_e.ptr, d_s.ptr, 2, ldir, 1))
        b.record()
        _lib.sync()
        ms = a.elapsed_ms(b) / 3
        cells = int(np.prod(shape))
        lines = (shape[ldir % 3] - 1) * (shape[(ldir + 1) % 3] - 1) // 4
        print(f"{name}-lines shape {shape}: {lines} lines per colour, {ms:.2f} ms (nu=2), "
              f"{2 * cells / ms / 1e6:.2f} G cell-sweeps/s", flush=True)
        del lv, d_s, d_e

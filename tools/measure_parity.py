"""Print the measured parity numbers behind the tolerances of tests/test_gpu_*.py: every golden
kernel case and golden solve, CUDA (order='lex') against the reference outputs."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import emg3d_b200 as eb
from emg3d_b200 import core, _lib
from helpers import kernel_case, solve_case, split_field

_lib.init()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))
gk = np.load(os.path.join(ROOT, 'tests/golden/kernels.npz'), allow_pickle=True)
GS = ['gauss_seidel', 'gauss_seidel_x', 'gauss_seidel_y', 'gauss_seidel_z']
ncase = int(gk['n_cases']) if 'n_cases' in gk.files else len({k.split('_')[0] for k in gk.files if k.startswith('k')})
worst = [0.0] * 4
for k in range(ncase):
    c = kernel_case(gk, k)
    margs = (c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'], c['hx'], c['hy'], c['hz'])
    for ldir in range(4):
        for nu in (1, 2):
            key = c['prefix'] + f'gs{ldir}_nu{nu}'
            if key not in gk.files:
                continue
            e = c['e'].copy()
            getattr(core, GS[ldir])(*split_field(c['shape'], e), *split_field(c['shape'], c['s']), *margs, nu, order='lex')
            worst[ldir] = max(worst[ldir], rel(e, gk[key]))
print("kernels (lex) worst rel. error, point / x / y / z:", ["%.1e" % w for w in worst])
gs = np.load(os.path.join(ROOT, 'tests/golden/solves.npz'), allow_pickle=True)
for prefix in ['res_F_', 'res_W_', 'res_V_', 'res_bic_', 'reg2_', 'lap_F_', 'lap_bic_',
               'config1_', 'config2_', 'config3_', 'config4_', 'config5_']:
    c = solve_case(gs, prefix)
    grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], c['origin'])
    model = eb.Model(grid, **c['model'])
    sfield = eb.Field(grid, c['sfield'].copy(), frequency=c['frequency'])
    e, info = eb.solve(model, sfield, return_info=True, order='lex', **dict(c['kwargs'], verb=-1))
    n = min(len(info['error_at_cycle']), len(c['error_at_cycle']))
    dcyc = np.abs(info['error_at_cycle'][:n] - c['error_at_cycle'][:n]).max() / c['ref_error']
    print(f"{prefix:10s} efield {rel(e.field, c['efield']):.1e}  max |err_at_cycle diff| / ||b|| {dcyc:.1e}  "
          f"it {info['it_mg']}/{c['it_mg']} ssl {info['it_ssl']}/{c['it_ssl']}")

"""GCROT(m,k) and CGS around the multigrid cycle against the reference (tests/golden/gcrot.npz,
cgs.npz).

Two GCROT drivers of the same algorithm: the device-resident restatement (the default; its driver
code is pinned against SciPy on a NumPy backend in tests/test_krylov_cpu.py) and SciPy on the host
with the GPU as operator and preconditioner (``EMG3D_B200_GCROT=host``).  ``order='lex'`` reproduces
the reference's Gauss-Seidel sweeps, so both must give the reference's iteration counts, exit
message and field:

* ``it_ssl``, ``it_mg``, ``exit_message`` identical;
* ``efield``: ``||e - e_ref|| / ||e_ref|| <= 1e-8`` (two outer iterations with modified
  Gram-Schmidt; the reductions are summed in another order than BLAS does);
* ``error_at_cycle`` within ``1e-8 ||b||``.

The reference only converges with GCROT when the source has a norm around one (its preconditioner
measures divergence against the norm of the original source, GCROT hands it unit vectors); the
as-is 'res' source reports DIVERGED in both, which test_gpu_solver.py covers.

The GCROT cases ran on a B200 in round 2 (6 passed); this file sorts last among the gpu tests on
purpose: the CGS cases were added after the round's GPU budget was spent.
"""
import numpy as np
import pytest

from conftest import rel_err
from helpers import solve_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eb():
    import emg3d_b200
    from emg3d_b200 import _lib
    _lib.init()
    return emg3d_b200


@pytest.mark.parametrize('driver', ['host', 'device'])
@pytest.mark.parametrize('prefix', ['res_gcrot_', 'res_gcrot_noprec_', 'config2_gcrot_'])
def test_gcrotmk_matches_reference(eb, golden, prefix, driver, monkeypatch, capsys):
    monkeypatch.setenv('EMG3D_B200_GCROT', driver)
    c = solve_case(golden('gcrot'), prefix)
    grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], c['origin'])
    model = eb.Model(grid, **c['model'])
    sfield = eb.Field(grid, c['sfield'].copy(), frequency=c['frequency'])
    efield, info = eb.solve(model, sfield, return_info=True, order='lex', **c['kwargs'])
    capsys.readouterr()
    assert info['exit_message'] == c['exit_message']
    assert (info['it_ssl'], info['it_mg']) == (c['it_ssl'], c['it_mg'])
    assert np.abs(info['error_at_cycle'] - c['error_at_cycle']).max() <= 1e-8 * c['ref_error']
    assert rel_err(efield.field, c['efield']) <= 1e-8


@pytest.mark.parametrize('prefix', ['res_cgs_', 'config2_cgs_'])
def test_cgs_matches_reference(eb, golden, prefix, capsys):
    """Device-resident CGS against CGS solves of the reference (tests/golden/cgs.npz); the same
    driver around the oracle's multigrid reproduces them to 2e-14 (tests/test_krylov_cpu.py).
    (Added after the round's GPU budget was spent: first run on a B200 is the driver's.)"""
    c = solve_case(golden('cgs'), prefix)
    grid = eb.TensorMesh([c['hx'], c['hy'], c['hz']], c['origin'])
    model = eb.Model(grid, **c['model'])
    sfield = eb.Field(grid, c['sfield'].copy(), frequency=c['frequency'])
    efield, info = eb.solve(model, sfield, return_info=True, order='lex', **c['kwargs'])
    capsys.readouterr()
    assert info['exit_message'] == c['exit_message']
    assert (info['it_ssl'], info['it_mg']) == (c['it_ssl'], c['it_mg'])
    assert np.abs(info['error_at_cycle'] - c['error_at_cycle']).max() <= 1e-8 * c['ref_error']
    assert rel_err(efield.field, c['efield']) <= 1e-8

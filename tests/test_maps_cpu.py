"""CPU checks of the interpolation row: the NumPy restatement (oracle/maps.py)
against the reference's outputs (tests/golden/maps.npz) and against SciPy itself (the third-party
library behind the reference's cubic interpolation), and the O(n) host logic of the product
(merged-node weights, sparse source fields) -- no GPU, no compute calls into the CUDA library."""
import numpy as np
import pytest

from conftest import rel_err
from helpers import maps_grid, split_field
from oracle import maps as om


def _nodes(hs, origin):
    return [np.r_[o, o + np.cumsum(h)] for h, o in zip(hs, origin)]


def test_oracle_volume_average_matches_reference(golden):
    gm = golden('maps')
    for k in range(int(gm['n_va'])):
        p = f'va{k}_'
        ni, no = _nodes(*maps_grid(gm, p, '_in')), _nodes(*maps_grid(gm, p, '_out'))
        for ax in range(3):
            w, ii, io = om.volume_average_weights(ni[ax], no[ax])
            assert np.array_equal(ii, gm[p + f'ii{ax}']) and np.array_equal(io, gm[p + f'io{ax}'])
            np.testing.assert_allclose(w, gm[p + f'w{ax}'], rtol=1e-15, atol=0)
        h = [np.diff(x) for x in no]
        vol = h[0][:, None, None] * h[1][None, :, None] * h[2][None, None, :]
        assert rel_err(om.interp_volume_average(ni, gm[p + 'values'], no, vol), gm[p + 'new']) < 1e-15
        logged = 10 ** om.interp_volume_average(ni, np.log10(gm[p + 'values']), no, vol)
        assert rel_err(logged, gm[p + 'new_log']) < 1e-14


def test_oracle_edges_to_vol_averages_matches_reference(golden):
    gm = golden('maps')
    for k in range(int(gm['n_ev'])):
        p = f'ev{k}_'
        hs, _ = maps_grid(gm, p)
        vol = hs[0][:, None, None] * hs[1][None, :, None] * hs[2][None, None, :]
        ex, ey, ez = split_field(vol.shape, gm[p + 'field'])
        for got, name in zip(om.interp_edges_to_vol_averages(ex, ey, ez, vol), ('ox', 'oy', 'oz')):
            assert rel_err(got, gm[p + name]) < 1e-15, (k, name)


@pytest.mark.parametrize('mode', ['constant', 'nearest'])
def test_oracle_cubic_spline_is_scipys(mode):
    """map_coordinates(order=3) restated: prefilter, weights, boundary rules of both modes the
    reference uses (maps.py:322-337), points inside, on and beyond the data."""
    ndi = pytest.importorskip('scipy.ndimage')
    rng = np.random.default_rng(3)
    for shape in [(9, 7, 11), (5, 6, 4), (2, 3, 30)]:
        for cplx in (False, True):
            d = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
            c = np.stack([rng.uniform(-1.5, s + 0.5, 150) for s in shape])
            c[:, :4] = np.array([[0, 0, 0], [s - 1 for s in shape], [0.5, 0, 1], [1, 1, 1]]).T
            want = ndi.map_coordinates(d, c, order=3, mode=mode, cval=np.nan)
            got = om.map_coordinates3(d, c, mode=mode, cval=np.nan)
            assert np.array_equal(np.isnan(want), np.isnan(got))
            ok = ~np.isnan(want)
            assert np.abs(want[ok] - got[ok]).max() < 1e-13 * np.abs(want[ok]).max()
            for m, mine in (('mirror', 'mirror'), ('reflect', 'reflect'), ('nearest', 'reflect'), ('constant', 'mirror')):
                ref = ndi.spline_filter(d.real, order=3, mode=m)
                assert np.abs(ref - om.spline_filter3(d.real, mine)).max() < 1e-12 * np.abs(ref).max()


def test_product_volume_average_weights(golden):
    """The product's own (bisection) form of maps._volume_average_weights against the reference."""
    from emg3d_b200 import maps
    gm = golden('maps')
    for k in range(int(gm['n_va'])):
        p = f'va{k}_'
        ni, no = _nodes(*maps_grid(gm, p, '_in')), _nodes(*maps_grid(gm, p, '_out'))
        for ax in range(3):
            w, ii, io = maps._volume_average_weights(ni[ax], no[ax])
            assert np.array_equal(ii, gm[p + f'ii{ax}']) and np.array_equal(io, gm[p + f'io{ax}'])
            np.testing.assert_allclose(w, gm[p + f'w{ax}'], rtol=1e-15, atol=0)
            assert np.all(np.diff(io) >= 0)          # what the per-output-cell offsets rely on


def test_source_field_is_sparse_and_equals_the_dense_reference(golden):
    """get_source_field keeps (indices, values, background); the dense array it materialises on
    demand is the reference's (tests/golden/host.npz), and touching it retires the sparse form."""
    import emg3d_b200 as eb
    gh = golden('host')
    grid = eb.TensorMesh([gh['hx'], gh['hy'], gh['hz']], gh['origin'])
    k = 0
    while f'src{k}' in gh.files:
        for freq in (1.0, -2.5):
            sf = eb.get_source_field(grid, gh[f'src{k}'], freq)
            idx, val, bg = sf.sparse
            assert idx.size < 200 and np.all(np.diff(idx) > 0)
            assert sf.dtype == (np.complex128 if freq > 0 else np.float64)
            dense = np.full(grid.n_edges, bg)
            dense[idx] = val
            want = gh[f'src{k}_f{freq}']
            assert rel_err(dense, want) < 1e-14
            assert rel_err(sf.field, want) < 1e-14 and sf.sparse is None
        k += 1
    assert k > 0


def test_sample_box_covers_the_reach_of_the_prefilter():
    from emg3d_b200 import maps
    shape = (300, 40, 200)
    c = [np.array([120.3, 131.9]), np.array([-3.0, 10.2, 45.0]), np.array([199.0])]
    lo, m = maps._sample_box(c, shape)
    assert lo == (120 - 48, 0, 199 - 48) and m == (132 + 48 + 1 - lo[0], 40, 200 - lo[2])
    lo, m = maps._sample_box([np.array([-5.0]), np.array([1.0]), np.array([1.0])], shape)   # nothing inside
    assert all(a >= 0 and a + b <= n for a, b, n in zip(lo, m, shape))
    assert maps._SPLINE_MARGIN >= 48 and abs(om.POLE) ** maps._SPLINE_MARGIN < 1e-27


def test_receiver_coordinates_formats():
    import emg3d_b200 as eb
    from emg3d_b200 import fields

    class Rx:
        def __init__(self, c):
            self.coordinates = c
    assert fields.receiver_coordinates(Rx((1, 2, 3, 4, 5))) == (1, 2, 3, 4, 5)
    got = fields.receiver_coordinates([Rx((1., 2., 3., 4., 5.)), Rx((6., 7., 8., 9., 10.))])
    assert np.array_equal(np.array(got), np.array([[1., 6.], [2., 7.], [3., 8.], [4., 9.], [5., 10.]]))
    with pytest.raises(ValueError, match='needs to be in the form'):
        fields.receiver_coordinates((1, 2, 3))
    assert eb.get_receiver is fields.get_receiver

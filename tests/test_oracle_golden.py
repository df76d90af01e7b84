"""Pin the CPU oracle (oracle/) against golden vectors produced by the reference.

Tolerances are norm-wise relative errors.  The reference's numba kernels run
with fastmath (reassociation, contraction), the oracle is plain C with a
different but algebraically identical summation order, so agreement is to
rounding: <= 1e-12 for single kernel calls on these well-conditioned models.
"""
import numpy as np
import pytest

import oracle
from oracle import mg
from conftest import rel_err
from helpers import hfield_case, kernel_case, solve_case, split_faces, split_field

KTOL = 1e-12


def test_amat_x(golden):
    gk = golden('kernels')
    for k in range(int(gk['n_cases'])):
        c = kernel_case(gk, k)
        r = c['s'].copy()
        oracle.amat_x(*split_field(c['shape'], r), *split_field(c['shape'], c['e']),
                      c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'], c['hx'], c['hy'], c['hz'])
        assert rel_err(r, c['r']) < 1e-14


@pytest.mark.parametrize('ldir', [0, 1, 2, 3])
def test_gauss_seidel(golden, ldir):
    gk = golden('kernels')
    fn = [oracle.gauss_seidel, oracle.gauss_seidel_x, oracle.gauss_seidel_y,
          oracle.gauss_seidel_z][ldir]
    for k in range(int(gk['n_cases'])):
        c = kernel_case(gk, k)
        for nu in (1, 2):
            e = c['e'].copy()
            fn(*split_field(c['shape'], e), *split_field(c['shape'], c['s']),
               c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'], c['hx'], c['hy'], c['hz'], nu)
            assert rel_err(e, gk[c['prefix'] + f'gs{ldir}_nu{nu}']) < KTOL


@pytest.mark.parametrize('ldir', [0, 1, 2, 3])
def test_relaxing_one_colour_twice_changes_nothing(golden, ldir):
    """Blocks of one colour class do not interact, so a second relaxation of the class
    reproduces the first (to rounding).  The CUDA smoothers rely on this: consecutive
    sweeps run the colours in opposite order and the repeated colour is skipped
    (csrc/gs_line.cu gs_dir, csrc/gs_point.cu launch_gs_point)."""
    gk = golden('kernels')
    c = kernel_case(gk, 0)
    shape = c['shape']
    seq = oracle.color_sequence(ldir, shape, 2)           # sweep 1 descending, sweep 2 ascending
    n = len(seq) // 2
    par = seq % 2
    k = 1                                                 # length of the first class of sweep 2
    while k < n and np.array_equal(par[n + k], par[n]):
        k += 1
    # the class that ends sweep 1 is the class that starts sweep 2
    assert {tuple(r) for r in seq[n - k:n]} == {tuple(r) for r in seq[n:n + k]}
    args = (c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'], c['hx'], c['hy'], c['hz'])
    e = c['e'].copy()
    oracle.gs_sequence(ldir, *split_field(shape, e), *split_field(shape, c['s']), *args, seq[:n])
    e2 = e.copy()
    oracle.gs_sequence(ldir, *split_field(shape, e2), *split_field(shape, c['s']), *args, seq[n:n + k])
    assert rel_err(e2, e) < 1e-13


def test_solve_known_answer():
    # band LDL^T against a dense solve (the reference tests do the same,
    # tests/test_core.py:203-262), real and complex, n = 6 and a long band
    rng = np.random.default_rng(5)
    for n, cplx in ((6, False), (6, True), (41, True), (1, True), (3, False)):
        A = np.zeros((n, n), dtype=complex if cplx else float)
        for i in range(n):
            for j in range(max(0, i - 5), i + 1):
                v = rng.standard_normal() + (1j * rng.standard_normal() if cplx else 0)
                A[i, j] = A[j, i] = v
            A[i, i] += 10
        b = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
        amat = np.zeros(6 * n, dtype=A.dtype)
        for j in range(n):
            for i in range(j, min(n, j + 6)):
                amat[i + 5 * j] = A[i, j]
        x = b.copy()
        oracle.solve(amat, x)
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-11)


def test_transfer(golden):
    gt = golden('transfer')
    for k in range(int(gt['n_cases'])):
        p = f"t{k}_"
        h = [gt[p + 'hx'], gt[p + 'hy'], gt[p + 'hz']]
        shape = tuple(a.size for a in h)
        g = mg.Grid(h, gt[p + 'origin'])
        for sc in gt[p + 'sc_dirs']:
            q = p + f"sc{sc}_"
            vm = mg.VolumeModel.from_arrays(
                g, *(np.asfortranarray(gt[p + n]) for n in ('eta_x', 'eta_y', 'eta_z', 'zeta')))
            cvm, cs, _ = mg.restriction(vm, gt[p + 'r'].copy(), sc)
            assert rel_err(cs, gt[q + 'cs']) < 1e-14
            for n, arr in (('ceta_x', cvm.eta_x), ('ceta_y', cvm.eta_y),
                           ('ceta_z', cvm.eta_z), ('czeta', cvm.zeta)):
                assert rel_err(arr, gt[q + n]) < 1e-15
            # weights (dummy ones where the axis is not coarsened)
            fl = oracle.SC_FLAGS[int(sc)]
            for a, ax in enumerate('xyz'):
                if fl[a]:
                    w = oracle.restrict_weights(g.nodes[a], g.centers[a], g.h[a],
                                                cvm.grid.nodes[a], cvm.grid.centers[a],
                                                cvm.grid.h[a])
                    np.testing.assert_allclose(np.array(w), gt[q + f'w{ax}'], rtol=1e-12, atol=1e-14)
            e = gt[p + 'e'].copy()
            mg.prolongation(g, e, cvm.grid, gt[q + 'ce'].copy(), sc)
            assert rel_err(e, gt[q + 'e_out']) < 1e-15
        assert shape == g.shape_cells


def _oracle_solve(c, **extra):
    g = mg.Grid([c['hx'], c['hy'], c['hz']], c['origin'])
    m = c['model']
    vm = mg.VolumeModel(g, m['property_x'], m.get('property_y'), m.get('property_z'),
                        m.get('mu_r'), m.get('epsilon_r'), c['frequency'])
    kw = {k: v for k, v in c['kwargs'].items() if k not in ('verb', 'plain')}
    if c['kwargs'].get('plain'):
        kw.setdefault('sslsolver', False)
        kw.setdefault('semicoarsening', False)
        kw.setdefault('linerelaxation', False)
    else:
        kw.setdefault('sslsolver', True)
        kw.setdefault('semicoarsening', True)
        kw.setdefault('linerelaxation', True)
    kw.update(extra)
    return mg.solve(vm, c['sfield'].copy(), **kw)


@pytest.mark.parametrize('prefix', ['res_F_', 'res_W_', 'res_V_', 'res_bic_', 'reg2_',
                                    'lap_F_', 'lap_bic_', 'config1_', 'config5_'])
def test_solves(golden, prefix):
    """Full solves: same iteration counts, per-cycle errors and fields."""
    c = solve_case(golden('solves'), prefix)
    e, info = _oracle_solve(c)
    assert info['it_mg'] == c['it_mg']
    assert info['it_ssl'] == c['it_ssl']
    assert info['exit_message'] == c['exit_message']
    np.testing.assert_allclose(info['error_at_cycle'] / c['ref_error'],
                               c['error_at_cycle'] / c['ref_error'], rtol=1e-6, atol=1e-12)
    assert rel_err(e, c['efield']) < 1e-9
    if prefix + 'regression' in golden('solves').files:
        # the reference's own regression data (tests/data/regression.npz)
        np.testing.assert_allclose(e, golden('solves')[prefix + 'regression'],
                                   rtol=1e-6, atol=1e-14 * np.abs(c['efield']).max() + 1e-30)


def test_edge_curl_factor_and_magnetic_field(golden):
    """fields._edge_curl_factor and get_magnetic_field (fields.py:617-659, 941-1009)."""
    gh = golden('hfield')
    for k in range(int(gh['n_cases'])):
        c = hfield_case(gh, k)
        shape = c['shape']
        h = np.zeros_like(c['h_k'])
        oracle.edge_curl_factor(*split_faces(shape, h), *split_field(shape, c['e']),
                                c['hx'], c['hy'], c['hz'], c['zeta_k'])
        assert rel_err(h, c['h_k']) < 1e-14
        # the caller: zeta = V / mu_r divided by s mu_0
        g = mg.Grid([c['hx'], c['hy'], c['hz']])
        vm = mg.VolumeModel(g, c['property_x'], c['property_y'], c['property_z'], c['mu_r'], None,
                            c['frequency'])
        f = c['frequency']
        smu0 = (-f if f < 0 else 2j * np.pi * f) * mg.MU_0
        h = np.zeros_like(c['h'])
        oracle.edge_curl_factor(*split_faces(shape, h), *split_field(shape, c['e']),
                                c['hx'], c['hy'], c['hz'], vm.zeta / smu0)
        assert rel_err(h, c['h']) < 1e-14


@pytest.mark.parametrize('d', [0, 1, 2])
def test_line_system_structure_and_reduced_elimination(golden, d):
    """The algebra the CUDA line smoothers rely on (csrc/gs_line.cu),
    checked on the operator itself: in the system of one grid line a line edge L_i
    couples only to the transverse edges at its two end nodes, with opposite signs
    (c_i = -f_i); transverse edges of neighbouring nodes couple diagonally; and
    eliminating the L_i first leaves a block-tridiagonal system of 4x4 blocks with
    E_m = diag(d) + f f^T / dL whose forward / backward recurrences reproduce the dense
    solve of the line system (= what the reference's banded LDL^T computes)."""
    gk = golden('kernels')
    c = kernel_case(gk, 0)                                 # (6, 4, 8), complex, triaxial
    shape = c['shape']
    args = (c['eta_x'], c['eta_y'], c['eta_z'], c['zeta'], c['hx'], c['hy'], c['hz'])
    p, q = (1 if d == 0 else 0), (1 if d == 2 else 2)
    N = shape[d]
    tp, tq = 2, 3                                          # an interior line
    offs, shp = [], []
    for comp in range(3):
        sh = tuple(shape[a] + (a != comp) for a in range(3))
        offs.append(sum(int(np.prod(x)) for x in shp))
        shp.append(sh)

    def eidx(comp, idx):
        sh = shp[comp]
        return offs[comp] + idx[0] + sh[0] * (idx[1] + sh[1] * idx[2])

    def pos(**kw):
        i = [0, 0, 0]
        for a, v in kw.items():
            i[int(a[1])] = v
        return i

    L = [eidx(d, pos(**{f'a{d}': i, f'a{p}': tp, f'a{q}': tq})) for i in range(N)]
    T = [[eidx(p, pos(**{f'a{d}': m, f'a{p}': tp - 1, f'a{q}': tq})),
          eidx(p, pos(**{f'a{d}': m, f'a{p}': tp, f'a{q}': tq})),
          eidx(q, pos(**{f'a{d}': m, f'a{p}': tp, f'a{q}': tq - 1})),
          eidx(q, pos(**{f'a{d}': m, f'a{p}': tp, f'a{q}': tq}))] for m in range(N + 1)]
    n_all = c['e'].size

    def apply_A(e):
        r = np.zeros(n_all, dtype=complex)
        oracle.amat_x(*split_field(shape, r), *split_field(shape, e), *args)
        return -r                                           # amat_x: r -= A e

    unknowns = L + [j for m in range(1, N) for j in T[m]]
    cols = {}
    for j in unknowns:
        e = np.zeros(n_all, dtype=complex)
        e[j] = 1.0
        cols[j] = apply_A(e)
    A = lambda i, j: cols[j][i]
    scale = max(abs(A(j, j)) for j in unknowns)
    # structure
    for i in range(N):
        for j in range(N):
            if i != j:
                assert abs(A(L[i], L[j])) == 0
        for m in range(1, N):
            for k in range(4):
                v = A(L[i], T[m][k])
                if m not in (i, i + 1):
                    assert abs(v) == 0
        if 1 <= i and i + 1 <= N - 1:
            f = np.array([A(L[i], T[i][k]) for k in range(4)])
            cc = np.array([A(L[i], T[i + 1][k]) for k in range(4)])
            assert np.all(np.abs(f.imag) < 1e-15 * scale) and np.allclose(cc, -f, rtol=1e-13, atol=0)
    for m in range(2, N):
        B = np.array([[A(T[m][r], T[m - 1][k]) for k in range(4)] for r in range(4)])
        assert np.all(np.abs(B - np.diag(np.diag(B))) == 0) and np.all(np.abs(B.imag) < 1e-15 * scale)

    # right-hand side of the line system for the golden field: b = s - A e with the
    # line's unknowns removed from e
    e0 = c['e'].copy()
    e0[unknowns] = 0
    b_all = c['s'] - apply_A(e0)
    M = np.array([[A(i, j) for j in unknowns] for i in unknowns])
    x_dense = np.linalg.solve(M, b_all[unknowns])

    # reduced elimination (fixed end-plane values are already inside b here)
    dL = np.array([A(L[i], L[i]) for i in range(N)])
    bL = np.array([b_all[L[i]] for i in range(N)])
    # couplings of L_i to the transverse edges at its lower (f) and upper (cu) node
    f = [np.array([A(L[i], T[i][k]) for k in range(4)]).real if i >= 1 else None for i in range(N)]
    cu = [np.array([A(L[i], T[i + 1][k]) for k in range(4)]).real if i + 1 <= N - 1 else None
          for i in range(N)]
    Tsol = {}
    X, g = {}, {}
    for m in range(1, N):
        C = np.array([[A(T[m][r], T[m][k]) for k in range(4)] for r in range(4)])
        D = C - np.outer(cu[m - 1], cu[m - 1]) / dL[m - 1]
        r = np.array([b_all[j] for j in T[m]]) - cu[m - 1] * bL[m - 1] / dL[m - 1]
        if m <= N - 1 and f[m] is not None:
            D = D - np.outer(f[m], f[m]) / dL[m]
            r = r - f[m] * bL[m] / dL[m]
        if m >= 2:
            dk = np.array([A(T[m][k], T[m - 1][k]) for k in range(4)]).real
            E = np.diag(dk) + np.outer(f[m - 1], f[m - 1]) / dL[m - 1]
            np.testing.assert_allclose(E, np.diag(dk) - np.outer(cu[m - 1], f[m - 1]) / dL[m - 1],
                                       rtol=1e-13)
            D = D - E @ X[m - 1] @ E
            r = r - E @ g[m - 1]
        X[m] = np.linalg.inv(D)
        g[m] = X[m] @ r
    for m in range(N - 1, 0, -1):
        if m == N - 1:
            Tsol[m] = g[m]
        else:
            dk = np.array([A(T[m + 1][k], T[m][k]) for k in range(4)]).real
            E = np.diag(dk) + np.outer(f[m], f[m]) / dL[m]
            Tsol[m] = g[m] - X[m] @ (E @ Tsol[m + 1])
    Lsol = np.zeros(N, dtype=complex)
    for i in range(N):
        acc = bL[i]
        if i >= 1:
            acc = acc - f[i] @ Tsol[i]
        if i + 1 <= N - 1:
            acc = acc - cu[i] @ Tsol[i + 1]
        Lsol[i] = acc / dL[i]
    x_red = np.r_[Lsol, np.concatenate([Tsol[m] for m in range(1, N)])]
    assert rel_err(x_red, x_dense) < 1e-10

    # and the dense solve is what one line relaxation of the oracle does
    e1 = c['e'].copy()
    oracle.gs_sequence(d + 1, *split_field(shape, e1), *split_field(shape, c['s']), *args,
                       np.array([[tp, tq]], dtype=np.int32))
    assert rel_err(e1[unknowns], x_dense) < 1e-10


def test_orderings_respect_the_block_interactions():
    """What the parallel orderings of the CUDA smoothers assume,
    checked on the operator: two point blocks (nodes) interact -- share an edge or are
    coupled by A -- only if they are face or face-diagonal neighbours; then the
    hyperplane index t = ix + 2 iy + 3 iz differs with the sign of the lexicographic
    order (so hyperplane-by-hyperplane == the reference's sequential sweep) and their
    parity classes differ (8 colours are conflict-free).  Same for lines with
    t = tp + 2 tq and 4 colours; 2 colours would not be enough."""
    rng = np.random.default_rng(0)
    shape = (5, 4, 6)
    g = mg.Grid([rng.uniform(1, 2, n) for n in shape])
    vm = mg.VolumeModel(g, rng.uniform(1, 10, shape), rng.uniform(1, 10, shape),
                        rng.uniform(1, 10, shape), rng.uniform(1, 2, shape), None, 1.0)
    args = (vm.eta_x, vm.eta_y, vm.eta_z, vm.zeta, *g.h)
    n_all = g.n_edges
    shp = [tuple(shape[a] + (a != comp) for a in range(3)) for comp in range(3)]
    offs = [0, int(np.prod(shp[0])), int(np.prod(shp[0])) + int(np.prod(shp[1]))]

    def eidx(comp, i, j, k):
        return offs[comp] + i + shp[comp][0] * (j + shp[comp][1] * k)

    nz = np.zeros((n_all, n_all), dtype=bool)
    for j in range(n_all):
        e = np.zeros(n_all, dtype=complex)
        e[j] = 1.0
        r = np.zeros(n_all, dtype=complex)
        oracle.amat_x(*split_field(shape, r), *split_field(shape, e), *args)
        nz[:, j] = r != 0
    nx, ny, nzc = shape

    def node_block(ix, iy, iz):
        return {eidx(0, ix - 1, iy, iz), eidx(0, ix, iy, iz), eidx(1, ix, iy - 1, iz),
                eidx(1, ix, iy, iz), eidx(2, ix, iy, iz - 1), eidx(2, ix, iy, iz)}

    def interact(a, b):
        return bool(a & b) or bool(nz[np.ix_(sorted(a), sorted(b))].any())

    nodes = [(ix, iy, iz) for iz in range(1, nzc) for iy in range(1, ny) for ix in range(1, nx)]
    blocks = {n: node_block(*n) for n in nodes}
    for ia, a in enumerate(nodes):
        for b in nodes[ia + 1:]:                               # a before b lexicographically
            if not interact(blocks[a], blocks[b]):
                continue
            dx, dy, dz = (b[0] - a[0], b[1] - a[1], b[2] - a[2])
            assert max(abs(dx), abs(dy), abs(dz)) == 1 and (dx, dy, dz).count(0) >= 1
            assert (b[0] + 2 * b[1] + 3 * b[2]) > (a[0] + 2 * a[1] + 3 * a[2])
            assert (dx % 2, dy % 2, dz % 2) != (0, 0, 0)
    # x-lines: blocks = all edges touching the interior nodes of the line (+ its own edges)
    lines = [(iy, iz) for iz in range(1, nzc) for iy in range(1, ny)]
    lblocks = {}
    for iy, iz in lines:
        bl = {eidx(0, i, iy, iz) for i in range(nx)}
        for m in range(1, nx):
            bl |= {eidx(1, m, iy - 1, iz), eidx(1, m, iy, iz), eidx(2, m, iy, iz - 1), eidx(2, m, iy, iz)}
        lblocks[(iy, iz)] = bl
    diagonal_pairs = 0
    for ia, a in enumerate(lines):
        for b in lines[ia + 1:]:
            if not interact(lblocks[a], lblocks[b]):
                continue
            dp, dq = b[0] - a[0], b[1] - a[1]
            assert max(abs(dp), abs(dq)) == 1
            assert b[0] + 2 * b[1] > a[0] + 2 * a[1]
            assert (dp % 2, dq % 2) != (0, 0)
            diagonal_pairs += abs(dp) == 1 and abs(dq) == 1
    assert diagonal_pairs > 0          # diagonal neighbours interact: red-black is not enough


def test_point_block_structure():
    """Structure of the 6x6 node system the CUDA point smoother exploits
    (csrc/gs_point.cu NodeSys): the two x-edges do not couple, neither do the two y-
    nor the two z-edges, all off-diagonal entries are real, and eliminating the x-edges
    first (a 4x4 Schur complement) reproduces the dense solve."""
    rng = np.random.default_rng(1)
    shape = (4, 5, 3)
    g = mg.Grid([rng.uniform(1, 2, n) for n in shape])
    vm = mg.VolumeModel(g, rng.uniform(1, 10, shape), rng.uniform(1, 10, shape),
                        rng.uniform(1, 10, shape), rng.uniform(1, 2, shape),
                        rng.uniform(1, 5, shape), 2.0)
    args = (vm.eta_x, vm.eta_y, vm.eta_z, vm.zeta, *g.h)
    n_all = g.n_edges
    shp = [tuple(shape[a] + (a != comp) for a in range(3)) for comp in range(3)]
    offs = [0, int(np.prod(shp[0])), int(np.prod(shp[0])) + int(np.prod(shp[1]))]
    eidx = lambda comp, i, j, k: offs[comp] + i + shp[comp][0] * (j + shp[comp][1] * k)
    ix, iy, iz = 2, 3, 1
    blk = [eidx(0, ix - 1, iy, iz), eidx(0, ix, iy, iz), eidx(1, ix, iy - 1, iz),
           eidx(1, ix, iy, iz), eidx(2, ix, iy, iz - 1), eidx(2, ix, iy, iz)]
    M = np.zeros((6, 6), dtype=complex)
    for c, j in enumerate(blk):
        e = np.zeros(n_all, dtype=complex)
        e[j] = 1.0
        r = np.zeros(n_all, dtype=complex)
        oracle.amat_x(*split_field(shape, r), *split_field(shape, e), *args)
        M[:, c] = -r[blk]
    assert np.allclose(M, M.T, rtol=1e-14, atol=0)                       # complex symmetric
    assert M[0, 1] == 0 and M[2, 3] == 0 and M[4, 5] == 0
    off = M - np.diag(np.diag(M))
    assert np.all(off.imag == 0)
    b = rng.standard_normal(6) + 1j * rng.standard_normal(6)
    x = np.linalg.solve(M, b)
    dX, B, C = np.diag(M)[:2], M[:2, 2:].real, M[2:, 2:]
    S = C - B.T @ np.diag(1 / dX) @ B
    xT = np.linalg.solve(S, b[2:] - B.T @ (b[:2] / dX))
    xX = (b[:2] - B @ xT) / dX
    # the node system is ill-conditioned (curl-curl is singular on the gradient of the
    # node's potential; only the eta term regularises it), so two elimination orders
    # agree to cond x eps only -- the reason for the 1e-11 / 2e-10 bars of the GPU tests
    assert np.linalg.cond(M) > 1e3
    assert rel_err(np.r_[xX, xT], x) < 1e-16 * np.linalg.cond(M) * 10

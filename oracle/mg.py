"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's multigrid
driver on top of the C oracle kernels.

Follows emg3d/solver.py: ``solve`` (52-449), ``multigrid`` (471-649), ``krylov``
(652-784), ``smoothing`` (788-846), ``restriction`` (849-944), ``prolongation``
(947-1019), ``residual`` (1022-1070), ``MGParameters`` level/cycle rules
(1202-1381), ``_current_sc_dir`` (1482-1531), ``_current_lr_dir`` (1534-1588),
``_terminate`` (1591-1664), ``_restrict_model_parameters`` (1667-1718) and
models.VolumeModel (emg3d/models.py:654-691).  Works on bare arrays; no logging.
Like the reference it rebuilds the coarse grid, model and weights at every visit.
"""
import itertools

import numpy as np
import scipy.sparse.linalg as ssl

from . import (SC_FLAGS, amat_x, gauss_seidel, gauss_seidel_x, gauss_seidel_y,
               gauss_seidel_z, prolong, restrict, restrict_weights)

from scipy.constants import epsilon_0 as EPS_0, mu_0 as MU_0  # as the reference


class Grid:
    """Tensor mesh: widths, nodes, centres (emg3d/meshes.py:72-107)."""

    def __init__(self, h, origin=(0., 0., 0.)):
        self.h = [np.array(a, dtype=float) for a in h]
        self.origin = np.array(origin, dtype=float)
        self.shape_cells = tuple(a.size for a in self.h)
        self.shape_nodes = tuple(a.size + 1 for a in self.h)
        self.nodes = [np.r_[0., a.cumsum()] + o for a, o in zip(self.h, self.origin)]
        self.centers = [(n[1:] + n[:-1]) / 2 for n in self.nodes]
        nx, ny, nz = self.shape_cells
        self.shape_edges = ((nx, ny + 1, nz + 1), (nx + 1, ny, nz + 1),
                            (nx + 1, ny + 1, nz))
        self.n_edges = sum(int(np.prod(s)) for s in self.shape_edges)
        self.n_cells = nx * ny * nz

    def split(self, field):
        """Views (fx, fy, fz) of a 1-D field array (emg3d/fields.py:201-259)."""
        out, i0 = [], 0
        for s in self.shape_edges:
            n = int(np.prod(s))
            out.append(field[i0:i0 + n].reshape(s, order='F'))
            i0 += n
        return out


class VolumeModel:
    """eta_a = -s mu0 V (sigma_a + s eps0 eps_r), zeta = V / mu_r."""

    def __init__(self, grid, res_x, res_y=None, res_z=None, mu_r=None,
                 epsilon_r=None, frequency=1.0):
        self.grid = grid
        sval = -frequency if frequency < 0 else 2j * np.pi * frequency
        smu0 = sval * MU_0
        shape = grid.shape_cells
        vol = (grid.h[0][:, None, None] * grid.h[1][None, :, None] *
               grid.h[2][None, None, :])

        def eta(res):
            cond = 1.0 / (np.asarray(res, dtype=float) * np.ones(shape))
            if epsilon_r is None:
                return np.asfortranarray(-smu0 * vol * cond)
            smu = sval * EPS_0 * (np.asarray(epsilon_r) * np.ones(shape))
            return np.asfortranarray(-smu0 * vol * (cond + smu))

        self.eta_x = eta(res_x)
        self.eta_y = self.eta_x if res_y is None else eta(res_y)
        self.eta_z = self.eta_x if res_z is None else eta(res_z)
        zeta = vol.copy()
        if mu_r is not None:
            zeta = zeta / (np.asarray(mu_r) * np.ones(shape))
        self.zeta = np.asfortranarray(zeta)

    @classmethod
    def from_arrays(cls, grid, eta_x, eta_y, eta_z, zeta):
        self = cls.__new__(cls)
        self.grid = grid
        self.eta_x, self.eta_y, self.eta_z, self.zeta = eta_x, eta_y, eta_z, zeta
        return self


def _margs(vm):
    g = vm.grid
    return (vm.eta_x, vm.eta_y, vm.eta_z, vm.zeta, g.h[0], g.h[1], g.h[2])


def residual(vm, s, e, norm=False):
    g = vm.grid
    r = s.copy()
    amat_x(*g.split(r), *g.split(e), *_margs(vm))
    return np.linalg.norm(r) if norm else r


def current_lr_dir(lr_dir, shape):
    """Drop line directions along which the grid has only two cells."""
    dirs = {0: (), 1: (0,), 2: (1,), 3: (2,), 4: (1, 2), 5: (0, 2), 6: (0, 1),
            7: (0, 1, 2)}[int(lr_dir)]
    keep = tuple(d for d in dirs if shape[d] != 2)
    inv = {(): 0, (0,): 1, (1,): 2, (2,): 3, (1, 2): 4, (0, 2): 5, (0, 1): 6,
           (0, 1, 2): 7}
    return inv[keep]


def current_sc_dir(sc_dir, shape):
    """Which axes can and shall be halved on this grid."""
    stop = [shape[a] % 2 != 0 or shape[a] < 3 or sc_dir == a + 1 for a in range(3)]
    table = {(0, 0, 0): 0, (1, 0, 0): 1, (0, 1, 0): 2, (0, 0, 1): 3,
             (0, 1, 1): 4, (1, 0, 1): 5, (1, 1, 0): 6, (1, 1, 1): 6}
    return table[tuple(int(b) for b in stop)]


def smoothing(vm, s, e, nu, lr_dir):
    g = vm.grid
    c = current_lr_dir(lr_dir, g.shape_cells)
    args = (*g.split(e), *g.split(s), *_margs(vm), nu)
    if c == 0:
        gauss_seidel(*args)
    if c in (1, 5, 6, 7):
        gauss_seidel_x(*args)
    if c in (2, 4, 6, 7):
        gauss_seidel_y(*args)
    if c in (3, 4, 5, 7):
        gauss_seidel_z(*args)


def restrict_param(p, sc_dir):
    """Coarse cell value = sum over its 8/4/2 fine cells."""
    fl = SC_FLAGS[int(sc_dir)]
    out = p
    for a in range(3):
        if fl[a]:
            lo = [slice(None)] * 3
            hi = [slice(None)] * 3
            lo[a], hi[a] = slice(0, None, 2), slice(1, None, 2)
            out = out[tuple(lo)] + out[tuple(hi)]
    return np.asfortranarray(out)


def restriction(vm, res, sc_dir):
    g = vm.grid
    fl = SC_FLAGS[int(sc_dir)]
    ch = [np.diff(g.nodes[a][::2 if fl[a] else 1]) for a in range(3)]
    cg = Grid(ch, g.origin)
    cex = restrict_param(vm.eta_x, sc_dir)
    cey = cex if vm.eta_y is vm.eta_x else restrict_param(vm.eta_y, sc_dir)
    cez = cex if vm.eta_z is vm.eta_x else restrict_param(vm.eta_z, sc_dir)
    cvm = VolumeModel.from_arrays(cg, cex, cey, cez, restrict_param(vm.zeta, sc_dir))
    w = []
    for a in range(3):
        if fl[a]:
            w.append(restrict_weights(g.nodes[a], g.centers[a], g.h[a],
                                      cg.nodes[a], cg.centers[a], cg.h[a]))
        else:
            z = np.zeros(g.shape_nodes[a])
            w.append((z, np.ones(g.shape_nodes[a]), z))
    cs = np.zeros(cg.n_edges, dtype=res.dtype)
    restrict(*cg.split(cs), *g.split(res), *w, sc_dir)
    return cvm, cs, np.zeros(cg.n_edges, dtype=res.dtype)


def prolongation(g, e, cg, ce, sc_dir):
    prolong(*g.split(e), *cg.split(ce), g.nodes, cg.nodes, sc_dir)


class Params:
    """Cycle bookkeeping (solver.py:1202-1232, 1272-1381)."""

    def __init__(self, shape, cycle='F', semicoarsening=False, linerelaxation=False,
                 sslsolver=False, tol=1e-6, maxit=50, nu_init=0, nu_pre=2,
                 nu_coarse=1, nu_post=2, clevel=-1):
        self.cycle, self.tol, self.maxit = cycle, tol, maxit
        self.nu_init, self.nu_pre, self.nu_coarse, self.nu_post = (
            nu_init, nu_pre, nu_coarse, nu_post)
        cl = np.zeros(3, dtype=int)
        for a in range(3):
            n = shape[a]
            while n % 2 == 0 and n > 2:
                cl[a] += 1
                n //= 2
            if -1 < clevel < cl[a]:
                cl[a] = clevel
        self.clevel = [cl.max(), max(cl[1], cl[2]), max(cl[0], cl[2]),
                       max(cl[0], cl[1])]

        def cyc(flag, default, hi):
            if flag is True:
                seq = default
            elif flag is False or (isinstance(flag, (int, np.integer)) and 0 <= flag <= hi):
                return [int(flag)], None
            else:
                seq = [int(c) for c in str(abs(flag))]
            return seq, itertools.cycle(seq)

        self.sc_seq, self.sc_cycle = cyc(semicoarsening, [1, 2, 3], 3)
        self.lr_seq, self.lr_cycle = cyc(linerelaxation, [4, 5, 6], 7)
        self.sc_dir = next(self.sc_cycle) if self.sc_cycle else self.sc_seq[0]
        self.lr_dir = next(self.lr_cycle) if self.lr_cycle else self.lr_seq[0]
        self.cycmax = 2 if cycle in ('F', 'W') else 1
        self.maxcycle = max(len(self.sc_seq), len(self.lr_seq))
        self.sslsolver = 'bicgstab' if sslsolver is True else sslsolver
        self.ssl_maxit = 0
        if self.sslsolver:
            self.ssl_maxit = maxit
            if cycle is not None:
                self.maxit = self.maxcycle
        self.it = 0
        self.ssl_it = 0
        self.l2 = 1.0
        self.l2_refe = 1.0
        self.exit_message = ''
        self.error_at_cycle = [0.0]
        self.sweeps = 0               # cell-sweeps counted (for throughput)


class ConvergenceError(Exception):
    pass


def _terminate(var, l2_last, l2_stag, it):
    finished = abort = False
    if l2_last < var.tol * var.l2_refe:
        var.exit_message, finished = "CONVERGED", True
    elif l2_last > 10 * var.l2_refe or not np.isfinite(l2_last):
        var.exit_message, finished, abort = "DIVERGED", True, True
    elif it > 2 and l2_last >= l2_stag:
        var.exit_message, finished, abort = "STAGNATED", True, True
    elif it == var.maxit:
        if not var.sslsolver:
            var.exit_message = "MAX. ITERATION REACHED, NOT CONVERGED"
        finished = True
    if finished and var.sslsolver and abort:
        raise ConvergenceError
    return finished


def _count(var, vm, nu, lr_dir):
    c = current_lr_dir(lr_dir, vm.grid.shape_cells)
    ndirs = {0: 1, 1: 1, 2: 1, 3: 1, 4: 2, 5: 2, 6: 2, 7: 3}[c]
    var.sweeps += vm.grid.n_cells * nu * ndirs


def multigrid(vm, s, e, var, level=0, new_cycmax=0):
    it = 0
    if level == var.clevel[var.sc_dir]:
        cycmax = 1
    elif new_cycmax == 0 or var.cycle != 'F':
        cycmax = var.cycmax
    else:
        cycmax = new_cycmax
    cyc = 0
    l2_last = residual(vm, s, e, True) if level == 0 else 0.0
    l2_stag = np.ones(var.maxcycle) * l2_last
    if level == 0 and var.nu_init > 0:
        smoothing(vm, s, e, var.nu_init, var.lr_dir)
        _count(var, vm, var.nu_init, var.lr_dir)
    while level == 0 or it < cycmax:
        l2_stag[(it - 1) % var.maxcycle] = l2_last
        if level == var.clevel[var.sc_dir]:
            smoothing(vm, s, e, var.nu_coarse, var.lr_dir)
            _count(var, vm, var.nu_coarse, var.lr_dir)
        else:
            if var.nu_pre > 0:
                smoothing(vm, s, e, var.nu_pre, var.lr_dir)
                _count(var, vm, var.nu_pre, var.lr_dir)
            sc = current_sc_dir(var.sc_dir, vm.grid.shape_cells)
            res = residual(vm, s, e)
            cvm, cs, ce = restriction(vm, res, sc)
            multigrid(cvm, cs, ce, var, level + 1, cycmax - cyc)
            prolongation(vm.grid, e, cvm.grid, ce, sc)
            if var.nu_post > 0:
                smoothing(vm, s, e, var.nu_post, var.lr_dir)
                _count(var, vm, var.nu_post, var.lr_dir)
        it += 1
        if level > 0:
            cyc += 1
        else:
            var.it += 1
            l2_last = residual(vm, s, e, True)
            var.error_at_cycle.append(l2_last)
            if var.sc_cycle:
                var.sc_dir = next(var.sc_cycle)
            if var.lr_cycle:
                var.lr_dir = next(var.lr_cycle)
            if _terminate(var, l2_last, l2_stag[(it - 1) % var.maxcycle], it):
                break
    var.l2 = l2_last


def krylov(vm, s, e, var):
    g = vm.grid

    def amatvec(x):
        r = np.zeros_like(x)
        amat_x(*g.split(r), *g.split(np.ascontiguousarray(x)), *_margs(vm))
        return -r

    def mg_matvec(b):
        x = np.zeros_like(b)
        multigrid(vm, np.ascontiguousarray(b), x, var)
        return x

    n = s.size
    A = ssl.LinearOperator((n, n), dtype=s.dtype, matvec=amatvec)
    M = ssl.LinearOperator((n, n), dtype=s.dtype, matvec=mg_matvec) if var.cycle else None

    def callback(x):
        var.ssl_it += 1
        var.l2 = residual(vm, s, np.ascontiguousarray(x), True)
        var.error_at_cycle.append(var.l2)

    try:
        x, i = getattr(ssl, var.sslsolver)(A=A, b=s, x0=e, rtol=var.tol,
                                           maxiter=var.ssl_maxit, atol=1e-30, M=M,
                                           callback=callback)
        e[:] = x
    except ConvergenceError:
        i = -1
        e[:] = 0
        var.exit_message += " (returned field is zero)"
    if i < 0:
        if var.exit_message == '':
            var.exit_message = f"Error in {var.sslsolver} ({i})"
    elif i > 0:
        var.exit_message = "MAX. ITERATION REACHED, NOT CONVERGED"
    else:
        var.exit_message = "CONVERGED"


def solve(vm, s, efield=None, **kwargs):
    """Restatement of emg3d.solver.solve on bare arrays; returns (e, info)."""
    var = Params(vm.grid.shape_cells, **kwargs)
    var.l2_refe = np.linalg.norm(s)
    var.error_at_cycle[0] = var.l2_refe
    e = np.zeros_like(s) if efield is None else efield
    if var.sslsolver:
        krylov(vm, s, e, var)
    else:
        multigrid(vm, s, e, var)
    info = {'exit': int(var.exit_message != 'CONVERGED'),
            'exit_message': var.exit_message, 'abs_error': var.l2,
            'rel_error': var.l2 / var.l2_refe, 'ref_error': var.l2_refe,
            'it_mg': var.it, 'it_ssl': var.ssl_it,
            'error_at_cycle': np.array(var.error_at_cycle),
            'cell_sweeps': var.sweeps}
    return e, info

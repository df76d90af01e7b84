"""Distributed (z-slab) solves through emg3d_b200.solve(..., comm=) against the single-GPU solver.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py [case ...]

Every rank runs its slab; rank 0 additionally solves the whole problem on its GPU with the
single-GPU driver and compares fields, cycle counts and per-cycle residual norms.  A case is
``name:config:n:key=value,...`` (solver keywords; ints / bools / strings), e.g.
``sc:config2:128:sslsolver=False,cycle=F,semicoarsening=True,linerelaxation=False``.
"""
import json
import os
import sys
import time

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, parallel, recipes  # noqa: E402

DEFAULT = [
    "plainF:config5:128:plain=True,cycle=F",
    "sc:config2:128:sslsolver=False,cycle=F,semicoarsening=True,linerelaxation=False",
    "sc+xy:config2:128:sslsolver=False,cycle=F,semicoarsening=True,linerelaxation=6",
    "z-lines:config2:128:sslsolver=False,cycle=V,semicoarsening=False,linerelaxation=3",
    "sc+lr:config4:128:sslsolver=False,cycle=F,semicoarsening=True,linerelaxation=True",
    "default:config3:128:tol=1e-6,zshift=-130",
    # (the marine model with air is only defined to ~1e-6: tools/dist_ops_check.py measures the
    # linearity of one V-cycle at 1e-6 there, on one GPU and on many; 7e-14 on config 2)
    # BiCGSTAB's shadow residual is r0 = b, a point source.  Gauss-Seidel leaves the residual
    # exactly zero at the nodes it relaxed last; when those hold the whole source (a dipole at the
    # slab interface, relaxed by the last z-half of the last sweep), <r0, r1> = 0 to rounding and
    # SciPy's rule |rho| < eps^2 ends the iteration ("Error in bicgstab (-10)"): a property of
    # BiCGSTAB + point source + ordering (measured: rho_1 = 1e-28 against |r0| |r1| = 2e-15), not
    # of the distributed operators (tools/dist_ops_check.py).  The Krylov cases move the source off
    # the interface.
    "bicgstab-air:config3:128:sslsolver=bicgstab,cycle=V,semicoarsening=False,linerelaxation=False,tol=1e-6,zshift=-130",
    "bicgstab:config2:128:sslsolver=bicgstab,cycle=V,semicoarsening=False,linerelaxation=False,zshift=-130",
    "bicgstab-sc-xy:config4:128:sslsolver=bicgstab,cycle=F,semicoarsening=True,linerelaxation=6",
    "cgs:config2:128:sslsolver=cgs,cycle=V,semicoarsening=False,linerelaxation=False,tol=1e-6,zshift=-130",
]


def parse(case):
    name, config, n, kws = case.split(':', 3)
    kw = {}
    for item in filter(None, kws.split(',')):
        k, v = item.split('=')
        if v in ('True', 'False'):
            kw[k] = v == 'True'
        elif v.lstrip('-').isdigit():
            kw[k] = int(v)
        else:
            try:
                kw[k] = float(v)
            except ValueError:
                kw[k] = v
    return name, config, int(n), kw


def main():
    cases = [parse(c) for c in (sys.argv[1:] or DEFAULT)]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    dist.init_process_group('gloo')
    _lib.init(local_rank)

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = parallel.NcclComm(rank, world, bcast)
    ok = True
    for name, config, n, kw in cases:
        cfg = recipes.config(config, n)
        kw = dict(kw)
        shift = kw.pop('zshift', 0.0)                  # (see the note at DEFAULT)
        if shift:
            src = list(cfg['source'])
            src[2] += float(shift)
            cfg['source'] = tuple(src)
        grid = eb.TensorMesh(cfg['h'], cfg['origin'])
        model = eb.Model(grid, **cfg['model'])
        sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
        kw = dict(dict(tol=1e-8, maxit=50), **kw)
        _lib.sync()
        dist.barrier()
        t0 = time.perf_counter()
        try:
            kws = {k: v for k, v in kw.items() if k not in ('warm', 'nosingle')}
            if kw.get('nosingle'):                     # (the field stays on the GPUs)
                e, info = None, eb.solve(model, sfield, comm=comm, return_info=True, return_field=False, **kws)
            else:
                e, info = eb.solve(model, sfield, comm=comm, return_info=True, **kws)
        except Exception as err:                       # noqa: BLE001
            if rank == 0:
                print(json.dumps({'case': name, 'error': repr(err)}), flush=True)
            ok = False
            dist.barrier()
            continue
        _lib.sync()
        dt = time.perf_counter() - t0
        if kw.pop('nosingle', False):                  # the distributed solve alone (full-size runs)
            if rank == 0:
                hist = [float(f"{v:.3e}") for v in info['error_at_cycle'] / info['ref_error']]
                print(json.dumps({'case': name, 'shape': [int(v) for v in grid.shape_cells], 'nranks': world,
                                  'kw': kw, 'dist': {'it_mg': info['it_mg'], 'it_ssl': info['it_ssl'],
                                                     'exit': info['exit_message'], 'wall_s': round(dt, 3),
                                                     'rel_error': float(info['rel_error']), 'err_hist': hist}}),
                      flush=True)
            dist.barrier()
            continue
        warm = None
        if kw.get('warm'):                             # second solve on a live solver: warm timings
            kw2 = {k: v for k, v in kw.items() if k != 'warm'}
            for key in ('sslsolver', 'semicoarsening', 'linerelaxation'):      # solve()'s defaults
                kw2.setdefault(key, True)
            dmg = parallel.DistributedMultigrid(model, sfield, comm, semicoarsening=kw2['semicoarsening'],
                                                linerelaxation=kw2['linerelaxation'])
            dmg.solve(**kw2)
            _lib.sync()
            dist.barrier()
            t1 = time.perf_counter()
            i2 = dmg.solve(**kw2)
            _lib.sync()
            warm = {'wall_s': round(time.perf_counter() - t1, 3), 'it_mg': i2['it_mg'],
                    'cycle_s': [round(float(b - a), 4) for a, b in
                                zip([0] + list(i2['runtime_at_cycle'][:-1]), i2['runtime_at_cycle'])]}
            dmg.close()
            del dmg
            import gc
            gc.collect()
            dist.barrier()
        kw = {k: v for k, v in kw.items() if k != 'warm'}
        if rank == 0:
            ws1 = eb.Workspace()
            t0 = time.perf_counter()
            e1, i1 = eb.solve(model, sfield, return_info=True, workspace=ws1, **kw)
            _lib.sync()
            dt1 = time.perf_counter() - t0
            warm1 = None
            if warm is not None:
                t0 = time.perf_counter()
                _, i1w = eb.solve(model, sfield, return_info=True, workspace=ws1, **kw)
                _lib.sync()
                warm1 = {'wall_s': round(time.perf_counter() - t0, 3), 'it_mg': i1w['it_mg'],
                         'cycle_s': [round(float(b - a), 4) for a, b in
                                     zip([0] + list(i1w['runtime_at_cycle'][:-1]), i1w['runtime_at_cycle'])]}
            ws1.clear()
            err = float(np.linalg.norm(e.field - e1.field) / np.linalg.norm(e1.field))
            hist = lambda i: [float(f"{v:.3e}") for v in i['error_at_cycle'] / i['ref_error']]
            good = (info['exit_message'] == i1['exit_message'] == 'CONVERGED' and err < max(1e-6, 10 * float(kw.get('tol', 1e-6)))
                    and abs(info['it_mg'] - i1['it_mg']) <= (3 if kw.get('sslsolver') else 1)
                    and abs(info['it_ssl'] - i1['it_ssl']) <= (1 if kw.get('sslsolver') else 0))
            ok = ok and good
            print(json.dumps({
                'case': name, 'shape': [int(v) for v in grid.shape_cells], 'nranks': world, 'kw': kw,
                'ok': bool(good), 'efield_rel_diff': err, 'warm_dist': warm, 'warm_single': warm1,
                'dist': {'it_mg': info['it_mg'], 'it_ssl': info['it_ssl'], 'exit': info['exit_message'],
                         'wall_s': round(dt, 3), 'err_hist': hist(info)},
                'single': {'it_mg': i1['it_mg'], 'it_ssl': i1['it_ssl'], 'exit': i1['exit_message'],
                           'wall_s': round(dt1, 3), 'err_hist': hist(i1)}}), flush=True)
        dist.barrier()
    comm.destroy()
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == '__main__':
    main()

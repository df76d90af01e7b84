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
    "bicgstab:config3:128:sslsolver=bicgstab,cycle=V,semicoarsening=False,linerelaxation=False",
    "cgs:config2:128:sslsolver=cgs,cycle=V,semicoarsening=False,linerelaxation=False",
]


def parse(case):
    name, config, n, kws = case.split(':', 3)
    kw = {}
    for item in filter(None, kws.split(',')):
        k, v = item.split('=')
        kw[k] = {'True': True, 'False': False}.get(v, int(v) if v.lstrip('-').isdigit() else v)
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
        grid = eb.TensorMesh(cfg['h'], cfg['origin'])
        model = eb.Model(grid, **cfg['model'])
        sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
        kw = dict(dict(tol=1e-8, maxit=50), **kw)
        _lib.sync()
        dist.barrier()
        t0 = time.perf_counter()
        try:
            e, info = eb.solve(model, sfield, comm=comm, return_info=True, **kw)
        except Exception as err:                       # noqa: BLE001
            if rank == 0:
                print(json.dumps({'case': name, 'error': repr(err)}), flush=True)
            ok = False
            dist.barrier()
            continue
        _lib.sync()
        dt = time.perf_counter() - t0
        if rank == 0:
            t0 = time.perf_counter()
            e1, i1 = eb.solve(model, sfield, return_info=True, **kw)
            _lib.sync()
            dt1 = time.perf_counter() - t0
            err = float(np.linalg.norm(e.field - e1.field) / np.linalg.norm(e1.field))
            hist = lambda i: [float(f"{v:.3e}") for v in i['error_at_cycle'] / i['ref_error']]
            good = (info['exit_message'] == i1['exit_message'] == 'CONVERGED' and err < 1e-6
                    and abs(info['it_mg'] - i1['it_mg']) <= 1 and info['it_ssl'] == i1['it_ssl'])
            ok = ok and good
            print(json.dumps({
                'case': name, 'shape': [int(v) for v in grid.shape_cells], 'nranks': world, 'kw': kw,
                'ok': bool(good), 'efield_rel_diff': err,
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

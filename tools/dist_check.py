"""Distributed (z-slab) multigrid against the single-GPU solver.  Launch with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py [n] [cycle]

Every rank runs its slab; rank 0 additionally solves the whole problem on its GPU
with the single-GPU driver and compares fields, cycle counts and residual norms.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, parallel, recipes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cycle = sys.argv[2] if len(sys.argv) > 2 else 'V'
    lr = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    dist.init_process_group('gloo')
    _lib.init(local_rank)

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = parallel.NcclComm(rank, world, bcast)
    cfg = recipes.config('config5', n)                       # stretched grid, triaxial
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    model = eb.Model(grid, **cfg['model'])
    sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])

    dmg = parallel.DistributedMultigrid(model, sfield, comm)
    _lib.sync()
    dist.barrier()
    t0 = time.perf_counter()
    info = dmg.solve(cycle=cycle, tol=1e-9, maxit=40, linerelaxation=lr)
    _lib.sync()
    dt = time.perf_counter() - t0
    out = np.zeros(grid.n_edges, dtype=complex)
    dmg.download_owned(out)
    t = torch.from_numpy(out.view(np.float64))
    dist.all_reduce(t)                                        # disjoint owned parts: sum = gather

    if rank == 0:
        t0 = time.perf_counter()
        e1, i1 = eb.solve(model, sfield, plain=True, cycle=cycle, tol=1e-9, maxit=40,
                          linerelaxation=lr, return_info=True)
        _lib.sync()
        dt1 = time.perf_counter() - t0
        err = np.linalg.norm(out - e1.field) / np.linalg.norm(e1.field)
        print(json.dumps({
            'shape': grid.shape_cells, 'nranks': world, 'n_dist': dmg.n_dist, 'cycle': cycle, 'lr': lr,
            'dist': {'it_mg': info['it_mg'], 'rel_error': info['rel_error'],
                     'exit': info['exit_message'], 'wall_s': round(dt, 3),
                     'err_hist': [float(f"{v:.3e}") for v in info['error_at_cycle'] / info['ref_error']]},
            'single': {'it_mg': i1['it_mg'], 'rel_error': i1['rel_error'], 'exit': i1['exit_message'],
                       'wall_s': round(dt1, 3),
                       'err_hist': [float(f"{v:.3e}") for v in i1['error_at_cycle'] / i1['ref_error']]},
            'efield_rel_diff': err, 'ref_error_diff': abs(info['ref_error'] - i1['ref_error']) / i1['ref_error'],
        }), flush=True)
        assert info['exit_message'] == 'CONVERGED' and err < 1e-6, err
    comm.destroy()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

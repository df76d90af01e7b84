"""GCROT(m,k) around the distributed cycle against the single-GPU solve -- without torch.

    python tools/dist_gcrot_check.py [nranks] [n]

The launcher is the standard library: one spawned process per GPU, the NCCL id handed over through
queues (``parallel.NcclComm`` only asks for a broadcast callable), a ``multiprocessing.Barrier``
between the steps.  Every rank runs its z-slab; rank 0 then solves the whole problem on its GPU
and compares iteration counts, exit message and fields.  The source is scaled to norm one (the
reference's GCROT + multigrid needs that, see tests/test_gpu_solver_gcrot.py).
"""
import json
import multiprocessing as mp
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    ('gcrot-plain', dict(sslsolver='gcrotmk', semicoarsening=False, linerelaxation=False, cycle='F')),
    ('gcrot-sc-lr', dict(sslsolver='gcrotmk', semicoarsening=True, linerelaxation=True, cycle='V')),
    ('gcrot-noprec', dict(sslsolver='gcrotmk', semicoarsening=False, linerelaxation=False, cycle=None, maxit=3)),
]


def worker(rank, world, n, queues, barrier):
    try:
        import emg3d_b200 as eb
        from emg3d_b200 import _lib, parallel, recipes
        _lib.init(rank)

        def bcast(obj):
            if rank == 0:
                for q in queues[1:]:
                    q.put(obj)
                return obj
            return queues[rank].get(timeout=60)

        comm = parallel.NcclComm(rank, world, bcast)
        cfg = recipes.config('config2', n)
        grid = eb.TensorMesh(cfg['h'], cfg['origin'])
        model = eb.Model(grid, **cfg['model'])
        src = list(cfg['source'])
        src[2] -= 130.0                                   # off the slab interface
        dense = np.asarray(eb.get_source_field(grid, tuple(src), cfg['frequency']).field)
        sfield = eb.Field(grid, dense / np.linalg.norm(dense), frequency=cfg['frequency'])
        for name, kw in CASES:
            kw = dict(kw, tol=1e-8)
            barrier.wait(60)
            t0 = time.perf_counter()
            e, info = eb.solve(model, sfield, comm=comm, return_info=True, **kw)
            _lib.sync()
            dt = time.perf_counter() - t0
            barrier.wait(60)
            if rank == 0:
                e1, info1 = eb.solve(model, sfield, return_info=True, order='color', **kw)
                print(json.dumps({
                    'case': name, 'shape': [int(v) for v in grid.shape_cells], 'nranks': world,
                    'dist': [info['exit_message'], info['it_ssl'], info['it_mg'], float(info['rel_error'])],
                    'single': [info1['exit_message'], info1['it_ssl'], info1['it_mg'], float(info1['rel_error'])],
                    'field_rel_diff': float(np.linalg.norm(e.field - e1.field) / np.linalg.norm(e1.field)),
                    'wall_s': round(dt, 3)}), flush=True)
            barrier.wait(60)
        comm.destroy()
    except Exception:                                     # noqa: BLE001
        print(f"rank {rank}:\n{traceback.format_exc()}", flush=True)
        try:
            barrier.abort()
        except Exception:                                 # noqa: BLE001
            pass
        raise


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    ctx = mp.get_context('spawn')
    queues = [ctx.Queue() for _ in range(world)]
    barrier = ctx.Barrier(world)
    procs = [ctx.Process(target=worker, args=(r, world, n, queues, barrier)) for r in range(world)]
    for p in procs:
        p.start()
    deadline = time.time() + float(os.environ.get('DIST_GCROT_TIMEOUT', '80'))
    for p in procs:
        p.join(max(0.0, deadline - time.time()))
    bad = False
    for p in procs:
        if p.is_alive():                                  # (our own children, by handle)
            p.terminate()
            bad = True
        elif p.exitcode != 0:
            bad = True
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()

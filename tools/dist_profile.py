"""Where the time of a distributed V-cycle goes (run under torchrun, see dist_check.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/dist_profile.py [n_per_gpu] [n_dist ...]

Prints, per requested number of distributed levels: the cycle time, the cycle time
with the halo exchanges switched off (wrong numerics; compute + launch overhead
only) and the cost of one exchange on every distributed level.
"""
import json
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, parallel, recipes, solver  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ndists = [int(a) for a in sys.argv[2:]] or [None]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    dist.init_process_group('gloo')
    _lib.init(local_rank)

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = parallel.NcclComm(rank, world, bcast)
    cfg = recipes.bench_grid(world, n)
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    model = eb.Model(grid, **cfg['model'])
    sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
    shape = tuple(int(v) for v in grid.shape_cells)
    kw = dict(verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
              shape_cells=shape, cycle='V', maxit=1)

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        _lib.sync()
        dist.barrier()
        a, b = _lib.Event(), _lib.Event()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        _lib.sync()
        ms = a.elapsed_ms(b) / reps
        t = [None] * world
        dist.all_gather_object(t, ms)
        return max(t)

    for nd in ndists:
        dmg = parallel.DistributedMultigrid(model, sfield, comm, n_dist=nd)

        def step():
            var = solver.MGParameters(**kw)
            var.l2_refe = 1.0
            dmg.e.zero()
            dmg.multigrid(var)

        out = {"n_dist": dmg.n_dist, "shape": shape, "cycle_ms": timed(step)}
        real_exchange = dmg.exchange
        dmg.exchange = lambda dl, f: None
        out["cycle_ms_no_exchange"] = timed(step)
        dmg.exchange = real_exchange
        out["exchange_ms"] = [timed(lambda dl=dl: dmg.exchange(dl, dl.lv.e if dl.index else dmg.e), 10)
                              for dl in dmg.levels[:dmg.n_dist]]
        out["local_shapes"] = [tuple(dl.lv.shape) for dl in dmg.levels]
        # pieces of the level-0 visit
        d0, d1 = dmg.levels[0], dmg.levels[1]
        var = solver.MGParameters(**kw)
        var.l2_refe = 1.0
        lib = _lib.load()

        def res_restrict():
            r = dmg.residual(d0, dmg.s, dmg.e)
            _lib.check(lib.emg3d_b200_restrict(d1.lv.handle.ptr, r.ptr, d1.lv.s.ptr))
            d1.lv.e.zero()

        def prolong():
            _lib.check(lib.emg3d_b200_prolong(d1.lv.handle.ptr, dmg.e.ptr, d1.lv.e.ptr))
            dmg.exchange(d0, dmg.e)

        out["pieces_ms"] = {
            "smoothing_nu2": timed(lambda: dmg.smoothing(d0, dmg.s, dmg.e, 2, 0)),
            "residual_norm": timed(lambda: dmg.residual(d0, dmg.s, dmg.e, norm=True)),
            "residual_restrict": timed(res_restrict),
            "descend": timed(lambda: dmg._descend(var, d1, 1, var.cycmax)),
            "prolong": timed(prolong),
        }
        if rank == 0:
            print(json.dumps(out), flush=True)
        del dmg
    comm.destroy()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

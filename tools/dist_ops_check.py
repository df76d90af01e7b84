"""Building blocks of the distributed Krylov wrappers against their single-GPU counterparts.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/dist_ops_check.py [n]

dot / norm over owned edges, the operator (halo refresh + window matvec) and the multigrid
preconditioner (one V-cycle from zero): value against the single-GPU result, repeatability
(the same call twice) and linearity.
"""
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emg3d_b200 as eb  # noqa: E402
from emg3d_b200 import _lib, parallel, recipes, solver  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cname = sys.argv[2] if len(sys.argv) > 2 else 'config3'
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo')
    _lib.init(int(os.environ.get('LOCAL_RANK', rank)))

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = parallel.NcclComm(rank, world, bcast)
    cfg = recipes.config(cname, n)
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    model = eb.Model(grid, **cfg['model'])
    sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
    dmg = parallel.DistributedMultigrid(model, sfield, comm)
    ops = dmg._Ops(dmg)
    rng = np.random.default_rng(5)
    ne = int(grid.n_edges)

    def rand_field():
        f = eb.Field(grid, dtype=complex, frequency=cfg['frequency'])
        f.field[:] = rng.standard_normal(ne) + 1j * rng.standard_normal(ne)
        return f

    def up(f):
        d = dmg.level0.lv.new_field()
        dmg.upload_field(f, dst=d)
        return d

    def gather(d):
        out = np.zeros(ne, dtype=complex)
        dmg.download_owned(out, src=d)
        t = __import__('torch').from_numpy(out.view(float))
        dist.all_reduce(t)
        return out

    def report(name, a, b):
        if rank == 0:
            print(f"{name:<44s} {np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300):.3e}", flush=True)

    def pec(f):
        """zero the tangential boundary edges (what every vector of a solve satisfies)"""
        vx, vy, vz = grid.shape_edges_x, grid.shape_edges_y, grid.shape_edges_z
        nx_, ny_ = int(np.prod(vx)), int(np.prod(vy))
        fx = f.field[:nx_].reshape(vx, order='F')
        fy = f.field[nx_:nx_ + ny_].reshape(vy, order='F')
        fz = f.field[nx_ + ny_:].reshape(vz, order='F')
        fx[:, [0, -1], :] = 0; fx[:, :, [0, -1]] = 0
        fy[[0, -1], :, :] = 0; fy[:, :, [0, -1]] = 0
        fz[[0, -1], :, :] = 0; fz[:, [0, -1], :] = 0
        return f

    x, y = pec(rand_field()), pec(rand_field())
    dx, dy = up(x), up(y)
    # --- dot / norm
    report("norm(x)", np.array([ops.norm(dx)]), np.array([np.linalg.norm(x.field)]))
    report("dot(x, y)", np.array([ops.dot(dx, dy)]), np.array([np.vdot(x.field, y.field)]))
    # --- operator
    dv = ops.new()
    ops.matvec(dx, dv)
    v1 = gather(dv)
    ops.matvec(dx, dv)
    v2 = gather(dv)
    report("matvec repeat", v1, v2)
    lv = solver._Level.from_volume_model(eb.VolumeModel(model, sfield), np.dtype(complex)) if rank == 0 else None
    if rank == 0:
        sx = _lib.DeviceArray.from_host(np.asarray(x.field))
        sv = lv.new_field()
        _lib.check(_lib.load().emg3d_b200_apply(lv.handle.ptr, sx.ptr, sv.ptr))
        report("matvec vs single GPU", v1, sv.download())
    # --- preconditioner
    def var_():
        v = solver.MGParameters(verb=-1, sslsolver='bicgstab', semicoarsening=False, linerelaxation=False,
                                shape_cells=dmg.gshape, cycle='V', return_info=True)
        v.order = 'color'
        v.l2_refe = 1e3 * float(np.linalg.norm(x.field))
        return v
    var = var_()
    dp = ops.new()
    ops.psolve(dx, dp, var)
    p1 = gather(dp)
    ops.psolve(dx, dp, var)
    p2 = gather(dp)
    report("psolve repeat (same var)", p2, p1)
    ops.psolve(dx, dp, var_())
    report("psolve repeat (fresh var)", gather(dp), p1)
    ops.psolve(dy, dp, var)
    py = gather(dp)
    z = eb.Field(grid, dtype=complex, frequency=cfg['frequency'])
    z.field[:] = 0.3 * x.field - (0.2 + 0.4j) * y.field
    dz = up(z)
    ops.psolve(dz, dp, var)
    report("psolve linearity", gather(dp), 0.3 * p1 - (0.2 + 0.4j) * py)
    # vectors whose halo planes are garbage (what axpby leaves behind)
    junk = dmg.level0.lv.new_field()
    junk.upload(np.full(dmg.level0.lv.n_edges, 1e30 + 1e30j))
    copies, _, _ = parallel.gather_plan(dmg.part0, 0, dmg.rank, dmg.gshape[0], dmg.gshape[1])
    isz = 16
    for loff, goff, cnt in copies:
        _lib.check(_lib.load().emg3d_b200_d2d(junk.ptr + loff * isz, dx.ptr + loff * isz, cnt * isz))
    ops.psolve(junk, dp, var)
    report("psolve with garbage outside the owned part", gather(dp), p1)
    ops.matvec(junk, dv)
    report("matvec with garbage outside the owned part", gather(dv), v1)
    report("norm with garbage outside", np.array([ops.norm(junk)]), np.array([np.linalg.norm(x.field)]))
    if rank == 0:
        var1 = var_()
        sp = lv.new_field()
        sp.zero()
        var1.e_is_zero, var1.s_norm = True, None
        solver._multigrid(lv, sx, sp, var1)
        sp1 = sp.download()
        report("psolve vs single GPU (other colour order)", p1, sp1)

        def single(f):
            d_in = _lib.DeviceArray.from_host(np.asarray(f.field))
            v_ = var_()
            sp.zero()
            v_.e_is_zero, v_.s_norm = True, None
            solver._multigrid(lv, d_in, sp, v_)
            return sp.download()
        report("single-GPU psolve linearity", single(z), 0.3 * sp1 - (0.2 + 0.4j) * single(y))
    dist.barrier()
    dmg.close()
    comm.destroy()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""CPU tests of the z-slab decomposition logic (emg3d_b200/parallel.py).

The index arithmetic (ownership, local grids, halo-exchange and gather plans) is
independent of the transport.  It is checked (i) in one process with a virtual
transport that pairs every send with the peer's receive, for several rank counts
and levels, and (ii) with real point-to-point messages between two ``gloo`` ranks.
"""
import os
import socket

import numpy as np
import pytest

from emg3d_b200 import parallel


def field_sizes(nx, ny, nzc):
    return nx * (ny + 1) * (nzc + 1) + (nx + 1) * ny * (nzc + 1) + (nx + 1) * (ny + 1) * nzc


def global_field(nx, ny, nz):
    """A field whose value encodes (component, ix, iy, iz): easy to verify."""
    out = []
    for c, shp in enumerate(((nx, ny + 1, nz + 1), (nx + 1, ny, nz + 1), (nx + 1, ny + 1, nz))):
        i, j, k = np.meshgrid(*[np.arange(n) for n in shp], indexing='ij')
        out.append((1e6 * (c + 1) + 1e4 * k + 1e2 * j + i + 0.5j * (k + 1)).ravel('F'))
    return np.concatenate(out)


def local_from_global(part, level, rank, nx, ny, gfield, owned_only):
    """Rank-local field: owned part from the global field, NaN elsewhere (or the whole slab)."""
    lo, hi = part.local(level, rank)
    loc = np.full(field_sizes(nx, ny, hi - lo), np.nan + 0j)
    if owned_only:
        nzl = part.nz_level(level)
        copies, _, _ = parallel.gather_plan(part, level, rank, nx, ny)
        for loff, goff, n in copies:
            loc[loff:loff + n] = gfield[goff:goff + n]
    else:
        for goff, loff, n in parallel.scatter_ranges(part, level, rank, nx, ny):
            loc[loff:loff + n] = gfield[goff:goff + n]
    return loc


@pytest.mark.parametrize('nz,nranks,n_dist', [(64, 2, 3), (64, 4, 2), (256, 8, 3), (96, 3, 2),
                                               (32, 2, 1), (512, 8, 4)])
def test_partition_is_consistent(nz, nranks, n_dist):
    part = parallel.SlabPartition(nz, nranks, n_dist)
    for level in range(n_dist + 1):
        b = part.bounds(level)
        nzl = part.nz_level(level)
        assert b[0] == 0 and b[-1] == nzl + 1 and all(x < y for x, y in zip(b[:-1], b[1:]))
        for r in range(nranks):
            lo, hi = part.local(level, r)
            p0, p1 = part.owned(level, r)
            assert 0 <= lo <= p0 and p1 - 1 <= hi <= nzl
            if level < n_dist:
                # local grids coarsen onto the next level's local grids
                assert lo % 2 == 0 and (hi - lo) % 2 == 0
                clo, chi = part.local(level + 1, r)
                assert (clo, chi) == (lo // 2, hi // 2)
                # ownership coarsens consistently: even owned planes <-> coarse owned planes
                c0, c1 = part.owned(level + 1, r)
                assert c0 == (p0 + 1) // 2 and c1 - 1 == (p1 - 1) // 2
            # halo: one plane above, >= 1 below
            if r > 0:
                assert p0 - lo == 2 ** (n_dist - level)
            if r < nranks - 1:
                assert hi == p1


def test_partition_rejects_bad_input():
    with pytest.raises(ValueError):
        parallel.SlabPartition(60, 2, 3)        # not a multiple of 8
    with pytest.raises(ValueError):
        parallel.SlabPartition(16, 4, 3)        # 2 aligned blocks for 4 ranks


@pytest.mark.parametrize('nz,nranks,n_dist', [(64, 2, 3), (64, 4, 2), (96, 3, 2)])
def test_exchange_and_gather_plans_virtual(nz, nranks, n_dist):
    """All ranks in one process; sends are paired with the peers' receives in order."""
    nx, ny = 6, 4
    part = parallel.SlabPartition(nz, nranks, n_dist)
    for level in range(n_dist + 1):
        sx, sy = nx, ny                                   # x, y extents do not matter here
        nzl = part.nz_level(level)
        g = global_field(sx, sy, nzl)
        locs = [local_from_global(part, level, r, sx, sy, g, owned_only=True) for r in range(nranks)]
        plans = [parallel.exchange_plan(part, level, r, sx, sy) for r in range(nranks)]
        # pair sends and receives per (src, dst), in order
        for src in range(nranks):
            for dst in range(nranks):
                sends = [(o, n) for s, p, o, n in plans[src] if s and p == dst]
                recvs = [(o, n) for s, p, o, n in plans[dst] if not s and p == src]
                assert [n for _, n in sends] == [n for _, n in recvs]
                for (so, n), (ro, _) in zip(sends, recvs):
                    locs[dst][ro:ro + n] = locs[src][so:so + n]
        for r in range(nranks):
            want = local_from_global(part, level, r, sx, sy, g, owned_only=False)
            lo, hi = part.local(level, r)
            p0, p1 = part.owned(level, r)
            px, py, pz = sx * (sy + 1), (sx + 1) * sy, (sx + 1) * (sy + 1)
            ox, oy, oz = 0, px * (hi - lo + 1), px * (hi - lo + 1) + py * (hi - lo + 1)
            # planes the smoother / residual / transfer operators read around the owned block
            for p in range(max(p0 - 1, lo), min(p1, hi) + 1):
                for off, sz in ((ox, px), (oy, py)):
                    a = off + sz * (p - lo)
                    np.testing.assert_array_equal(locs[r][a:a + sz], want[a:a + sz])
            first_layer = max(p0 - 2, lo) if level < n_dist else max(p0 - 1, lo)
            for k in range(first_layer, min(p1, hi)):
                a = oz + pz * (k - lo)
                np.testing.assert_array_equal(locs[r][a:a + pz], want[a:a + pz])
        # gather: every global element is owned by exactly one rank
        out = np.full(g.size, np.nan + 0j)
        cover = np.zeros(g.size, dtype=int)
        for r in range(nranks):
            copies, sends, recvs = parallel.gather_plan(part, level, r, sx, sy)
            for loff, goff, n in copies:
                out[goff:goff + n] = locs[r][loff:loff + n]
                cover[goff:goff + n] += 1
            # what r sends to q is what q expects from r
            for q in range(nranks):
                if q == r:
                    continue
                _, _, qrecvs = parallel.gather_plan(part, level, q, sx, sy)
                assert [n for p, _, n in sends if p == q] == [n for p, _, n in qrecvs if p == r]
        assert np.all(cover == 1)
        np.testing.assert_array_equal(out, g)


@pytest.mark.parametrize('nz,nranks,n_dist', [(64, 2, 3), (64, 4, 2), (96, 3, 2)])
def test_pull_plan_equals_send_recv_pairs(nz, nranks, n_dist):
    """The peer-memory exchange (one kernel pulling from the neighbours) moves exactly
    what the NCCL send/recv plan moves."""
    nx, ny = 5, 3
    part = parallel.SlabPartition(nz, nranks, n_dist)
    for level in range(n_dist):
        g = global_field(nx, ny, part.nz_level(level))
        base = [local_from_global(part, level, r, nx, ny, g, owned_only=True) for r in range(nranks)]
        a = [b.copy() for b in base]
        plans = [parallel.exchange_plan(part, level, r, nx, ny) for r in range(nranks)]
        for src in range(nranks):
            for dst in range(nranks):
                sends = [(o, n) for s, p, o, n in plans[src] if s and p == dst]
                recvs = [(o, n) for s, p, o, n in plans[dst] if not s and p == src]
                for (so, n), (ro, _) in zip(sends, recvs):
                    a[dst][ro:ro + n] = base[src][so:so + n]
        b = [x.copy() for x in base]
        for r in range(nranks):
            pulls = parallel.pull_plan(part, level, r, nx, ny)
            assert len(pulls) <= 8                        # P2P_MAX_SEG of csrc/comm.cu
            for from_upper, mo, po, n in pulls:
                q = r + 1 if from_upper else r - 1
                b[r][mo:mo + n] = base[q][po:po + n]
        for r in range(nranks):
            np.testing.assert_array_equal(a[r], b[r])


@pytest.mark.parametrize('nz,nranks,n_dist', [(64, 2, 3), (64, 4, 2), (96, 3, 2)])
def test_owned_plane_window_equals_owned_ranges(nz, nranks, n_dist):
    """The residual kernel sums |r|^2 over the planes [own0, own1) of the rank's z-window
    (emg3d_b200_level_set_owned: x/y-edges on those node planes, z-edges of the layers whose
    upper plane is among them).  That set must be exactly the element ranges of
    `owned_ranges`, which the unfused norm (sum_owned) uses, and the ranks' sets must
    partition the global field."""
    nx, ny = 5, 3
    part = parallel.SlabPartition(nz, nranks, n_dist)
    for level in range(n_dist):
        nzl = part.nz_level(level)
        total = 0
        for rank in range(nranks):
            lo, hi = part.local(level, rank)
            p0, p1 = part.owned(level, rank)
            z0 = 0 if rank == 0 else p0 - 1 - lo           # window start (parallel._DLevel)
            own0, own1 = p0 - lo - z0, p1 - lo - z0
            nzw = hi - lo - z0                              # cells of the window
            assert 0 <= own0 < own1 <= nzw + 1
            nloc = hi - lo
            px, py, pz = nx * (ny + 1), (nx + 1) * ny, (nx + 1) * (ny + 1)
            mask = np.zeros(px * (nloc + 1) + py * (nloc + 1) + pz * nloc, dtype=bool)
            ox, oy, oz = 0, px * (nloc + 1), px * (nloc + 1) + py * (nloc + 1)
            for k in range(nzw + 1):                        # window-local plane k = local plane z0 + k
                if own0 <= k < own1:
                    mask[ox + px * (z0 + k):ox + px * (z0 + k + 1)] = True
                    mask[oy + py * (z0 + k):oy + py * (z0 + k + 1)] = True
                if k < nzw and own0 <= k + 1 < own1:
                    mask[oz + pz * (z0 + k):oz + pz * (z0 + k + 1)] = True
            want = np.zeros_like(mask)
            for off, n in parallel.owned_ranges(part, level, rank, nx, ny):
                want[off:off + n] = True
            np.testing.assert_array_equal(mask, want)
            total += int(mask.sum())
        assert total == px * (nzl + 1) + py * (nzl + 1) + pz * nzl


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, nz, n_dist, nx, ny, results):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        part = parallel.SlabPartition(nz, world, n_dist)
        ok = True
        for level in range(n_dist + 1):
            nzl = part.nz_level(level)
            g = global_field(nx, ny, nzl)
            loc = local_from_global(part, level, rank, nx, ny, g, owned_only=True)
            view = torch.from_numpy(loc.view(np.float64))           # (re, im) pairs
            reqs = []
            for is_send, peer, off, n in parallel.exchange_plan(part, level, rank, nx, ny):
                t = view[2 * off:2 * (off + n)]
                reqs.append(dist.isend(t, peer) if is_send else dist.irecv(t, peer))
            for q in reqs:
                q.wait()
            want = local_from_global(part, level, rank, nx, ny, g, owned_only=False)
            lo, hi = part.local(level, rank)
            p0, p1 = part.owned(level, rank)
            px = nx * (ny + 1)
            for p in range(max(p0 - 1, lo), min(p1, hi) + 1):
                a = px * (p - lo)
                ok &= bool(np.array_equal(loc[a:a + px], want[a:a + px]))
            # all-gather into the global layout
            out = np.full(g.size, np.nan + 0j)
            oview = torch.from_numpy(out.view(np.float64))
            copies, sends, recvs = parallel.gather_plan(part, level, rank, nx, ny)
            for loff, goff, n in copies:
                out[goff:goff + n] = loc[loff:loff + n]
            reqs = [dist.isend(view[2 * o:2 * (o + n)], p) for p, o, n in sends]
            reqs += [dist.irecv(oview[2 * o:2 * (o + n)], p) for p, o, n in recvs]
            for q in reqs:
                q.wait()
            ok &= bool(np.array_equal(out, g))
            # owned sums add up to the global sum (norms)
            s = sum(float(np.sum(np.abs(loc[o:o + n]) ** 2))
                    for o, n in parallel.owned_ranges(part, level, rank, nx, ny))
            t = torch.tensor([s], dtype=torch.float64)
            dist.all_reduce(t)
            ok &= bool(abs(t.item() - float(np.sum(np.abs(g) ** 2))) <= 1e-12 * t.item())
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def test_exchange_and_gather_with_gloo_world_size_2():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_gloo_worker, args=(world, port, 32, 2, 5, 3, results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}


# ---- semicoarsening: per-level z-shifts and the planning of the distributed hierarchies --------

@pytest.mark.parametrize('nz,nranks,zshifts,zmax', [(64, 2, [0, 0, 1, 1], 3), (256, 4, [0, 1, 1, 2], 2),
                                                      (128, 8, [0, 0, 0], 1), (64, 4, [0, 1, 2], 2)])
def test_partition_with_z_shifts(nz, nranks, zshifts, zmax):
    """Levels that do not coarsen z keep plane ownership; local grids stay nested; level 0 depends
    on (nz, nranks, zmax) only, so hierarchies of several patterns share it."""
    n_dist = len(zshifts) - 1
    part = parallel.SlabPartition(nz, nranks, n_dist, zshifts, zmax)
    ref0 = parallel.SlabPartition(nz, nranks, 0, [0], zmax)
    for r in range(nranks):
        assert part.local(0, r) == ref0.local(0, r) and part.owned(0, r) == ref0.owned(0, r)
    for level in range(n_dist + 1):
        assert part.nz_level(level) == nz >> zshifts[level]
        b = part.bounds(level)
        assert b[0] == 0 and b[-1] == part.nz_level(level) + 1
        for r in range(nranks):
            lo, hi = part.local(level, r)
            p0, p1 = part.owned(level, r)
            assert 0 <= lo <= p0 and p1 - 1 <= hi <= part.nz_level(level)
            if r > 0:
                assert p0 - lo == part.depth(level) == 2 ** (zmax - zshifts[level])
            if level < n_dist:
                clo, chi = part.local(level + 1, r)
                if zshifts[level + 1] == zshifts[level]:
                    assert (clo, chi) == (lo, hi) and part.owned(level + 1, r) == (p0, p1)
                else:
                    assert lo % 2 == 0 and (hi - lo) % 2 == 0 and (clo, chi) == (lo // 2, hi // 2)
    # halo plans of neighbouring ranks still pair up on every level
    for level in range(n_dist):
        for r in range(nranks):
            parallel.pull_plan(part, level, r, 6, 5)
            parallel.pull_plan(part, level, r, 6, 5, shared_from_lower=True)


def test_hierarchy_plan_follows_the_single_gpu_coarsening():
    """Global shapes / transitions of the distributed levels equal what the single-GPU driver
    does level by level (solver._current_sc_dir), for every semicoarsening pattern."""
    from emg3d_b200 import core, solver
    shape = (512, 512, 256)
    var = solver.MGParameters(verb=-1, sslsolver=False, semicoarsening=True, linerelaxation=True,
                              shape_cells=shape, cycle='F')
    assert sorted(set(var.raw_sc_cycle)) == [1, 2, 3]
    for pat in (0, 1, 2, 3):
        shapes, trans, zs = parallel.hierarchy_plan(shape, pat, 4, int(var.clevel[pat]) - 1)
        assert shapes[0] == shape and len(shapes) == len(trans) + 1 == len(zs)
        cur = shape
        for l, sc in enumerate(trans):
            assert sc == int(solver._current_sc_dir(pat, parallel._Shape(cur)))
            flag = core.SC_FLAGS[sc]
            cur = tuple(n // 2 if f else n for n, f in zip(cur, flag))
            assert cur == shapes[l + 1] and zs[l + 1] == zs[l] + int(bool(flag[2]))
        # every distributed level (but the finest) has more than a million cells; the first
        # replicated one leaves two planes per rank
        assert all(int(np.prod(s)) > 1_000_000 for s in shapes[1:-1])
        assert shapes[-1][2] // 4 >= 2
        if pat == 3:
            assert zs[-1] == 0                      # z is never coarsened: slabs keep their planes
        if pat == 1:
            assert all(s[0] == 512 for s in shapes)
    with pytest.raises(ValueError):
        parallel.hierarchy_plan((2, 2, 2), 0, 2, 0)


def test_shared_layer_direction_of_the_exchange_plans():
    """The fz layer between the planes next to an interface travels downwards by default and
    upwards with shared_from_lower (the lower rank relaxed it last); everything else is equal."""
    part = parallel.SlabPartition(32, 2, 1)
    nx, ny = 4, 3
    up0 = parallel.exchange_plan(part, 0, 0, nx, ny)
    lo0 = parallel.exchange_plan(part, 0, 0, nx, ny, shared_from_lower=True)
    up1 = parallel.exchange_plan(part, 0, 1, nx, ny)
    lo1 = parallel.exchange_plan(part, 0, 1, nx, ny, shared_from_lower=True)
    pz = (nx + 1) * (ny + 1)
    only = lambda a, b: [x for x in a if x not in b]
    assert [(s, c) for s, _, _, c in only(up0, lo0)] == [(0, pz)]      # rank 0 receives the layer ...
    assert [(s, c) for s, _, _, c in only(lo0, up0)] == [(1, pz)]      # ... or sends it
    assert [(s, c) for s, _, _, c in only(up1, lo1)] == [(1, pz)]
    assert [(s, c) for s, _, _, c in only(lo1, up1)] == [(0, pz)]


def test_sparse_source_slab_equals_dense_slab():
    """A sparse source (indices, values) cut to a rank's slab places the same entries as slicing
    the dense field (halo planes included), for every rank of 1-, 2- and 4-rank partitions."""
    from emg3d_b200 import parallel
    nx, ny, nz = 6, 5, 16
    n_edges = nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(n_edges, 60, replace=False)).astype(np.int64)
    val = rng.standard_normal(60) + 1j * rng.standard_normal(60)
    dense = np.zeros(n_edges, dtype=complex)
    dense[idx] = val
    for nranks in (1, 2, 4):
        part = parallel.SlabPartition(nz, nranks, 0, [0], 2)
        for rank in range(nranks):
            ranges = parallel.scatter_ranges(part, 0, rank, nx, ny)
            nloc = sum(n for _, _, n in ranges)
            want = np.zeros(nloc, dtype=complex)
            for goff, loff, n in ranges:
                want[loff:loff + n] = dense[goff:goff + n]
            li, lv = parallel.slab_sparse(part, rank, nx, ny, idx, val)
            got = np.zeros(nloc, dtype=complex)
            got[li] = lv
            assert np.array_equal(got, want), (nranks, rank)
            assert len(np.unique(li)) == len(li)


@pytest.mark.parametrize('nranks,nbatch', [(1, 1), (2, 1), (2, 4), (4, 4), (8, 8), (8, 3)])
def test_zline_schedule_is_a_pipeline(nranks, nbatch):
    """Every rank takes the same number of steps (each is followed by a collective exchange); a
    batch is relaxed by rank r only after rank r - 1 (forward) resp. r + 1 (backward) relaxed it
    and exchanged; the backward phase of a batch starts after its forward phase on the top rank."""
    from emg3d_b200.parallel import zline_schedule
    sched = [zline_schedule(nranks, nbatch, r) for r in range(nranks)]
    assert len({len(s) for s in sched}) == 1 and len(sched[0]) == 2 * (nbatch + nranks - 1)
    when = {}
    for r, steps in enumerate(sched):
        for t, (phase, b) in enumerate(steps):
            if b is not None:
                assert (phase, b, r) not in when
                when[(phase, b, r)] = t
    for b in range(nbatch):
        for r in range(nranks):
            assert (1, b, r) in when and (2, b, r) in when
            if r > 0:
                assert when[(1, b, r)] > when[(1, b, r - 1)]
            if r < nranks - 1:
                assert when[(2, b, r)] > when[(2, b, r + 1)]
        assert when[(2, b, nranks - 1)] > when[(1, b, nranks - 1)]


def test_chained_line_solve_equals_the_global_solve():
    """The scheme of the z-lines cut by slabs (parallel.zline_smoothing, csrc/gs_line.cu) on a scalar
    tridiagonal model: pivots continued through the cuts, forward substitution rank after rank
    with the intermediate of the piece's last node handed up through the halo, backward substitution
    downwards with the solved value handed down -- equals the solve of the whole line."""
    rng = np.random.default_rng(3)
    n, cuts = 40, [0, 9, 23, 31, 40]                    # unknowns 0 .. n-1, four pieces
    lo = rng.uniform(-1, 1, n)
    up = np.r_[lo[1:], 0.0]                             # symmetric: up[i] = lo[i + 1]
    dg = 2.5 + np.abs(lo) + np.abs(up) + 1j * rng.uniform(0, 1, n)
    rhs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    A = np.diag(dg) + np.diag(lo[1:], -1) + np.diag(up[:-1], 1)
    want = np.linalg.solve(A, rhs)
    # factorisation chained through the cuts (level_line_chain): pivot of the piece's first unknown
    # continues from the lower piece's last pivot
    piv = np.zeros(n, dtype=complex)
    carry = None
    for p0, p1 in zip(cuts[:-1], cuts[1:]):
        for i in range(p0, p1):
            prev = carry if i == p0 else piv[i - 1]
            piv[i] = dg[i] - (0 if prev is None else lo[i] * up[i - 1] / prev)
        carry = piv[p1 - 1]
    # forward bottom-up: halo value = g of the lower piece's last unknown
    g = np.zeros(n, dtype=complex)
    halo = 0.0
    for p0, p1 in zip(cuts[:-1], cuts[1:]):
        prev = halo
        for i in range(p0, p1):
            g[i] = (rhs[i] - lo[i] * prev) / piv[i]
            prev = g[i]
        halo = g[p1 - 1]
    # backward top-down: halo value = solved value of the upper piece's first unknown
    x = np.zeros(n, dtype=complex)
    halo = 0.0
    for p0, p1 in reversed(list(zip(cuts[:-1], cuts[1:]))):
        nxt = halo
        for i in range(p1 - 1, p0 - 1, -1):
            x[i] = g[i] - up[i] * nxt / piv[i]
            nxt = x[i]
        halo = x[p0]
    assert np.linalg.norm(x - want) < 1e-13 * np.linalg.norm(want)


@pytest.mark.parametrize('nz,nranks,n_dist', [(16, 2, 1), (32, 4, 2), (64, 8, 2)])
@pytest.mark.parametrize('low', [False, True])
def test_push_plan_mirrors_the_pull_plans(nz, nranks, n_dist, low):
    """Executing every rank's push plan moves exactly what executing every rank's pull plan
    moves (virtual ranks, host arrays), for both directions of the shared fz layer."""
    from emg3d_b200 import parallel
    nx, ny = 5, 4
    part = parallel.SlabPartition(nz, nranks, n_dist)
    for level in range(n_dist):
        rng = np.random.default_rng(level)
        fields = []
        for r in range(nranks):
            lo, hi = part.local(level, r)
            (px, py, pz), (ox, oy, oz) = parallel._comp_sizes(nx >> level or 1, ny >> level or 1, hi - lo)
            fields.append(rng.standard_normal(oz + pz * (hi - lo)))
        nxl, nyl = nx >> level or 1, ny >> level or 1
        pulled = [f.copy() for f in fields]
        for r in range(nranks):
            for from_upper, mo, po, cnt in parallel.pull_plan(part, level, r, nxl, nyl, low):
                q = r + 1 if from_upper else r - 1
                pulled[r][mo:mo + cnt] = fields[q][po:po + cnt]
        pushed = [f.copy() for f in fields]
        for r in range(nranks):
            for to_upper, mo, po, cnt in parallel.push_plan(part, level, r, nxl, nyl, low):
                q = r + 1 if to_upper else r - 1
                pushed[q][po:po + cnt] = fields[r][mo:mo + cnt]
        for a, b in zip(pulled, pushed):
            assert np.array_equal(a, b)


def _split_local(loc, nx, ny, nzl):
    """F-ordered 3-D views (fx, fy, fz) of a local 1-D field with nzl cells along z."""
    shp = ((nx, ny + 1, nzl + 1), (nx + 1, ny, nzl + 1), (nx + 1, ny + 1, nzl))
    out, i0 = [], 0
    for s_ in shp:
        n = int(np.prod(s_))
        out.append(loc[i0:i0 + n].reshape(s_, order='F'))
        i0 += n
    return out


@pytest.mark.parametrize('ldir', [0, 1, 2])
@pytest.mark.parametrize('nranks', [2, 4])
def test_exact_halves_across_slabs_equal_the_single_sweep_on_the_oracle(nranks, ldir):
    """The exact halo variant (DistributedMultigrid.smoothing) emulated on the CPU with the real
    plans and the C oracle as the smoother (point, x-line and y-line relaxation): every rank relaxes
    the blocks of one global z-parity on its z-window, the halos are exchanged (the shared fz layer from the lower rank after the half
    that relaxed the lower rank's top plane), then the other parity -- the assembled field equals
    one multicolour sweep of the whole grid, for descending and ascending sweeps."""
    import oracle
    oracle.build()
    nx, ny, nz = 4, 3, 16
    rng = np.random.default_rng(9)
    hs = [rng.uniform(1, 2, n) for n in (nx, ny, nz)]
    cplx = lambda shape: np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    eta = [cplx((nx, ny, nz)) * 0.1 - 1j for _ in range(3)]
    zeta = np.asfortranarray(rng.uniform(1, 2, (nx, ny, nz)))
    ne = field_sizes(nx, ny, nz)
    e0 = rng.standard_normal(ne) + 1j * rng.standard_normal(ne)
    s = rng.standard_normal(ne) + 1j * rng.standard_normal(ne)
    ex, ey, ez = _split_local(e0, nx, ny, nz)              # PEC: tangential boundary edges are zero
    ex[:, [0, -1], :] = 0; ex[:, :, [0, -1]] = 0
    ey[[0, -1], :, :] = 0; ey[:, :, [0, -1]] = 0
    ez[[0, -1], :, :] = 0; ez[:, [0, -1], :] = 0
    part = parallel.SlabPartition(nz, nranks, 1)

    # point smoother: 8 node colours, bit 2 = z-parity; x- / y-lines: 4 colours of the transverse
    # node indices (y or x, and z), bit 1 = z-parity.  Blocks are (ix, iy, iz) resp. (t1, iz).
    ncls, zbit = (8, 2) if ldir == 0 else (4, 1)

    def classes(back, pz):
        return [c for c in (range(ncls - 1, -1, -1) if back else range(ncls)) if (c >> zbit) & 1 == pz]

    def nodes_of(c, planes):
        if ldir == 0:
            px, py, pz = c & 1, (c >> 1) & 1, (c >> 2) & 1
            return [(ix, iy, iz) for iz in planes if 1 <= iz < nz and (iz - 1) & 1 == pz
                    for iy in range(1 + py, ny, 2) for ix in range(1 + px, nx, 2)]
        pp, pz = c & 1, c >> 1
        n1 = ny if ldir == 1 else nx                       # transverse axis next to z
        return [(a, iz) for iz in planes if 1 <= iz < nz and (iz - 1) & 1 == pz for a in range(1 + pp, n1, 2)]

    for back in (True, False):
        # --- one sweep of the whole grid
        ref = e0.copy()
        seq = [n for pz in ((1, 0) if back else (0, 1)) for c in classes(back, pz) for n in nodes_of(c, range(nz + 1))]
        oracle.gs_sequence(ldir, *_split_local(ref, nx, ny, nz), *_split_local(s, nx, ny, nz), *eta, zeta, *hs,
                           np.array(seq, dtype=np.int32))
        # --- the same on slabs
        locs = [local_from_global(part, 0, r, nx, ny, e0, owned_only=False) for r in range(nranks)]
        srcs = [local_from_global(part, 0, r, nx, ny, s, owned_only=False) for r in range(nranks)]
        for pz in ((1, 0) if back else (0, 1)):
            for r in range(nranks):
                lo, hi = part.local(0, r)
                p0, p1 = part.owned(0, r)
                z0 = 0 if r == 0 else p0 - lo - 1           # the window: one halo plane below
                w0 = lo + z0                                # global plane of window plane 0
                win = [a[:, :, z0:] for a in _split_local(locs[r], nx, ny, hi - lo)]
                swin = [a[:, :, z0:] for a in _split_local(srcs[r], nx, ny, hi - lo)]
                sl = slice(w0, hi)
                rows = [(*n[:-1], n[-1] - w0) for c in classes(back, pz) for n in nodes_of(c, range(p0, p1))]
                if rows:
                    oracle.gs_sequence(ldir, *win, *swin, *[np.asfortranarray(a[:, :, sl]) for a in eta],
                                       np.asfortranarray(zeta[:, :, sl]), hs[0], hs[1], hs[2][sl],
                                       np.array(rows, dtype=np.int32))
            # boundaries of the ownership blocks are even planes: the plane below an interface is
            # odd, i.e. of class bit 0 -- after that half the shared layer travels upwards
            plans = [parallel.exchange_plan(part, 0, r, nx, ny, shared_from_lower=(pz == 0)) for r in range(nranks)]
            new = [a.copy() for a in locs]
            for src in range(nranks):
                for dst in range(nranks):
                    sends = [(o, n) for s_, p, o, n in plans[src] if s_ and p == dst]
                    recvs = [(o, n) for s_, p, o, n in plans[dst] if not s_ and p == src]
                    for (so, n), (ro, _) in zip(sends, recvs):
                        new[dst][ro:ro + n] = locs[src][so:so + n]
            locs = new
        out = np.full(ne, np.nan + 0j)
        for r in range(nranks):
            copies, _, _ = parallel.gather_plan(part, 0, r, nx, ny)
            for loff, goff, n in copies:
                out[goff:goff + n] = locs[r][loff:loff + n]
        assert not np.isnan(out).any()
        assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref), (nranks, back)
        assert np.linalg.norm(ref - e0) > 0.1 * np.linalg.norm(e0)      # the sweep did something

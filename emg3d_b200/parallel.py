"""Multi-GPU multigrid by z-slab decomposition.

New functionality relative to the reference (which only runs independent solves
in a process pool, emg3d/_multiprocessing.py:33-65): ONE solve is spread over N
GPUs, one process per GPU, through ``emg3d_b200.solve(..., comm=)`` (inside a multi-process
launch) or ``solve(..., n_gpus=)`` (ranks spawned) -- same arguments, log, info dict and return
conventions as on one GPU.

* **Partition.**  z is the slowest axis of every array, so a z-slab of a field or
  coefficient array is one contiguous range.  Node planes are owned in contiguous
  blocks whose boundaries are multiples of ``2**zmax`` (``zmax``: the largest number of
  z-coarsenings of any distributed hierarchy), so that ownership coarsens consistently.  Every
  rank works on a *local grid*: its owned planes, one halo plane above and, below, as many
  halo planes as keep the local grid aligned with its own coarsening; the unchanged
  single-GPU kernels run on z-windows of the local grids, window boundaries playing the role
  of the PEC boundary.
* **Hierarchies.**  One chain of distributed levels per semicoarsening pattern of the cycle
  (:func:`hierarchy_plan`: patterns 1 / 2 keep halving z, pattern 3 leaves z alone), sharing
  the finest level; a level is distributed while it has more than about a million cells.
* **Smoothing is a true Gauss-Seidel sweep across slabs** (``exact=True``, default): every
  multicolour sweep runs as its two z-halves with a halo exchange after each
  (:meth:`DistributedMultigrid.smoothing`); ``exact=False`` exchanges once per sweep (the
  relaxed variant ``north_star`` words: block-Jacobi across slabs).  Point smoother, x- and
  y-lines work on the slab; **z-lines are cut by the slabs** and solved exactly as pieces of
  the global lines (:meth:`DistributedMultigrid.zline_smoothing`).
* **Halo exchange** = one peer-memory kernel per exchange (CUDA IPC mapping of the
  neighbours' arrays, remote loads over NVLink, flag handshake; csrc/comm.cu), NCCL
  send/recv groups as the fallback.
* **Norms and dot products** are sums over owned edges, all-reduced with NCCL; BiCGSTAB, CGS
  and GCROT(m,k) run around the distributed cycle through the backend-neutral drivers of
  solver.py (tools/dist_check.py, tools/dist_gcrot_check.py: same iteration counts and fields
  as on one GPU).
* **Coarse levels are replicated**: below the distributed levels the restricted residual is
  all-gathered and every rank runs the remaining coarse sub-cycle redundantly with the
  single-GPU driver; each rank keeps its slab of the correction.

Multicolour order only (``order='color'``).

The index arithmetic (:class:`SlabPartition`, :func:`exchange_plan`,
:func:`gather_plan`, :func:`hierarchy_plan`) is pure Python and is tested on CPU with two
``gloo`` ranks; the transport is pluggable (:class:`NcclComm` on device pointers,
``tests/test_parallel_cpu.py`` plugs in a gloo transport on host arrays).
"""
import numpy as np

__all__ = ['SlabPartition', 'exchange_plan', 'pull_plan', 'gather_plan', 'owned_ranges', 'NcclComm',
           'DistributedMultigrid']


# =========================================================================== #
# Pure index arithmetic (CPU-testable)
# =========================================================================== #

class SlabPartition:
    """Ownership of node planes along z for ``nranks`` slabs and ``n_dist`` levels.

    ``nz``: global number of cells along z on the finest grid.  Levels ``0 .. n_dist - 1`` are
    distributed, level ``n_dist`` is the first replicated one (slabs exist there only to
    gather).  ``zshifts[l]`` = number of z-coarsenings between level 0 and level ``l`` (default
    ``l``: standard coarsening; with semicoarsening some transitions leave z alone), so level
    ``l`` has ``nz >> zshifts[l]`` cells along z.  ``zmax`` (default ``zshifts[n_dist]``):
    ownership boundaries on the finest grid are multiples of ``2**zmax`` and the local grid of
    level 0 carries ``2**zmax`` halo planes below -- hierarchies of several semicoarsening
    patterns share level 0 when they are built with a common ``zmax``.
    """

    def __init__(self, nz, nranks, n_dist, zshifts=None, zmax=None):
        self.nz, self.nranks, self.n_dist = int(nz), int(nranks), int(n_dist)
        self.zshifts = list(range(self.n_dist + 1)) if zshifts is None else [int(z) for z in zshifts]
        if len(self.zshifts) != self.n_dist + 1 or self.zshifts[0] != 0 or any(
                b - a not in (0, 1) for a, b in zip(self.zshifts[:-1], self.zshifts[1:])):
            raise ValueError(f"zshifts={zshifts} must start at 0 and grow by 0 or 1 per level")
        self.zmax = self.zshifts[-1] if zmax is None else int(zmax)
        if self.zmax < self.zshifts[-1]:
            raise ValueError("zmax is smaller than the number of z-coarsenings")
        align = 1 << self.zmax
        if self.nz % align:
            raise ValueError(f"nz={nz} must be a multiple of 2**{self.zmax}={align}")
        nblocks = self.nz // align
        if nblocks < self.nranks:
            raise ValueError(f"nz={nz} gives {nblocks} aligned blocks for {nranks} ranks; "
                             "use fewer distributed levels or fewer ranks")
        # inner ownership boundaries on the finest grid: multiples of `align`
        self.bounds0 = [0] + [align * ((k * nblocks) // self.nranks)
                              for k in range(1, self.nranks)] + [self.nz + 1]
        for a, b in zip(self.bounds0[:-1], self.bounds0[1:]):
            if b - a < align:
                raise ValueError("a rank would own no plane on the coarsest distributed level")

    def nz_level(self, level):
        return self.nz >> self.zshifts[level]

    def depth(self, level):
        """Halo planes below the owned planes of a rank's local grid (ranks > 0)."""
        return 1 << (self.zmax - self.zshifts[level])

    def bounds(self, level):
        """Owned plane ranges [B[r], B[r+1]) on ``level`` (plane 0 .. nz_level)."""
        inner = [b >> self.zshifts[level] for b in self.bounds0[1:-1]]
        return [0] + inner + [self.nz_level(level) + 1]

    def owned(self, level, rank):
        b = self.bounds(level)
        return b[rank], b[rank + 1]

    def local(self, level, rank):
        """First and last node plane (inclusive) of the rank's local grid."""
        b = self.bounds(level)
        lo = 0 if rank == 0 else b[rank] - self.depth(level)
        hi = self.nz_level(level) if rank == self.nranks - 1 else b[rank + 1]
        return lo, hi


def _comp_sizes(nx, ny, nplanes_cells):
    """Elements per z-plane / z-layer and component offsets of a local field.

    Local grid with ``nplanes_cells`` cells along z.  Returns (plane sizes of
    fx, fy, fz; offsets of fx, fy, fz in the [fx | fy | fz] array).
    """
    px, py, pz = nx * (ny + 1), (nx + 1) * ny, (nx + 1) * (ny + 1)
    ox = 0
    oy = ox + px * (nplanes_cells + 1)
    oz = oy + py * (nplanes_cells + 1)
    return (px, py, pz), (ox, oy, oz)


def owned_ranges(part, level, rank, nx, ny):
    """Element ranges [(offset, count)] x 3 of the owned part of a local field.

    fx, fy on owned node planes; fz on the layers below owned planes (a layer
    belongs to the owner of its upper plane); the global boundary planes belong to
    the first / last rank, so that norms include them like the reference does.
    """
    lo, hi = part.local(level, rank)
    p0, p1 = part.owned(level, rank)
    (px, py, pz), (ox, oy, oz) = _comp_sizes(nx, ny, hi - lo)
    l0, l1 = max(p0 - 1, 0), p1 - 1                  # owned fz layers [l0, l1)
    return [(ox + px * (p0 - lo), px * (p1 - p0)),
            (oy + py * (p0 - lo), py * (p1 - p0)),
            (oz + pz * (l0 - lo), pz * (l1 - l0))]


def exchange_plan(part, level, rank, nx, ny, shared_from_lower=False):
    """Halo exchange of one field on one level.

    Returns a list of ``(is_send, peer, offset, count)`` in elements of the local
    field.  Upwards a rank sends fx, fy of its top owned plane and the fz layer
    below it; downwards fx, fy of its bottom owned plane and the fz layer below
    that plane.  (Two layers of fz are needed below an interface by the
    restriction of fz, one by the smoother.)

    The fz layer between the two planes next to an interface is SHARED: both ranks hold it and
    the smoothers of both relax it (an interior edge belongs to the blocks of both its end
    nodes).  By default the upper rank's copy wins (a layer belongs to the owner of its upper
    plane: residuals, prolongation, norms).  ``shared_from_lower=True`` sends the lower rank's
    copy up instead: used after the half sweep that relaxed the lower rank's top plane, whose
    update of the shared layer would otherwise be discarded.
    """
    lo, hi = part.local(level, rank)
    p0, p1 = part.owned(level, rank)
    (px, py, pz), (ox, oy, oz) = _comp_sizes(nx, ny, hi - lo)
    plan = []

    def plane(is_send, peer, p):
        plan.append((is_send, peer, ox + px * (p - lo), px))
        plan.append((is_send, peer, oy + py * (p - lo), py))

    def layer(is_send, peer, k):
        if k - lo >= 0 and k <= hi - 1:
            plan.append((is_send, peer, oz + pz * (k - lo), pz))

    if rank > 0:                                      # interface below: plane p0
        depth_mine = p0 - lo
        plane(1, rank - 1, p0)
        if not shared_from_lower:
            layer(1, rank - 1, p0 - 1)
        plane(0, rank - 1, p0 - 1)
        if shared_from_lower:
            layer(0, rank - 1, p0 - 1)
        if depth_mine >= 2:                           # the layer exists in my local grid
            layer(0, rank - 1, p0 - 2)
    if rank < part.nranks - 1:                        # interface above: plane p1
        depth_up = part.depth(level)                  # halo depth of the upper neighbour
        plane(1, rank + 1, p1 - 1)
        if shared_from_lower:
            layer(1, rank + 1, p1 - 1)
        if depth_up >= 2:
            layer(1, rank + 1, p1 - 2)
        plane(0, rank + 1, p1)
        if not shared_from_lower:
            layer(0, rank + 1, p1 - 1)
    return plan


def pull_plan(part, level, rank, nx, ny, shared_from_lower=False):
    """The halo exchange of :func:`exchange_plan` seen from the receiving side.

    Returns ``[(from_upper, my_offset, peer_offset, count)]``: the ``count`` elements
    at ``peer_offset`` of the neighbour's local field (``rank + 1`` if ``from_upper``
    else ``rank - 1``) belong at ``my_offset`` of this rank's.  A rank's receives
    from a neighbour are matched, in order, with that neighbour's sends to it.
    """
    mine = exchange_plan(part, level, rank, nx, ny, shared_from_lower)
    out = []
    for q in (rank - 1, rank + 1):
        if q < 0 or q >= part.nranks:
            continue
        recvs = [(off, cnt) for s, p, off, cnt in mine if not s and p == q]
        sends = [(off, cnt) for s, p, off, cnt in exchange_plan(part, level, q, nx, ny, shared_from_lower)
                 if s and p == rank]
        if [c for _, c in recvs] != [c for _, c in sends]:
            raise AssertionError("halo plans of neighbouring ranks do not match")
        out += [(int(q > rank), ro, so, cnt) for (ro, cnt), (so, _) in zip(recvs, sends)]
    return out


def push_plan(part, level, rank, nx, ny, shared_from_lower=False):
    """The halo exchange seen from the SENDING side: ``[(to_upper, my_offset, peer_offset, count)]``,
    the ``count`` elements at ``my_offset`` of this rank's local field belong at ``peer_offset`` of
    the neighbour's (``rank + 1`` if ``to_upper`` else ``rank - 1``): the neighbours' pull plans
    read backwards."""
    out = []
    for q in (rank - 1, rank + 1):
        if q < 0 or q >= part.nranks:
            continue
        for from_upper, q_off, my_off, cnt in pull_plan(part, level, q, nx, ny, shared_from_lower):
            if bool(from_upper) == (rank > q):           # q pulls this one from me
                out.append((int(q > rank), my_off, q_off, cnt))
    return out


def gather_plan(part, level, rank, nx, ny):
    """All-gather of the owned parts of a local field into the global layout.

    Returns ``(copies, sends, recvs)``: ``copies`` = [(local_off, global_off,
    count)] for the rank's own part, ``sends`` = [(peer, local_off, count)],
    ``recvs`` = [(peer, global_off, count)].
    """
    nzl = part.nz_level(level)
    (gx, gy, gz), (gox, goy, goz) = _comp_sizes(nx, ny, nzl)
    mine = owned_ranges(part, level, rank, nx, ny)

    def global_ranges(r):
        p0, p1 = part.owned(level, r)
        l0, l1 = max(p0 - 1, 0), p1 - 1
        return [(gox + gx * p0, gx * (p1 - p0)), (goy + gy * p0, gy * (p1 - p0)),
                (goz + gz * l0, gz * (l1 - l0))]

    copies = [(lo_, go_, n) for (lo_, n), (go_, _) in zip(mine, global_ranges(rank))]
    sends, recvs = [], []
    for q in range(part.nranks):
        if q == rank:
            continue
        for (lo_, n) in mine:
            sends.append((q, lo_, n))
        for (go_, n) in global_ranges(q):
            recvs.append((q, go_, n))
    return copies, sends, recvs


def scatter_ranges(part, level, rank, nx, ny):
    """[(global_off, local_off, count)] copying a rank's local slab (with halos)
    out of a global field."""
    lo, hi = part.local(level, rank)
    nzl = part.nz_level(level)
    (gx, gy, gz), (gox, goy, goz) = _comp_sizes(nx, ny, nzl)
    (px, py, pz), (ox, oy, oz) = _comp_sizes(nx, ny, hi - lo)
    return [(gox + gx * lo, ox, px * (hi - lo + 1)), (goy + gy * lo, oy, py * (hi - lo + 1)),
            (goz + gz * lo, oz, pz * (hi - lo))]


def zline_schedule(nranks, nbatch, rank):
    """Steps of one colour class of the z-line relaxation across slabs, for ``rank``: a list of
    ``(phase, batch)``, phase 1 = forward substitution (ranks bottom-up), 2 = backward (top-down);
    ``batch`` = the batch of lines this rank relaxes in the step, or None (it only takes part in
    the interface exchange that follows EVERY step on every rank).  Batch b reaches rank r in
    forward step b + r and in backward step b + (nranks - 1 - r): a pipeline of
    ``nbatch + nranks - 1`` steps per phase."""
    steps = []
    for phase in (1, 2):
        pos = rank if phase == 1 else nranks - 1 - rank
        for t in range(nbatch + nranks - 1):
            b = t - pos
            steps.append((phase, b if 0 <= b < nbatch else None))
    return steps


def slab_sparse(part, rank, nx, ny, idx, val):
    """The entries (global flat indices ``idx``, values ``val``) of a sparse field that fall into
    the local slab of ``rank`` (halo planes included), as local flat indices and values."""
    loc_i, loc_v = [], []
    for goff, loff, n in scatter_ranges(part, 0, rank, nx, ny):
        m = (idx >= goff) & (idx < goff + n)
        loc_i.append(idx[m] - goff + loff)
        loc_v.append(val[m])
    return np.concatenate(loc_i), np.concatenate(loc_v)


# =========================================================================== #
# Device transport and distributed driver
# =========================================================================== #

class NcclComm:
    """NCCL communicator owned by the C library (one per process)."""

    def __init__(self, rank, nranks, broadcast):
        """``broadcast(obj_or_None)``: returns rank 0's object on every rank."""
        import ctypes
        from emg3d_b200 import _lib
        self._lib, self.rank, self.nranks = _lib, int(rank), int(nranks)
        lib = _lib.init()
        uid = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(lib.emg3d_b200_comm_unique_id(uid))
        raw = broadcast(bytes(uid.raw) if self.rank == 0 else None)
        buf = ctypes.create_string_buffer(raw, 128)
        _lib.check(lib.emg3d_b200_comm_init(buf, self.nranks, self.rank))

    def sendrecv(self, base_ptr, itemsize, plan):
        """plan: [(is_send, peer, offset, count)] in elements of the array at base_ptr."""
        import ctypes
        n = len(plan)
        if n == 0:
            return
        ptrs = (ctypes.c_void_p * n)(*[base_ptr + off * itemsize for _, _, off, _ in plan])
        nbytes = (ctypes.c_size_t * n)(*[cnt * itemsize for _, _, _, cnt in plan])
        peers = (ctypes.c_int * n)(*[p for _, p, _, _ in plan])
        sends = (ctypes.c_int * n)(*[s for s, _, _, _ in plan])
        self._lib.check(self._lib.load().emg3d_b200_comm_sendrecv(n, ptrs, nbytes, peers, sends))

    def sendrecv_two(self, src_ptr, dst_ptr, itemsize, sends, recvs):
        """sends from one array, receives into another (gather)."""
        plan_ptrs = [(1, p, src_ptr + off * itemsize, cnt * itemsize) for p, off, cnt in sends]
        plan_ptrs += [(0, p, dst_ptr + off * itemsize, cnt * itemsize) for p, off, cnt in recvs]
        import ctypes
        n = len(plan_ptrs)
        if n == 0:
            return
        ptrs = (ctypes.c_void_p * n)(*[x[2] for x in plan_ptrs])
        nbytes = (ctypes.c_size_t * n)(*[x[3] for x in plan_ptrs])
        peers = (ctypes.c_int * n)(*[x[1] for x in plan_ptrs])
        snd = (ctypes.c_int * n)(*[x[0] for x in plan_ptrs])
        self._lib.check(self._lib.load().emg3d_b200_comm_sendrecv(n, ptrs, nbytes, peers, snd))

    # ---- halo exchange over peer memory (one kernel per exchange, csrc/comm.cu) ----
    def p2p_enable(self):
        """Collective.  True if every rank can map its neighbours' memory (CUDA IPC)."""
        import ctypes
        import os
        if os.environ.get('EMG3D_B200_P2P', '1') == '0' or self.nranks < 2:
            self.p2p = False
            return False
        on = ctypes.c_int(0)
        self._lib.check(self._lib.load().emg3d_b200_p2p_init(ctypes.byref(on)))
        self.p2p = bool(on.value)
        return self.p2p

    def p2p_register(self, ptr):
        """Collective, same order on all ranks.  Slot id, or -1 (use sendrecv)."""
        import ctypes
        slot = ctypes.c_int(-1)
        self._lib.check(self._lib.load().emg3d_b200_p2p_register(ptr, ctypes.byref(slot)))
        return slot.value

    @staticmethod
    def p2p_args(pulls, itemsize):
        """ctypes argument arrays of p2p_exchange for a :func:`pull_plan`."""
        import ctypes
        n = len(pulls)
        A = ctypes.c_size_t * n
        return (n, A(*[mo * itemsize for _, mo, _, _ in pulls]),
                A(*[po * itemsize for _, _, po, _ in pulls]),
                A(*[c * itemsize for _, _, _, c in pulls]),
                (ctypes.c_int * n)(*[u for u, _, _, _ in pulls]))

    def p2p_exchange(self, slot, args, push=False):
        if args[0]:
            self._lib.check(self._lib.load().emg3d_b200_p2p_exchange(slot, *args, int(bool(push))))

    def p2p_release(self):
        """Unmap every registered array (before the arrays are freed)."""
        if getattr(self, 'p2p', False):
            self._lib.check(self._lib.load().emg3d_b200_p2p_release())

    def p2p_status(self):
        import ctypes
        st = ctypes.c_int(0)
        self._lib.check(self._lib.load().emg3d_b200_p2p_status(ctypes.byref(st)))
        return st.value

    def allreduce_sum(self, dev_array, n=None):
        self._lib.check(self._lib.load().emg3d_b200_comm_allreduce_sum(
            dev_array.ptr, dev_array.size if n is None else int(n)))

    def destroy(self):
        self._lib.load().emg3d_b200_comm_destroy()


class _DLevel:
    """One distributed level: the rank's local `_Level` plus its exchange plans."""

    def __init__(self, level, lv, part, rank):
        nx, ny = lv.shape[0], lv.shape[1]
        self.index, self.lv = level, lv
        self.plan = exchange_plan(part, level, rank, nx, ny)
        self.pulls = pull_plan(part, level, rank, nx, ny)
        self.pull_args = None                        # ctypes arrays, built on first use
        # the same exchange with the shared fz layer of every interface taken from the LOWER rank
        self.plan_low = exchange_plan(part, level, rank, nx, ny, shared_from_lower=True)
        self.pulls_low = pull_plan(part, level, rank, nx, ny, shared_from_lower=True)
        self.pull_args_low = None
        # the same two exchanges from the sending side (push variant of the peer-memory kernel)
        self.pushes = push_plan(part, level, rank, nx, ny)
        self.pushes_low = push_plan(part, level, rank, nx, ny, shared_from_lower=True)
        self.push_args = self.push_args_low = None
        self.owned = owned_ranges(part, level, rank, nx, ny)
        # Smoother and residual run on the z-window [p0 - 1, hi] of the local grid:
        # one halo plane on either side, refreshed by the exchange and fixed during a
        # sweep.  The deeper halo planes below only exist to keep the local grids
        # nested under coarsening; no kernel result ever depends on them.
        lo, hi = part.local(level, rank)
        p0, p1 = part.owned(level, rank)
        z0 = 0 if rank == 0 else p0 - 1 - lo
        self.win = lv.handle if z0 == 0 else lv.handle.window(z0, hi - lo - z0)
        # norms evaluated by the residual kernel on the window count owned edges only
        import ctypes
        from emg3d_b200 import _lib
        lib = _lib.load()
        _lib.check(lib.emg3d_b200_level_set_owned(self.win.ptr, p0 - lo - z0, p1 - lo - z0))
        # Exact (true Gauss-Seidel) sweeps across slabs: the colour classes are run in two
        # z-halves with a halo exchange after each.  Node / line colours use the GLOBAL z-parity
        # (local node plane 1 of the window is global plane lo + z0 + 1) ...
        _lib.check(lib.emg3d_b200_level_set_zflip(self.win.ptr, int((lo + z0 + 1) % 2 == 0)))
        # ... the tile-fused point schedule colours by local tile index: planes on either side of
        # an interface fall into different halves iff the top tile of the lower rank is odd
        kind, tile = ctypes.c_int(0), (ctypes.c_int * 3)()
        self.point_kind = kind
        _lib.check(lib.emg3d_b200_point_schedule_kind(self.win.ptr, ctypes.byref(kind)))
        _lib.check(lib.emg3d_b200_point_tile_shape(tile))
        nzw = hi - lo - z0                               # cells of the window
        self.point_halves_ok = (kind.value != 2 or rank == part.nranks - 1
                                or ((nzw - 2) // tile[2]) % 2 == 1)
        lv.res_buffer().zero()


class _Shape:
    """Stand-in for a grid where only ``shape_cells`` matters (global level shapes)."""

    def __init__(self, shape):
        self.shape_cells = tuple(int(n) for n in shape)


def hierarchy_plan(gshape, pattern, nranks, clevel, min_cells=1_000_000):
    """Global shapes and transitions of the distributed levels of one semicoarsening pattern.

    ``pattern``: the cycle's ``sc_dir`` (0 = standard coarsening, 1 / 2 / 3 = no coarsening along
    x / y / z; solver.py:1482-1531).  Returns ``(shapes, trans, zshifts)``: ``shapes[l]`` the
    GLOBAL cell shape of level ``l`` (``l = 0 .. n_dist``), ``trans[l]`` the effective pattern of
    the transition ``l -> l + 1`` (which axes are halved), ``zshifts[l]`` the number of
    z-coarsenings above level ``l``.  Level ``l >= 1`` is distributed while it has more than
    ``min_cells`` cells (below that a level is launch-latency bound and replicating it is cheaper
    than exchanging its halos), the level below it leaves every rank two planes, and the
    hierarchy goes on below it (``clevel``: number of coarsening steps of the pattern).
    """
    from emg3d_b200 import core, solver
    shapes, trans, zshifts = [tuple(int(n) for n in gshape)], [], [0]

    def step(shape):
        sc = int(solver._current_sc_dir(pattern, _Shape(shape)))
        flag = core.SC_FLAGS[sc]
        return sc, tuple(n // 2 if f else n for n, f in zip(shape, flag)), int(bool(flag[2]))

    while True:
        k = len(trans)                                   # levels 0 .. k exist; decide on k + 1
        if k + 1 > clevel:
            break
        sc, nxt, dz = step(shapes[-1])
        if nxt == shapes[-1]:
            break
        if k >= 1:
            # level k becomes distributed only if it is big enough and level k + 1 still leaves
            # two planes per rank
            if int(np.prod(shapes[-1])) <= min_cells or nxt[2] // nranks < 2 or min(nxt[:2]) < 2:
                break
        shapes.append(nxt)
        trans.append(sc)
        zshifts.append(zshifts[-1] + dz)
    if not trans:
        raise ValueError("grid cannot be coarsened: nothing to distribute")
    return shapes, trans, zshifts


class _Chain:
    """Distributed levels of one semicoarsening pattern (built on first use)."""

    def __init__(self, dmg, pattern, shapes, trans, zshifts):
        self.pattern, self.shapes, self.trans = pattern, shapes, trans
        self.n_dist = len(trans)
        self.part = SlabPartition(dmg.gshape[2], dmg.nranks, self.n_dist, zshifts, dmg.zmax)
        self.levels = [dmg.level0]
        for l in range(1, self.n_dist + 1):              # level n_dist: gather buffer only
            child = self.levels[-1].lv.coarse(trans[l - 1])
            self.levels.append(_DLevel(l, child, self.part, dmg.rank))
        for dl, shp in zip(self.levels, shapes):
            dl.gshape = shp
        top = self.levels[self.n_dist]
        tnx, tny = top.lv.shape[0], top.lv.shape[1]
        self.glevel = dmg._build_global_level(self)
        self.g_s, self.g_e = self.glevel.new_field(), self.glevel.new_field()
        self.gather = gather_plan(self.part, self.n_dist, dmg.rank, tnx, tny)
        self.scatter = scatter_ranges(self.part, self.n_dist, dmg.rank, tnx, tny)


class DistributedMultigrid:
    """One solve on N GPUs (one instance per rank): multigrid cycles and Krylov wrappers.

    Parameters
    ----------
    model : Model
        The GLOBAL model (every rank holds it on the host; only the rank's slab
        is uploaded).
    sfield : Field
        The GLOBAL source field.
    comm : NcclComm
    semicoarsening, linerelaxation : as in :func:`emg3d_b200.solve`; they select the
        hierarchies that are planned (one per semicoarsening pattern in the cycle).
    n_dist : int, optional
        Number of distributed levels (standard coarsening only); default: the levels with more
        than about a million cells (the coarser ones are replicated on every GPU).
    exact : bool
        True (default): two halo exchanges per sweep, a true Gauss-Seidel sweep across slabs;
        False: one exchange per sweep (see :meth:`smoothing`).
    """

    def __init__(self, model, sfield, comm, n_dist=None, order=None, exact=None,
                 semicoarsening=False, linerelaxation=False):
        import os
        from emg3d_b200 import _lib, core, meshes, models, solver
        if exact is None:
            exact = os.environ.get('EMG3D_B200_DIST_EXACT', '1') != '0'
        self.exact = bool(exact)
        self._lib, self._solver = _lib, solver
        self.comm, self.rank, self.nranks = comm, comm.rank, comm.nranks
        self._slots = {}                             # device pointer -> peer-memory slot
        self._graphs = {}                            # captured visits of level 1
        if not getattr(comm, 'p2p', False) and hasattr(comm, 'p2p_enable'):
            comm.p2p_enable()
        self.order = core.order_id(order)
        self.gshape = tuple(int(n) for n in model.grid.shape_cells)
        nx, ny, nz = self.gshape
        self.model_grid = model.grid
        self.dtype = solver._field_dtype(sfield)
        self.frequency = sfield._frequency

        # --- plan the hierarchies of every semicoarsening pattern of the cycle -----------
        probe = solver.MGParameters(verb=-1, sslsolver=False, semicoarsening=semicoarsening,
                                    linerelaxation=linerelaxation, shape_cells=self.gshape,
                                    cycle='V')
        self.patterns = sorted(set(int(v) for v in probe.raw_sc_cycle))
        self._plans = {}
        for pat in self.patterns:
            if n_dist is not None and pat == 0:
                shapes, trans, zs = hierarchy_plan(self.gshape, 0, self.nranks, int(n_dist), min_cells=0)
            else:
                shapes, trans, zs = hierarchy_plan(self.gshape, pat, self.nranks,
                                                   int(probe.clevel[pat]) - 1)
            self._plans[pat] = (shapes, trans, zs)
        self.zmax = max(zs[-1] for _, _, zs in self._plans.values())
        self.n_dist = len(self._plans[self.patterns[0]][1])       # (of the first pattern)

        # --- local finest level: sliced model, device-side VolumeModel -----------------
        part0 = SlabPartition(nz, self.nranks, 0, [0], self.zmax)
        lo, hi = part0.local(0, self.rank)
        g = model.grid
        nodes_z = np.r_[0., np.asarray(g.h[2]).cumsum()] + g.origin[2]
        lgrid = meshes.TensorMesh([g.h[0], g.h[1], np.asarray(g.h[2])[lo:hi]],
                                  (g.origin[0], g.origin[1], nodes_z[lo]))
        sl = {}
        for name in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r'):
            a = getattr(model, name)
            sl[name] = None if a is None else np.asfortranarray(np.asarray(a)[:, :, lo:hi])
        lmodel = models.Model(lgrid, mapping=getattr(model.map, 'name', 'Resistivity'), **sl)
        lmodel.case = model.case
        lv0 = solver._Level.from_model(lmodel, sfield)
        self.part0 = part0
        self.level0 = _DLevel(0, lv0, part0, self.rank)
        self.level0.gshape = self.gshape
        self._chains = {}
        self.levels = self.chain(self.patterns[0]).levels         # (first pattern; kept for tools)
        self.part = self.chain(self.patterns[0]).part

        # --- local source and field ----------------------------------------------------
        self.s = lv0.new_field()
        self.e = lv0.new_field()
        self.upload_source(sfield)
        self.p2p_push = os.environ.get('EMG3D_B200_P2P_PUSH', '1') != '0'
        self._sums = _lib.DeviceArray(8, np.float64)
        self._krylov_pool = []
        self._rec = self._seg = None                     # graph segments being recorded (_descend)
        self._sums.zero()

    def chain(self, pattern):
        """Distributed hierarchy of a semicoarsening pattern (device levels built on first use)."""
        pattern = int(pattern)
        ch = self._chains.get(pattern)
        if ch is None:
            if pattern not in self._plans:
                raise ValueError(f"semicoarsening pattern {pattern} was not planned for this solver")
            ch = self._chains[pattern] = _Chain(self, pattern, *self._plans[pattern])
        return ch

    def close(self):
        """Release the peer-memory mappings of this solver's arrays.  Call on every rank (and
        synchronise the ranks) before the instance is dropped and another one is created."""
        self._lib.sync()
        if hasattr(self.comm, 'p2p_release'):
            self.comm.p2p_release()
        self._slots.clear()
        self._graphs.clear()
        # Every rank has now unmapped its neighbours' arrays; nobody may FREE its own (exported)
        # arrays before all ranks got here -- freeing memory a peer still has open through CUDA IPC
        # is undefined (seen as an illegal address in a later, unrelated solve of the same
        # process).  One all-reduce + sync is the barrier.
        if hasattr(self.comm, 'allreduce_sum') and getattr(self, '_sums', None) is not None:
            self.comm.allreduce_sum(self._sums, 1)
            self._lib.sync()

    # ---- setup helpers ------------------------------------------------------------
    def _build_global_level(self, chain):
        """First replicated level: all-gather the owned cell layers of the local
        coarse coefficient slabs (device to device) into global arrays."""
        from emg3d_b200 import core, meshes, solver
        _lib = self._lib
        level = chain.n_dist
        top = chain.levels[level].lv
        nx, ny = top.shape[0], top.shape[1]
        nzg = chain.part.nz_level(level)
        g = self.model_grid
        ch = [np.asarray(h, dtype=float) for h in g.h]
        for sc in chain.trans:                               # coarsen the global widths alike
            flag = core.SC_FLAGS[sc]
            ch = [np.add.reduceat(h, np.arange(0, len(h), 2)) if f else h for h, f in zip(ch, flag)]
        cgrid = meshes.BaseMesh(ch, g.origin)
        nxy = nx * ny

        def layers(r):                               # owned cell layers [l0, l1) of rank r
            p0, p1 = chain.part.owned(level, r)
            return max(p0 - 1, 0), p1 - 1

        lo, _ = chain.part.local(level, self.rank)

        def gather(local):
            out = _lib.DeviceArray(nxy * nzg, local.dtype)
            isz = local.dtype.itemsize
            l0, l1 = layers(self.rank)
            _lib.check(_lib.load().emg3d_b200_d2d(out.ptr + l0 * nxy * isz,
                                                  local.ptr + (l0 - lo) * nxy * isz,
                                                  (l1 - l0) * nxy * isz))
            sends, recvs = [], []
            for q in range(self.nranks):
                if q == self.rank:
                    continue
                sends.append((q, (l0 - lo) * nxy, (l1 - l0) * nxy))
                q0, q1 = layers(q)
                recvs.append((q, q0 * nxy, (q1 - q0) * nxy))
            self.comm.sendrecv_two(local.ptr, out.ptr, isz, sends, recvs)
            return out

        eta = []
        for k, a in enumerate(top.eta):
            for j in range(k):
                if a is top.eta[j]:
                    eta.append(eta[j])
                    break
            else:
                eta.append(gather(a))
        zeta = gather(top.zeta)
        _lib.sync()
        return solver._Level(cgrid, self.dtype, top.case, eta, zeta)

    def _slab(self, field_1d):
        """Local slab (with halos) of a global host field (finest level)."""
        nx, ny = self.gshape[0], self.gshape[1]
        out = np.empty(self.level0.lv.n_edges, dtype=self.dtype)
        for goff, loff, n in scatter_ranges(self.part0, 0, self.rank, nx, ny):
            out[loff:loff + n] = field_1d[goff:goff + n]
        return out

    def upload_source(self, sfield):
        """The rank's slab of the source field; a sparse source (fields.SourceField) sends only the
        non-zero edges that fall into the slab (halo planes included)."""
        sparse = getattr(sfield, 'sparse', None)
        if sparse is None:
            self.s.upload(self._slab(np.asarray(sfield.field)))
            return
        idx, val, bg = sparse
        loc_i, loc_v = slab_sparse(self.part0, self.rank, self.gshape[0], self.gshape[1], idx, val)
        self.s.fill_scatter(bg, loc_i, loc_v)

    def gather_to_root(self, src=None, root=0):
        """The whole field on the device of rank ``root`` (global layout): every rank sends its owned
        parts over NVLink (NCCL send / recv).  Returns the device array on ``root``, None elsewhere."""
        src = self.e if src is None else src
        isz = self.dtype.itemsize
        nx, ny = self.gshape[0], self.gshape[1]
        copies, sends, recvs = gather_plan(self.part0, 0, self.rank, nx, ny)
        if self.rank == root:
            n = sum(c for _, _, c in copies) + sum(c for _, _, c in recvs)
            full = self._lib.DeviceArray(n, self.dtype, scratch=True)     # (pool memory: no cudaMalloc / cudaFree per solve)
            lib = self._lib.load()
            for loff, goff, cnt in copies:
                self._lib.check(lib.emg3d_b200_d2d(full.ptr + goff * isz, src.ptr + loff * isz, cnt * isz))
            self.comm.sendrecv_two(src.ptr, full.ptr, isz, [], recvs)
            return full
        self.comm.sendrecv_two(src.ptr, src.ptr, isz, [(p, o, c) for p, o, c in sends if p == root], [])
        return None

    def upload_field(self, efield, dst=None):
        (self.e if dst is None else dst).upload(self._slab(np.asarray(efield.field)))

    def download_owned(self, out_global, src=None):
        """Write the owned part of a local field (default: e) into a global host array."""
        loc = (self.e if src is None else src).download()
        nx, ny = self.gshape[0], self.gshape[1]
        copies, _, _ = gather_plan(self.part0, 0, self.rank, nx, ny)
        for loff, goff, n in copies:
            out_global[goff:goff + n] = loc[loff:loff + n]

    # ---- distributed building blocks -------------------------------------------------
    def exchange(self, dl, field, shared_from_lower=False):
        """Refresh the halo planes of `field` (a local field of level `dl`); the shared fz layer
        of an interface comes from the upper rank, or from the lower one (see exchange_plan)."""
        if self.comm.p2p:
            slot = self._slots.get(field.ptr)
            if slot is None:                         # first exchange of this array: collective
                slot = self._slots[field.ptr] = self.comm.p2p_register(field.ptr)
            if slot >= 0 and self.p2p_push:
                if shared_from_lower:
                    if dl.push_args_low is None:
                        dl.push_args_low = self.comm.p2p_args(dl.pushes_low, self.dtype.itemsize)
                    self.comm.p2p_exchange(slot, dl.push_args_low, push=True)
                    return
                if dl.push_args is None:
                    dl.push_args = self.comm.p2p_args(dl.pushes, self.dtype.itemsize)
                self.comm.p2p_exchange(slot, dl.push_args, push=True)
                return
            if slot >= 0:
                if shared_from_lower:
                    if dl.pull_args_low is None:
                        dl.pull_args_low = self.comm.p2p_args(dl.pulls_low, self.dtype.itemsize)
                    self.comm.p2p_exchange(slot, dl.pull_args_low)
                    return
                if dl.pull_args is None:
                    dl.pull_args = self.comm.p2p_args(dl.pulls, self.dtype.itemsize)
                self.comm.p2p_exchange(slot, dl.pull_args)
                return
        self.comm.sendrecv(field.ptr, self.dtype.itemsize, dl.plan_low if shared_from_lower else dl.plan)

    def check_transport(self):
        """Raise if a peer-memory halo exchange timed out (csrc/comm.cu sets a status word and
        lets the kernel finish; the fields are then inconsistent across ranks)."""
        if getattr(self.comm, 'p2p', False):
            st = self.comm.p2p_status()
            if st:
                raise self._lib.Emg3dB200Error(
                    f"peer-memory halo exchange timed out (status {st}): a neighbouring rank did "
                    "not arrive within the spin limit; the fields are inconsistent")

    def sum_owned(self, dl, x, y=None):
        """sum over owned edges of conj(x) y (default y = x), all-reduced."""
        lib = self._lib.load()
        y = x if y is None else y
        isz = self.dtype.itemsize
        for k, (off, n) in enumerate(dl.owned):
            self._lib.check(lib.emg3d_b200_dot(int(self.dtype.kind == 'c'), n, x.ptr + off * isz,
                                               y.ptr + off * isz, 1, self._sums.ptr + 16 * k))
        self.comm.allreduce_sum(self._sums)
        v = self._sums.download()
        return complex(v[0] + v[2] + v[4], v[1] + v[3] + v[5])

    def residual(self, dl, s, e, norm=False):
        lib = self._lib.load()
        if norm:
            # norm only: no residual is written; the kernel sums |r|^2 over the owned
            # edges of the window (level_set_owned) into a device scalar, all-reduced
            self._lib.check(lib.emg3d_b200_residual(dl.win.ptr, s.ptr, e.ptr, None, self._sums.ptr))
            self.comm.allreduce_sum(self._sums, 1)
            return float(np.sqrt(self._sums.download()[0]))
        r = dl.lv.res_buffer()
        self._lib.check(lib.emg3d_b200_residual(dl.win.ptr, s.ptr, e.ptr, r.ptr, None))
        self.exchange(dl, r)
        return r

    def apply(self, dl, src, dst):
        """dst = A src on the owned planes (halos of src refreshed first)."""
        self.exchange(dl, src)
        self._lib.check(self._lib.load().emg3d_b200_apply(dl.win.ptr, src.ptr, dst.ptr))

    def smoothing(self, dl, s, e, nu, lr_dir):
        """``nu`` sweeps per line direction on the rank's window.

        Exact variant (default): every multicolour sweep is run as its two z-halves
        (the classes of even, then odd z-parity, or the reverse in a descending sweep) with a
        halo exchange after each.  Planes on either side of an interface belong to different
        halves, so every relaxed node sees current neighbour values: a true multicolour
        Gauss-Seidel sweep, as on one GPU.  Relaxed variant (``exact=False``, what
        ``north_star`` words): one exchange per sweep, interface planes see values one sweep
        old (block-Jacobi across slabs) -- half the messages, weaker smoothing at interfaces.
        z-lines cross the slabs: :meth:`zline_smoothing`.
        """
        solver, lib = self._solver, self._lib.load()
        c_lr_dir = int(solver._current_lr_dir(lr_dir, _Shape(dl.gshape)))
        dirs = solver._LR_DIRS[c_lr_dir] or (0,)
        for ldir in dirs:
            if ldir == 3:
                self.zline_smoothing(dl, s, e, nu)
                continue
            halves = self.exact and self.order == 1 and (ldir != 0 or dl.point_halves_ok)
            for sweep in range(int(nu)):
                base = self.order | (sweep << 8)
                if not halves:
                    self._lib.check(lib.emg3d_b200_gauss_seidel(dl.win.ptr, e.ptr, s.ptr, 1, ldir, base))
                    self.exchange(dl, e)
                    continue
                # sweep 0 is the descending one (classes of odd z-parity first).  The half that
                # relaxed the LOWER rank's top plane is followed by an exchange that carries the
                # lower rank's copy of the shared fz layer upwards (its latest update); the other
                # half by the default exchange.  Global parity (node colours, lines): ownership
                # boundaries are even planes, the plane below an interface is odd = class bit 0
                # = half 1; tile-fused point schedule: the lower rank's top tile is odd = half 2.
                low_half = 2 if (ldir == 0 and dl.point_kind.value == 2) else 1
                for half in ((2, 1) if sweep % 2 == 0 else (1, 2)):
                    self._lib.check(lib.emg3d_b200_gauss_seidel(
                        dl.win.ptr, e.ptr, s.ptr, 1, ldir, base | (half << 16)))
                    self.exchange(dl, e, shared_from_lower=(half == low_half))

    def _zline_chain(self, dl):
        """Factorise the z-lines of a distributed level as pieces of the GLOBAL lines (once): the
        block recurrence S_m = D_m - E_m S_{m-1}^{-1} E_m runs through the slab cuts, every rank
        continues from the factors of the lower rank's last blocks (10 numbers per line)."""
        if getattr(dl, 'zchain', False):
            return
        import ctypes
        lib = self._lib.load()
        n = ctypes.c_size_t(0)
        self._lib.check(lib.emg3d_b200_level_line_chain(dl.win.ptr, 3, None, None, ctypes.byref(n)))
        n, isz = int(n.value), self.dtype.itemsize
        first, last = self.rank == 0, self.rank == self.nranks - 1
        xin = None if first else self._lib.DeviceArray(n, self.dtype)
        xout = None if last else self._lib.DeviceArray(n, self.dtype)
        if xin is not None:
            self.comm.sendrecv(xin.ptr, isz, [(0, self.rank - 1, 0, n)])
        if xin is not None or xout is not None:
            self._lib.check(lib.emg3d_b200_level_line_chain(
                dl.win.ptr, 3, None if xin is None else xin.ptr, None if xout is None else xout.ptr, None))
        if xout is not None:
            self.comm.sendrecv(xout.ptr, isz, [(1, self.rank + 1, 0, n)])
        self._lib.sync()                                 # (xin / xout are released on return)
        dl.zchain = True

    def zline_smoothing(self, dl, s, e, nu):
        """z-line relaxation (core.py:1071-1348) with the lines cut by the slabs: EXACT, every global
        line is solved as on one GPU.  Per colour class the forward block substitution runs rank
        after rank upwards -- the intermediate g of a piece's last node reaches the next rank in
        its halo plane, where the kernel reads its start value -- and the backward substitution
        rank after rank downwards (the true T of the upper piece's first node arrives in the top
        halo plane).  The shared fz layer of an interface is computed by the lower rank
        (``shared_from_lower`` exchanges only: the lower rank's copy also parks its intermediate
        there).  Same colour sequence as the single-GPU kernel (gs_line.cu: gs_dir)."""
        if self.order != 1:
            raise NotImplementedError("z-line relaxation across z-slabs: multicolour order only")
        lib = self._lib.load()
        self._zline_chain(dl)
        n = self.nranks
        # The lines of a colour class are independent: cut into `nb` batches, the ranks work as a
        # pipeline (rank r handles batch t - r in step t of the forward phase, and batch
        # t - (n - 1 - r) in the backward phase) -- nb + n - 1 steps per phase instead of the n
        # fully serial ones, every step followed by the interface exchange.
        import os
        # (measured on 2 B200, 256 x 256 x 128: batches of a few thousand one-thread-per-line
        # solves are latency-bound and the extra exchanges cost more than the overlap gains --
        # 0.42 s / 0.46 s / 0.54 s per solve for 1 / 2 / 4 batches; from 4 ranks on the serial
        # chain dominates: one batch per rank)
        nb = min(16, int(os.environ.get('EMG3D_B200_ZBATCH', 0)) or (n if n >= 4 else 1))
        for sweep in range(int(nu)):
            back = sweep % 2 == 0
            for cc in range(4):
                if sweep > 0 and cc == 0:
                    continue                             # idempotent repeat (gs_line.cu)
                cg = 3 - cc if back else cc
                for phase, b in zline_schedule(n, nb, self.rank):
                    if b is not None:
                        order = 1 | (phase << 18) | ((cg + 1) << 20) | (b << 23) | ((nb - 1) << 27)
                        self._lib.check(lib.emg3d_b200_gauss_seidel(dl.win.ptr, e.ptr, s.ptr, 1, 3, order))
                    self.exchange(dl, e, shared_from_lower=True)

    # ---- the cycle ---------------------------------------------------------------------
    def multigrid(self, var, level=0, new_cycmax=0, s=None, e=None):
        """Distributed counterpart of solver._multigrid (same control flow)."""
        solver, lib = self._solver, self._lib.load()
        ch = self.chain(var.sc_dir)
        dl = ch.levels[level]
        if level == 0:
            s = self.s if s is None else s
            e = self.e if e is None else e
        else:
            s, e = dl.lv.s, dl.lv.e
        it = 0
        if new_cycmax == 0 or var.cycle != 'F':
            cycmax = var.cycmax
        else:
            cycmax = new_cycmax
        cyc = 0
        if level == 0 and getattr(var, 'e_is_zero', False) and var.s_norm is not None:
            l2_last = float(var.s_norm)              # zero start field: ||r|| = ||s||
            var.e_is_zero = False
        else:
            var.e_is_zero = False
            l2_last = self.residual(dl, s, e, norm=True) if level == 0 else 0.0
        l2_stag = np.ones(var.maxcycle) * l2_last
        if level == 0 and var.nu_init > 0:
            self.smoothing(dl, s, e, var.nu_init, var.lr_dir)
        while level == 0 or it < cycmax:
            l2_prev = l2_last
            l2_stag[(it - 1) % var.maxcycle] = l2_last
            if level == 0:
                ch = self.chain(var.sc_dir)          # the pattern may change from cycle to cycle
            if var.nu_pre > 0:
                self.smoothing(dl, s, e, var.nu_pre, var.lr_dir)
            res = self.residual(dl, s, e)
            child = ch.levels[level + 1]
            self._lib.check(lib.emg3d_b200_restrict(child.lv.handle.ptr, res.ptr, child.lv.s.ptr))
            child.lv.e.zero()
            self._descend(var, ch, child, level + 1, cycmax - cyc)
            self._lib.check(lib.emg3d_b200_prolong(child.lv.handle.ptr, e.ptr, child.lv.e.ptr))
            self.exchange(dl, e)
            if var.nu_post > 0:
                self.smoothing(dl, s, e, var.nu_post, var.lr_dir)
            it += 1
            if level > 0:
                cyc += 1
            else:
                var.it += 1
                l2_last = self.residual(dl, s, e, norm=True)
                self.check_transport()
                solver._print_cycle_info(var, l2_last, l2_prev)
                if var.sc_cycle:                         # as solver._multigrid (solver.py:639-642)
                    var.sc_dir = next(var.sc_cycle)
                if var.lr_cycle:
                    var.lr_dir = next(var.lr_cycle)
                if solver._terminate(var, l2_last, l2_stag[(it - 1) % var.maxcycle], it):
                    break
        var.l2 = l2_last

    def _descend(self, var, ch, child, level, new_cycmax):
        """Everything between restriction to and prolongation from `child`.

        Below the finest level a visit is a fixed sequence of launches (smoothers, transfer
        kernels, halo-exchange kernels -- no norms, no host decisions) around the gather + replicated
        coarse sub-cycle.  With ``EMG3D_B200_DIST_GRAPHS=1`` the visit of level 1 is recorded on its
        second execution as a list of SEGMENTS: every stretch of launches between two gathers
        becomes one CUDA graph, the gather (NCCL send / recv, not capturable together with the
        peer-memory flags: that hung in r1) and the replicated sub-cycle (which replays its own
        single-GPU graphs) stay eager; later visits replay the list.  The kernels of levels 1-2
        are shorter than the 10-15 us of host time an eager launch from Python costs.
        """
        def run():
            if level < ch.n_dist:
                self.multigrid(var, level, new_cycmax)
                self.exchange(child, child.lv.e)
            else:
                self._coarse_replicated(var, ch, child, level, new_cycmax)

        solver = self._solver
        import os
        if not (level == 1 and solver.GRAPHS and os.environ.get('EMG3D_B200_DIST_GRAPHS', '0') == '1'
                and var.verb <= 3 and self._rec is None
                and not getattr(var, '_capturing', False)):
            return run()
        key = (int(new_cycmax), var.cycle, int(var.sc_dir), int(var.lr_dir), solver._order(var),
               int(var.nu_pre), int(var.nu_post), int(var.nu_coarse), tuple(var.clevel))
        g = self._graphs.get(key)
        if g is None:                      # first visit: eager (builds caches, registers arrays)
            self._graphs[key] = False
            return run()
        if g is False:                     # second visit: record the segments while executing them
            self._rec = []
            self._seg_begin(var)
            try:
                run()
            finally:
                self._seg_end(var)
                items, self._rec = self._rec, None
            self._graphs[key] = items
            return
        for item in g:                     # replay
            if isinstance(item, tuple):
                self._coarse_body(var, *item)
            else:
                item.launch()

    def _seg_begin(self, var):
        var._capturing = True
        self._seg = self._lib.Graph()
        self._seg.__enter__()

    def _seg_end(self, var):
        """Close the open segment, keep it and run it (captured launches have not executed)."""
        seg, self._seg = self._seg, None
        var._capturing = False
        if seg is not None:
            seg.__exit__(None, None, None)
            self._rec.append(seg)
            seg.launch()

    def _coarse_replicated(self, var, ch, top, level, new_cycmax):
        """Gather the coarse source, solve the coarse sub-cycle redundantly, keep our slab."""
        if self._rec is not None:          # recording: this part stays eager, between two segments
            self._seg_end(var)
            self._rec.append((ch, top, level, new_cycmax))
            self._coarse_body(var, ch, top, level, new_cycmax)
            self._seg_begin(var)
            return
        self._coarse_body(var, ch, top, level, new_cycmax)

    def _coarse_body(self, var, ch, top, level, new_cycmax):
        isz = self.dtype.itemsize
        lib = self._lib.load()
        copies, sends, recvs = ch.gather
        for loff, goff, n in copies:
            self._lib.check(lib.emg3d_b200_d2d(ch.g_s.ptr + goff * isz, top.lv.s.ptr + loff * isz, n * isz))
        self.comm.sendrecv_two(top.lv.s.ptr, ch.g_s.ptr, isz, sends, recvs)
        ch.g_e.zero()
        self._solver._multigrid(ch.glevel, ch.g_s, ch.g_e, var, level=level,
                                new_cycmax=new_cycmax)
        for goff, loff, n in ch.scatter:
            self._lib.check(lib.emg3d_b200_d2d(top.lv.e.ptr + loff * isz, ch.g_e.ptr + goff * isz, n * isz))

    # ---- Krylov backend (solver._DeviceOps interface) ------------------------------------
    class _Ops:
        def __init__(self, dmg):
            self.dmg, self.dl, self._next = dmg, dmg.level0, 0
            self.vec = dmg._solver._Vec(int(dmg.dtype.kind == 'c'), dmg.level0.lv.n_edges)

        def new(self):
            # Work vectors are kept by the solver and reused by later solves: an exchanged array
            # is registered with the neighbours by ADDRESS (peer-memory slots); a vector freed
            # after one solve and another allocated at the same address by the next would be
            # served by the stale mapping of the neighbour's freed array.
            pool = self.dmg._krylov_pool
            if self._next == len(pool):
                pool.append(self.dmg.level0.lv.new_field())
            a = pool[self._next]
            self._next += 1
            a.zero()
            return a

        def norm(self, x):
            return float(np.sqrt(self.dmg.sum_owned(self.dl, x).real))

        def dot(self, x, y):
            v = self.dmg.sum_owned(self.dl, x, y)
            return v if self.dmg.dtype.kind == 'c' else v.real

        def axpby(self, a, x, b, y):
            self.vec.axpby(a, x, b, y)                  # (halo planes ride along; never read)

        def matvec(self, src, dst):
            self.dmg.apply(self.dl, src, dst)

        def psolve(self, src, dst, var):
            if var.cycle:
                dst.zero()
                var.e_is_zero, var.s_norm = True, None
                self.dmg.exchange(self.dl, src)
                self.dmg.multigrid(var, s=src, e=dst)
            else:
                dst.copy_from(src)

        def residual_norm(self, s, x):
            self.dmg.exchange(self.dl, x)
            return self.dmg.residual(self.dl, s, x, norm=True)

    # ---- solve ------------------------------------------------------------------------------
    def solve(self, sslsolver=False, semicoarsening=False, linerelaxation=False, verb=0,
              zero_start=True, **kwargs):
        """Collective.  Arguments and info dict of :func:`emg3d_b200.solve` (``return_info`` is
        implied); the field stays distributed (``download_owned``).  ``zero_start=False``: start
        from the current content of ``self.e`` (halos are refreshed)."""
        solver = self._solver
        if kwargs.pop('plain', False):
            sslsolver = False if sslsolver is True else sslsolver
            semicoarsening = False if semicoarsening is True else semicoarsening
            linerelaxation = False if linerelaxation is True else linerelaxation
        if sslsolver not in (False, True, 'bicgstab', 'cgs', 'gcrotmk'):
            raise ValueError("distributed Krylov wrappers: 'bicgstab' (True), 'cgs' and 'gcrotmk'. "
                             f"Provided: {sslsolver!r}.")
        kwargs.pop('return_info', None)
        var = solver.MGParameters(verb=verb, sslsolver=sslsolver, semicoarsening=semicoarsening,
                                  linerelaxation=linerelaxation, shape_cells=self.gshape,
                                  return_info=True, **kwargs)
        var.order = {0: 'lex', 1: 'color'}[self.order]
        if self.rank != 0:                               # one log, from rank 0
            var.verb, var.log = -1, 0
        missing = set(int(v) for v in var.raw_sc_cycle) - set(self._plans)
        if missing:
            raise ValueError(f"semicoarsening patterns {sorted(missing)} were not planned: pass "
                             "`semicoarsening` to DistributedMultigrid")
        for pat in set(int(v) for v in var.raw_sc_cycle):
            if var.clevel[pat] <= len(self._plans[pat][1]):
                raise ValueError("grid too small for the number of distributed levels")
        var.cprint(f"\n:: emg3d START :: {var.time.now} :: v{solver.__version__}\n", 2)
        var.cprint(var, 2)
        var.l2_refe = float(np.sqrt(self.sum_owned(self.level0, self.s).real))
        var.error_at_cycle[0] = var.l2_refe
        info = ""
        if zero_start:
            self.e.zero()
            var.e_is_zero, var.s_norm = not var.sslsolver, var.l2_refe
        else:
            self.exchange(self.level0, self.e)
            var.user_start = True
        if var.l2_refe < 100 * np.finfo(float).tiny:
            var.l2_refe = np.nan
            var.sslsolver = var.cycle = None
            var.exit_message = "CONVERGED"
            info = "   > RETURN ZERO E-FIELD (provided sfield is zero)\n"
            self.e.zero()
        header = f"   [hh:mm:ss]  {'rel. error':<22}"
        if var.sslsolver:
            header += f"{'solver':<20}"
            if var.cycle:
                header += f"{'MG':<11} l s"
            var.cprint(header + "\n", 3)
        elif var.cycle:
            var.cprint(header + f"{'[abs. error, last/prev]':>29}   l s\n", 3)
        if var.sslsolver:
            solver._krylov(None, self.s, self.e, var, ops=self._Ops(self))
        elif var.cycle:
            self.multigrid(var)
        self.check_transport()
        var.do_return = False
        return solver._finish(var, None, info)


# =========================================================================== #
# One API for one and many GPUs: emg3d_b200.solve(..., comm= / n_gpus=)
# =========================================================================== #

def solve_distributed(model, sfield, comm, sslsolver=True, semicoarsening=True,
                      linerelaxation=True, verb=0, efield=None, order=None, always_return=False,
                      return_info=False, exact=None, receivers=None, receiver_method='cubic',
                      return_field=True, dist_solver=None, **kwargs):
    """What ``emg3d_b200.solve(model, sfield, ..., comm=comm)`` runs: COLLECTIVE over the ranks of
    ``comm`` (one process per GPU; every rank passes the same global model, source field and
    options).  Same arguments, log, info dict and return conventions as the single-GPU solve;
    every rank receives the whole field -- unless ``return_field=False``: with ``receivers`` the
    field is gathered on rank 0's GPU over NVLink, sampled there and only the responses are
    handed to every rank.  ``dist_solver``: a live :class:`DistributedMultigrid` of this model,
    frequency and option set to reuse (its coefficients, hierarchies and factorisations stay on
    the GPUs between solves, like a single-GPU ``Workspace``); it is not closed."""
    from emg3d_b200 import _lib, fields
    if kwargs.pop('plain', False):
        sslsolver = False if sslsolver is True else sslsolver
        semicoarsening = False if semicoarsening is True else semicoarsening
        linerelaxation = False if linerelaxation is True else linerelaxation
    kwargs.pop('workspace', None)
    if return_field not in (True, False):
        raise ValueError("distributed solve: return_field must be True or False")
    if sfield.frequency is None and getattr(sfield, '_frequency', None) is None:
        raise ValueError("Source field is missing frequency information.")
    if dist_solver is None:
        dmg = DistributedMultigrid(model, sfield, comm, order=order, exact=exact,
                                   semicoarsening=semicoarsening, linerelaxation=linerelaxation)
    else:
        dmg = dist_solver
        dmg.upload_source(sfield)
    try:
        do_return = efield is None or always_return
        if efield is not None:
            if dmg.dtype != efield.field.dtype:
                raise ValueError(
                    "Source field and electric field must have the same "
                    "dtype; complex (f-domain) or real (s-domain). Provided:"
                    f"sfield: {dmg.dtype}; efield: {efield.field.dtype}.")
            dmg.upload_field(efield)
            # PEC: tangential edges on the six GLOBAL boundary faces are zero (solver.py:350-355);
            # the local z-boundaries of inner ranks are halo planes, refreshed by the solve
            _lib.check(_lib.load().emg3d_b200_pec_zero(dmg.level0.lv.handle.ptr, dmg.e.ptr))
        info = dmg.solve(sslsolver=sslsolver, semicoarsening=semicoarsening,
                         linerelaxation=linerelaxation, verb=verb, zero_start=efield is None,
                         **kwargs)
        responses = None
        if receivers is not None:
            # the spline prefilter of the cubic interpolation runs along whole grid lines: the
            # field is gathered on ONE device (NVLink), sampled there, the responses shared
            full = dmg.gather_to_root()
            coords = np.broadcast_arrays(*[np.atleast_1d(np.asarray(c, dtype=float)) for c in
                                           fields.receiver_coordinates(receivers)[:3]])
            d_r = _lib.DeviceArray(2 * coords[0].size, np.float64, scratch=True)
            d_r.zero()
            if full is not None:
                resp = fields.get_receiver(fields.DeviceField(model.grid, full, dmg.dtype, sfield._frequency),
                                           receivers, receiver_method)
                d_r.upload(np.ascontiguousarray(resp.ravel(), dtype=np.complex128).view(np.float64))
                del full
            comm.allreduce_sum(d_r, d_r.size)
            responses = d_r.download().view(np.complex128).reshape(coords[0].shape)
            if dmg.dtype.kind != 'c':
                responses = responses.real
        if return_field:
            # assemble the field on every rank: owned parts into a zeroed global device array, summed
            n = int(model.grid.n_edges)
            full = _lib.DeviceArray(n, dmg.dtype)
            full.zero()
            isz = dmg.dtype.itemsize
            copies, _, _ = gather_plan(dmg.part0, 0, dmg.rank, dmg.gshape[0], dmg.gshape[1])
            lib = _lib.load()
            for loff, goff, cnt in copies:
                _lib.check(lib.emg3d_b200_d2d(full.ptr + goff * isz, dmg.e.ptr + loff * isz, cnt * isz))
            comm.allreduce_sum(full, n * (2 if dmg.dtype.kind == 'c' else 1))
            if efield is None:
                efield = fields.Field(model.grid, dtype=dmg.dtype, frequency=sfield._frequency)
            elif efield.frequency is None:
                efield._frequency = sfield._frequency
            full.download(out=np.asarray(efield.field).view(np.ndarray))
        else:
            efield = None
    finally:
        if dist_solver is None:
            dmg.close()
    out = []
    if do_return and efield is not None:
        out.append(efield)
    if responses is not None:
        out.append(responses)
    if return_info:
        out.append(info)
    if not out:
        return None
    return out[0] if len(out) == 1 else tuple(out)


def _spawned_rank(rank, nranks, uid, model, field, frequency, kwargs, queue):
    """Worker of :func:`solve_spawn`: one rank of one distributed solve."""
    try:
        import emg3d_b200 as eb
        from emg3d_b200 import _lib
        _lib.init(rank)
        comm = NcclComm(rank, nranks, lambda obj: uid)
        sfield = eb.Field(model.grid, field, frequency=frequency)
        kw = dict(kwargs)
        start = kw.pop('efield_array', None)
        if start is not None:
            kw['efield'] = eb.Field(model.grid, start, frequency=frequency)
            kw['always_return'] = True
        kw['return_info'] = True
        out = solve_distributed(model, sfield, comm, **kw)
        comm.destroy()
        if rank == 0:
            queue.put(('ok', np.asarray(out[0].field), out[1]))
    except BaseException as err:            # noqa: BLE001 -- report, the parent re-raises
        import traceback
        queue.put(('error', rank, traceback.format_exc()))
        raise


def solve_spawn(model, sfield, n_gpus, efield=None, return_info=False, always_return=False,
                **kwargs):
    """What ``emg3d_b200.solve(model, sfield, ..., n_gpus=N)`` runs from a single process: spawns
    one rank per GPU (devices 0 .. N-1), runs :func:`solve_distributed` in them and returns rank
    0's result.  Convenience mode: model and fields are pickled to the workers; inside a
    multi-process launch (torchrun, mpirun) pass ``comm=`` instead."""
    import ctypes
    import multiprocessing as mp
    from emg3d_b200 import _lib, fields
    uid = ctypes.create_string_buffer(128)
    _lib.check(_lib.load().emg3d_b200_comm_unique_id(uid))
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    kw = dict(kwargs)
    if efield is not None:
        kw['efield_array'] = np.asarray(efield.field)
    procs = [ctx.Process(target=_spawned_rank,
                         args=(r, int(n_gpus), bytes(uid.raw), model, np.asarray(sfield.field),
                               sfield._frequency, kw, queue)) for r in range(int(n_gpus))]
    for pr in procs:
        pr.start()
    msg = queue.get()
    for pr in procs:
        pr.join()
    if msg[0] != 'ok':
        raise _lib.Emg3dB200Error(f"distributed solve failed on rank {msg[1]}:\n{msg[2]}")
    _, arr, info = msg
    if efield is not None:
        np.asarray(efield.field).view(np.ndarray)[:] = arr
        out = efield
    else:
        out = fields.Field(model.grid, arr, frequency=sfield._frequency)
    do_return = efield is None or always_return
    if do_return and return_info:
        return out, info
    if do_return:
        return out
    if return_info:
        return info

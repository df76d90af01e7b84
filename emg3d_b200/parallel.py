"""Multi-GPU multigrid by z-slab decomposition.

New functionality relative to the reference (which only runs independent solves
in a process pool, emg3d/_multiprocessing.py:33-65): ONE solve is spread over N
GPUs, one process per GPU.

* **Partition.**  z is the slowest axis of every array, so a z-slab of a field or
  coefficient array is one contiguous range.  Node planes are owned in contiguous
  blocks whose boundaries are multiples of ``2**n_dist``, so that ownership
  coarsens consistently over the ``n_dist`` distributed levels.  Every rank works
  on a *local grid*: its owned planes, one halo plane above and, below, as many
  halo planes as are needed to keep the local grid aligned with its own
  coarsening (``2**(n_dist - level)``); the unchanged single-GPU kernels then run
  on local grids, local boundaries playing the role of the PEC boundary.
* **Halo exchange between smoothing sweeps** (the variant ``north_star`` names):
  after every Gauss-Seidel sweep, residual evaluation and prolongation the owner's
  planes next to an interface are sent to the neighbour with ncclSend/ncclRecv
  pairs in one NCCL group (``emg3d_b200_comm_sendrecv``), GPU to GPU over NVLink.
  Nodes at an interface therefore see neighbour values that are at most one sweep
  old (block-Jacobi across slabs, Gauss-Seidel inside).
* **Norms** are sums over owned edges, all-reduced with NCCL.
* **Coarse levels are replicated**: at level ``n_dist`` the restricted residual is
  all-gathered (NCCL point-to-point into the global layout) and every rank runs
  the remaining coarse sub-cycle redundantly with the single-GPU driver; each rank
  then keeps its slab of the correction -- no scatter is needed.

Supported in this mode: standard coarsening (``semicoarsening=False``), point
smoother or x/y line relaxation (``linerelaxation`` in {0, 1, 2, 6}), V/W/F cycles,
no Krylov wrapper.  z-lines cross slabs and are not distributed yet.

The index arithmetic (:class:`SlabPartition`, :func:`exchange_plan`,
:func:`gather_plan`) is pure Python and is tested on CPU with two ``gloo`` ranks;
the transport is pluggable (:class:`NcclComm` on device pointers,
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

    ``nz``: global number of cells along z on the finest grid.  Level ``l`` has
    ``nz >> l`` cells.  Levels ``0 .. n_dist - 1`` are distributed, level
    ``n_dist`` is the first replicated one (slabs exist there only to gather).
    """

    def __init__(self, nz, nranks, n_dist):
        self.nz, self.nranks, self.n_dist = int(nz), int(nranks), int(n_dist)
        align = 1 << self.n_dist
        if self.nz % align:
            raise ValueError(f"nz={nz} must be a multiple of 2**n_dist={align}")
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
        return self.nz >> level

    def bounds(self, level):
        """Owned plane ranges [B[r], B[r+1]) on ``level`` (plane 0 .. nz_level)."""
        inner = [b >> level for b in self.bounds0[1:-1]]
        return [0] + inner + [self.nz_level(level) + 1]

    def owned(self, level, rank):
        b = self.bounds(level)
        return b[rank], b[rank + 1]

    def local(self, level, rank):
        """First and last node plane (inclusive) of the rank's local grid."""
        b = self.bounds(level)
        depth = 1 << (self.n_dist - level)           # halo planes below
        lo = 0 if rank == 0 else b[rank] - depth
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
        depth_up = 1 << (part.n_dist - level)         # halo depth of the upper neighbour
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

    def p2p_exchange(self, slot, args):
        if args[0]:
            self._lib.check(self._lib.load().emg3d_b200_p2p_exchange(slot, *args))

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
        self.owned = owned_ranges(part, level, rank, nx, ny)
        # Smoother and residual run on the z-window [p0 - 1, hi] of the local grid:
        # one halo plane on either side, refreshed by the exchange and fixed during a
        # sweep.  The deeper halo planes below only exist to keep the local grids
        # nested under coarsening; no kernel result ever depends on them.
        lo, hi = part.local(level, rank)
        p0, _ = part.owned(level, rank)
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


class DistributedMultigrid:
    """Plain multigrid cycles of one solve on N GPUs (one instance per rank).

    Parameters
    ----------
    model : Model
        The GLOBAL model (every rank holds it on the host; only the rank's slab
        is uploaded).
    sfield : Field
        The GLOBAL source field.
    comm : NcclComm
    n_dist : int, optional
        Number of distributed levels; default: the levels with more than about a
        million cells (the coarser ones are replicated on every GPU).
    """

    def __init__(self, model, sfield, comm, n_dist=None, order=None, exact=None):
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
        self.gshape = tuple(model.grid.shape_cells)
        nx, ny, nz = self.gshape
        if n_dist is None:
            # distribute the levels that are worth it (more than ~1 M cells: below that
            # a level is launch-latency bound and replicating it is cheaper than
            # exchanging its halos), as long as every rank keeps two owned planes
            n_dist = 1
            while (nz % (1 << (n_dist + 1)) == 0 and (nz >> (n_dist + 1)) // self.nranks >= 2
                   and min(nx, ny) >> (n_dist + 1) >= 2
                   and (nx * ny * nz) >> (3 * n_dist) > 1_000_000):
                n_dist += 1
        self.n_dist = n_dist
        self.part = part = SlabPartition(nz, self.nranks, n_dist)
        self.dtype = np.dtype(np.asarray(sfield.field).dtype)
        self.frequency = sfield._frequency

        # --- local finest level: sliced model, device-side VolumeModel -----------------
        lo, hi = part.local(0, self.rank)
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
        self.levels = [_DLevel(0, lv0, part, self.rank)]
        for l in range(1, n_dist + 1):               # level n_dist: gather buffer only
            self.levels.append(_DLevel(l, self.levels[-1].lv.coarse(0), part, self.rank))

        # --- replicated coarse hierarchy: global level `n_dist` ----------------------
        self.glevel = self._build_global_level(model, n_dist)
        gl = self.glevel
        self.g_s, self.g_e = gl.new_field(), gl.new_field()
        top = self.levels[n_dist]
        tnx, tny = top.lv.shape[0], top.lv.shape[1]
        self._gather = gather_plan(part, n_dist, self.rank, tnx, tny)
        self._scatter = scatter_ranges(part, n_dist, self.rank, tnx, tny)

        # --- local source and field ----------------------------------------------------
        self.s = lv0.new_field()
        self.e = lv0.new_field()
        self.upload_source(sfield)
        self._sums = _lib.DeviceArray(8, np.float64)
        self._sums.zero()

    def close(self):
        """Release the peer-memory mappings of this solver's arrays.  Call on every rank (and
        synchronise the ranks) before the instance is dropped and another one is created."""
        self._lib.sync()
        if hasattr(self.comm, 'p2p_release'):
            self.comm.p2p_release()
        self._slots.clear()
        self._graphs.clear()

    # ---- setup helpers ------------------------------------------------------------
    def _build_global_level(self, model, level):
        """First replicated level: all-gather the owned cell layers of the local
        coarse coefficient slabs (device to device) into global arrays."""
        from emg3d_b200 import meshes, solver
        _lib = self._lib
        top = self.levels[level].lv
        nx, ny = top.shape[0], top.shape[1]
        nzg = self.part.nz_level(level)
        f = 1 << level
        g = model.grid
        ch = [np.add.reduceat(np.asarray(h, dtype=float), np.arange(0, len(h), f)) for h in g.h]
        cgrid = meshes.BaseMesh(ch, g.origin)
        nxy = nx * ny

        def layers(r):                               # owned cell layers [l0, l1) of rank r
            p0, p1 = self.part.owned(level, r)
            return max(p0 - 1, 0), p1 - 1

        lo, _ = self.part.local(level, self.rank)

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

    def _slab(self, field_1d, level=0):
        """Local slab (with halos) of a global host field."""
        nx, ny = self.gshape[0] >> level, self.gshape[1] >> level
        out = np.empty(self.levels[level].lv.n_edges, dtype=self.dtype)
        for goff, loff, n in scatter_ranges(self.part, level, self.rank, nx, ny):
            out[loff:loff + n] = field_1d[goff:goff + n]
        return out

    def upload_source(self, sfield):
        self.s.upload(self._slab(np.asarray(sfield.field)))

    def download_owned(self, out_global):
        """Write the owned part of the local field into a global host array."""
        loc = self.e.download()
        nx, ny = self.gshape[0], self.gshape[1]
        copies, _, _ = gather_plan(self.part, 0, self.rank, nx, ny)
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

    def smoothing(self, dl, s, e, nu, lr_dir):
        """``nu`` sweeps per line direction on the rank's window.

        Exact variant (default): every multicolour sweep is run as its two z-halves
        (the classes of even, then odd z-parity, or the reverse in a descending sweep) with a
        halo exchange after each.  Planes on either side of an interface belong to different
        halves, so every relaxed node sees current neighbour values: a true multicolour
        Gauss-Seidel sweep, as on one GPU.  Relaxed variant (``exact=False``, what
        ``north_star`` words): one exchange per sweep, interface planes see values one sweep
        old (block-Jacobi across slabs) -- half the messages, weaker smoothing at interfaces.
        """
        solver, lib = self._solver, self._lib.load()
        c_lr_dir = int(solver._current_lr_dir(lr_dir, dl.lv.grid))
        dirs = solver._LR_DIRS[c_lr_dir] or (0,)
        for ldir in dirs:
            if ldir == 3:
                raise NotImplementedError("z-line relaxation across z-slabs is not distributed")
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

    # ---- the cycle ---------------------------------------------------------------------
    def multigrid(self, var, level=0, new_cycmax=0):
        """Distributed counterpart of solver._multigrid (same control flow)."""
        solver, lib = self._solver, self._lib.load()
        dl = self.levels[level]
        s = self.s if level == 0 else dl.lv.s
        e = self.e if level == 0 else dl.lv.e
        it = 0
        if new_cycmax == 0 or var.cycle != 'F':
            cycmax = var.cycmax
        else:
            cycmax = new_cycmax
        cyc = 0
        if level == 0 and getattr(var, 'e_is_zero', False):
            l2_last = float(var.s_norm)              # zero start field: ||r|| = ||s||
            var.e_is_zero = False
        else:
            l2_last = self.residual(dl, s, e, norm=True) if level == 0 else 0.0
        l2_stag = np.ones(var.maxcycle) * l2_last
        if level == 0 and var.nu_init > 0:
            self.smoothing(dl, s, e, var.nu_init, var.lr_dir)
        while level == 0 or it < cycmax:
            l2_prev = l2_last
            l2_stag[(it - 1) % var.maxcycle] = l2_last
            if var.nu_pre > 0:
                self.smoothing(dl, s, e, var.nu_pre, var.lr_dir)
            res = self.residual(dl, s, e)
            child = self.levels[level + 1]
            self._lib.check(lib.emg3d_b200_restrict(child.lv.handle.ptr, res.ptr, child.lv.s.ptr))
            child.lv.e.zero()
            self._descend(var, child, level + 1, cycmax - cyc)
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
                if var.lr_cycle:                         # as solver._multigrid (solver.py:639-642)
                    var.lr_dir = next(var.lr_cycle)
                if solver._terminate(var, l2_last, l2_stag[(it - 1) % var.maxcycle], it):
                    break
        var.l2 = l2_last

    def _descend(self, var, child, level, new_cycmax):
        """Everything between restriction to and prolongation from `child`.

        Below the finest level a visit is a fixed sequence of launches (smoothers,
        transfer kernels, halo-exchange kernels, the gather and the replicated coarse
        sub-cycle; no norms, no host decisions), so the visit of level 1 is captured
        into ONE CUDA graph per rank and replayed: the first visit runs eagerly (it
        builds caches and registers the exchanged arrays with the neighbours), the
        second is captured.  The exchange kernels keep their sequence number in device
        memory.  OFF by default (EMG3D_B200_DIST_GRAPHS=1 enables it): with the NCCL
        gather of the replicated levels inside the captured region the replay hung on
        2 B200s in r1; until the gather runs over peer memory as well, the distributed
        levels are launched eagerly (the replicated coarse levels still replay the
        single-GPU graphs of solver._subcycle).
        """
        def run():
            if level < self.n_dist:
                self.multigrid(var, level, new_cycmax)
                self.exchange(child, child.lv.e)
            else:
                self._coarse_replicated(var, child, level, new_cycmax)

        solver = self._solver
        import os
        if not (level == 1 and solver.GRAPHS and os.environ.get('EMG3D_B200_DIST_GRAPHS', '0') == '1'
                and var.verb <= 3
                and not getattr(var, '_capturing', False)):
            return run()
        key = (int(new_cycmax), var.cycle, int(var.lr_dir), solver._order(var), int(var.nu_pre),
               int(var.nu_post), int(var.nu_coarse), tuple(var.clevel))
        g = self._graphs.get(key)
        if g is None:
            self._graphs[key] = False
            return run()
        if g is False:
            var._capturing = True
            try:
                with self._lib.Graph() as g:
                    run()
            finally:
                var._capturing = False
            self._graphs[key] = g
        g.launch()

    def _coarse_replicated(self, var, top, level, new_cycmax):
        """Gather the coarse source, solve the coarse sub-cycle redundantly, keep our slab."""
        isz = self.dtype.itemsize
        lib = self._lib.load()
        copies, sends, recvs = self._gather
        for loff, goff, n in copies:
            self._lib.check(lib.emg3d_b200_d2d(self.g_s.ptr + goff * isz, top.lv.s.ptr + loff * isz, n * isz))
        self.comm.sendrecv_two(top.lv.s.ptr, self.g_s.ptr, isz, sends, recvs)
        self.g_e.zero()
        self._solver._multigrid(self.glevel, self.g_s, self.g_e, var, level=level,
                                new_cycmax=new_cycmax)
        for goff, loff, n in self._scatter:
            self._lib.check(lib.emg3d_b200_d2d(top.lv.e.ptr + loff * isz, self.g_e.ptr + goff * isz, n * isz))

    def solve(self, cycle='V', tol=1e-6, maxit=50, nu_init=0, nu_pre=2, nu_coarse=1, nu_post=2,
              linerelaxation=False, verb=0, zero_start=True):
        """Run multigrid cycles; returns the info dict of solver.solve."""
        solver = self._solver
        lr_values = np.atleast_1d(linerelaxation)
        if linerelaxation is True or any(int(v) not in (0, 1, 2, 6) for v in lr_values.ravel()):
            raise ValueError("distributed line relaxation supports 0 (point), 1 (x), 2 (y) and 6 (x and "
                             f"y) only: z-lines cross the z-slabs. Provided: {linerelaxation!r}.")
        var = solver.MGParameters(verb=verb, sslsolver=False, semicoarsening=False,
                                  linerelaxation=linerelaxation, shape_cells=self.gshape,
                                  cycle=cycle, tol=tol, maxit=maxit, nu_init=nu_init,
                                  nu_pre=nu_pre, nu_coarse=nu_coarse, nu_post=nu_post)
        var.order = {0: 'lex', 1: 'color'}[self.order]
        if var.clevel[0] <= self.n_dist:
            raise ValueError("grid too small for the requested number of distributed levels")
        var.l2_refe = float(np.sqrt(self.sum_owned(self.levels[0], self.s).real))
        var.error_at_cycle[0] = var.l2_refe
        if zero_start:
            self.e.zero()
            var.e_is_zero, var.s_norm = True, var.l2_refe
        self.multigrid(var)
        self.check_transport()
        return {'exit': int(var.exit_message != 'CONVERGED'), 'exit_message': var.exit_message,
                'abs_error': var.l2, 'rel_error': var.l2 / var.l2_refe, 'ref_error': var.l2_refe,
                'it_mg': var.it, 'error_at_cycle': var.error_at_cycle,
                'runtime_at_cycle': var.runtime_at_cycle}

#!/usr/bin/env python
"""Benchmark of the emg3d multigrid hot path on B200 (contract: see DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): cells x smoother-sweeps per second on a 256^3 V-cycle,
next to the achieved HBM GB/s of the dominant kernel against the measured peak.

A "step" is one plain V(2,2) multigrid cycle (nu_coarse = 1) on the 256^3
marine CSEM model of BASELINE.json configs[2] (air / sea / VTI sediment /
resistor, stretched grid, 1 Hz x-dipole), inputs resident in HBM.  The work of a
step is  W = sum over levels of cells(level) x sweeps(level) = 76 695 816
cell-sweeps.  ``e2e`` times the public call
``emg3d_b200.solve(model, sfield, plain=True, cycle='V', maxit=1)`` with host
buffers (pinned source/field buffers), host<->device copies included.

``--impl reference`` times the CPU implementation of the same cycle (the oracle
port, one independent cycle per host core like the reference's process pool) on
a bounded sample (64^3 sibling of the same model).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

METRIC = "cell_smoother_sweeps_per_s_256cube_Vcycle"
UNIT = "cell-sweeps/s"


def vcycle_work(shape, nu_pre=2, nu_post=2, nu_coarse=1):
    """W of one plain V-cycle: cells x sweeps summed over the levels."""
    n = list(shape)
    work = 0
    while True:
        cells = n[0] * n[1] * n[2]
        if any(m % 2 or m <= 2 for m in n):
            # coarsest reachable grid (standard coarsening halves all axes together)
            work += cells * nu_coarse
            break
        work += cells * (nu_pre + nu_post)
        n = [m // 2 for m in n]
    return work


def workload_config(size, world, order):
    """`config` of the JSON line: the workload, a function of the command line only -- both arms
    (`--impl b200` and `--impl reference`) print the same dict.  Run-time facts (number of
    distributed levels, halo transport, the CPU sample) go to other keys."""
    from emg3d_b200 import recipes
    if world == 1:
        shape = (size, size, size)
        what = (f"plain V(2,2) multigrid cycle (nu_coarse=1) on the {size}^3 marine CSEM model of "
                "BASELINE.json configs[2], complex128, VTI")
        par = "single GPU"
    else:
        shape = tuple(recipes.bench_shape(world, size))
        what = ("plain V(2,2) multigrid cycle (nu_coarse=1) on the marine CSEM model of BASELINE.json "
                f"configs[2] grown to {shape[0]}x{shape[1]}x{shape[2]} cells ({size}^3 per GPU), "
                "complex128, VTI")
        par = (f"one solve on {world} GPUs: z-slabs of {shape[2] // world} cell layers, the levels with "
               "more than ~1 M cells distributed, coarser levels replicated; halo exchange of E after "
               "every half sweep (a true Gauss-Seidel sweep across slabs), residual and prolongation")
    work = vcycle_work(shape)
    return {"workload": f"{what}; {work} cell-sweeps per step", "order": order,
            "cells": int(np.prod(shape)), "cell_sweeps_per_step": int(work),
            "l2": "working set 3.2 GB per GPU and step, far larger than the 126 MB L2",
            "parallelism": par}


def peaks():
    fn = os.path.join(HERE, 'MEASURED_PEAKS.json')
    if os.path.exists(fn):
        with open(fn) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (rank 0)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.device}', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def recorded_traffic(kernel_prefix):
    """DRAM bytes per launch (read + write, mean over the captured launches) of a kernel from the
    committed ncu capture of the shipped kernels (profiles/r2_ncu_full_summary.csv, `ncu --set
    full`, 256^3 workload); None if absent.  bench.py never runs under a profiler, so this is a
    recorded, not a live, number."""
    import csv
    path = os.path.join(HERE, 'profiles', 'r2_ncu_full_summary.csv')
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    try:
        rows = [r for r in csv.reader(ln for ln in open(path) if not ln.startswith('#'))]
        names = next(r for r in rows if r[0] == 'Kernel Name')
        cols = [i for i, n in enumerate(names) if i >= 2 and kernel_prefix in n]
        total = 0.0
        for want in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            r = next(r for r in rows if r[0] == want)
            total += sum(float(r[i]) for i in cols) / len(cols) * unit[r[1]]
        return total
    except (OSError, StopIteration, ValueError, KeyError, IndexError, ZeroDivisionError):
        return None


def build_inputs(n):
    import emg3d_b200 as eb
    from emg3d_b200 import recipes
    cfg = recipes.config('config3', n)
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    model = eb.Model(grid, **cfg['model'])
    sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
    return cfg, grid, model, sfield


# --------------------------------------------------------------------------- #
# CPU legs (oracle port)
# --------------------------------------------------------------------------- #

def _cpu_cycle(n):
    """One plain V(2,2)-cycle of the oracle on the n^3 sibling; returns (W, seconds)."""
    from oracle import mg
    from emg3d_b200 import recipes
    import emg3d_b200 as eb
    cfg = recipes.config('config3', n)
    g = mg.Grid(cfg['h'], cfg['origin'])
    m = cfg['model']
    vm = mg.VolumeModel(g, m['property_x'], None, m['property_z'], None, None, cfg['frequency'])
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    s = np.asarray(eb.get_source_field(grid, cfg['source'], cfg['frequency']).field)
    t0 = time.perf_counter()
    _, info = mg.solve(vm, s.copy(), cycle='V', maxit=1)
    return info['cell_sweeps'], time.perf_counter() - t0


def cpu_baseline(n=128, repeats=3):
    """One thread of the C oracle port on the n^3 sibling: median of `repeats` cycles."""
    import oracle
    oracle.build()
    oracle.lib()
    _cpu_cycle(16)                                   # warm-up (library load)
    runs = [_cpu_cycle(n) for _ in range(repeats)]
    work = runs[0][0]
    secs = sorted(r[1] for r in runs)
    sec = secs[len(secs) // 2]
    return {"value": work / sec, "unit": UNIT, "cores": 1, "kind": "port",
            "spread": [work / secs[-1], work / secs[0]],
            "sample": f"one plain V(2,2)-cycle of the C oracle on the {n}^3 sibling of the "
                      f"workload ({work} cell-sweeps, median {sec:.1f} s of {repeats} runs, 1 thread)"}


def _pinned_cycle(job):
    """Pool worker: pin to one core (no migration between the repeats), run one cycle."""
    core, n = job
    try:
        os.sched_setaffinity(0, {core})
    except (AttributeError, OSError):
        pass
    return _cpu_cycle(n)


def run_reference(args):
    """--impl reference: the CPU implementation of the path on all host cores.

    The reference's own parallel mode is a process pool of independent solves
    (emg3d/_multiprocessing.py:33-65), one per core; the kernels are single-threaded.  A step
    = one plain V(2,2)-cycle per core, all cores at once, on a bounded sibling of the workload
    (128^3, 96^3, 64^3 ... : the largest that keeps the K + W steps within about four minutes);
    throughput = cell-sweeps of all cores / time of the slowest.  Workers are forked AFTER the
    oracle library is loaded, pinned to one core each and warmed on the same size."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    oracle.lib()                                      # loaded before the fork: workers inherit it
    try:
        cores_avail = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores_avail = list(range(os.cpu_count() or 1))
    cores = len(cores_avail)
    # sibling size: the largest whose run of K + W steps stays within about four minutes (one
    # all-core step costs ~ 12 s at 128^3, ~ 6 s at 96^3 on the boxes seen so far)
    rounds = max(1, args.steps + args.warmup)
    n = next((m for m, sec in ((128, 12.0), (96, 6.0), (64, 2.0), (48, 1.0)) if rounds * sec <= 240), 32)
    if args.size < 128:
        n = min(n, args.size)
    one = cpu_baseline(n, repeats=1)                  # one thread, same size, machine otherwise idle
    jobs = [(c, n) for c in cores_avail]
    per_step = []
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_pinned_cycle, jobs, chunksize=1)
        work, sec = 0, 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_pinned_cycle, jobs, chunksize=1)
            wall = time.perf_counter() - t0
            w = sum(r[0] for r in res)
            work += w
            sec += wall                              # wall clock of the step (dispatch included)
            per_step.append(w / wall)
    value = work / sec
    per_step.sort()
    sample = (f"{cores} concurrent independent plain V(2,2)-cycles (one per host core, the "
              f"reference's process-pool mode) of the C oracle port on the {n}^3 sibling of "
              f"the workload per step, {args.steps} steps after {args.warmup} warm-up steps of the same "
              f"size; wall clock per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.size, args.gpus, args.order),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample, "order": "lex (the reference's order)",
                         "per_core": value / cores,
                         "one_thread_alone": one["value"],
                         "median_step": per_step[len(per_step) // 2],
                         "spread": [per_step[0], per_step[-1]]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# GPU arm
# --------------------------------------------------------------------------- #

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200,
                    help='timed cycles (default 200: a 2 s timed region at ~10 ms per cycle)')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', type=int, default=256, help='cells per axis (256 = the metric)')
    ap.add_argument('--order', default='color', choices=['color', 'lex'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true',
                    help='N > 1: skip the N-GPU against 1-GPU correctness leg')
    ap.add_argument('--cycles-only', action='store_true',
                    help='run the warm-up and timed cycles and stop (for ncu launch lists: the '
                         'launches are then those of the step, plus the one-off setup kernels)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('cpu:gloo,cuda:nccl', device_id=torch.device('cuda', local_rank))

    import emg3d_b200 as eb
    from emg3d_b200 import _lib, solver
    _lib.init(local_rank)

    def barrier():
        _lib.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.size
    if world > 1:
        return run_distributed(args, rank, world, local_rank, dist, barrier, max_over_ranks)
    cfg, grid, model, sfield = build_inputs(n)
    work = vcycle_work(grid.shape_cells)
    vmodel = eb.VolumeModel(model, sfield)
    level = solver._Level.from_volume_model(vmodel, sfield.field.dtype)
    d_s = _lib.DeviceArray.from_host(sfield.field)
    d_e = level.new_field()
    kw = dict(verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
              shape_cells=grid.shape_cells, cycle='V', maxit=1)

    l2_refe = solver._Vec(level.cplx, level.n_edges).norm(d_s)

    def step():
        # what solve(model, sfield, plain=True, cycle='V', maxit=1) runs between the
        # upload of the source and the download of the field
        var = solver.MGParameters(**kw)
        var.order = args.order
        var.l2_refe = l2_refe
        d_e.zero()
        var.e_is_zero, var.s_norm = True, l2_refe      # zero start field: ||r|| = ||b||
        solver._multigrid(level, d_s, d_e, var)
        return var

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    n0 = _lib.launch_count()
    ev0, ev1 = _lib.Event(), _lib.Event()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_ms(ev1))
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    value = world * work * args.steps / (ms * 1e-3)

    if args.cycles_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                              "steps": args.steps, "ms_per_step": ms / args.steps,
                              "gpu_launches": int(launches), "note": "--cycles-only"}))
        return

    # --- dominant kernel: the point smoother on the finest grid -------------------
    # nu = 2 sweeps = 16 colour launches; algorithmic bytes per cell-sweep: E read +
    # write 96, S 48, eta 16 per distinct array, zeta 8.
    n_eta = len({id(a) for a in level.eta})
    bytes_per_cell_sweep = 96 + 48 + 16 * n_eta + 8
    reps = 5
    lib = _lib.load()
    order = solver.core.order_id(args.order)
    for _ in range(2):
        _lib.check(lib.emg3d_b200_gauss_seidel(level.handle.ptr, d_e.ptr, d_s.ptr, 2, 0, order))
    k0, k1 = _lib.Event(), _lib.Event()
    l0 = _lib.launch_count()
    k0.record()
    for _ in range(reps):
        _lib.check(lib.emg3d_b200_gauss_seidel(level.handle.ptr, d_e.ptr, d_s.ptr, 2, 0, order))
    k1.record()
    kms = k0.elapsed_ms(k1)
    klaunch = _lib.launch_count() - l0
    cells = grid.n_cells
    peak, peak_src = peaks()
    bytes_per_launch = bytes_per_cell_sweep * cells * 2 * reps / klaunch
    achieved = bytes_per_launch / (kms * 1e-3 / klaunch) / 1e9
    roofline = {"bound": "hbm", "kernel": "gs_point_tile_kernel (finest grid, one tile-colour launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src,
                "traffic": recorded_traffic('gs_point_tile_kernel') if n == 256 else None,
                "traffic_source": "profiles/r2_ncu_full_summary.csv (ncu --set full of the shipped kernel, same workload, mean of 8 launches); "
                                  "algorithmic bytes per launch: %d" % int(bytes_per_launch),
                "bytes_per_cell_sweep": bytes_per_cell_sweep,
                "launch_ms": kms / klaunch,
                "cell_sweeps_per_s": cells * 2 * reps / (kms * 1e-3)}

    # --- other hot kernels on the finest grid (same accounting) -------------------
    kernels = {}

    def time_call(fn, nrep=3):
        fn()
        a, b = _lib.Event(), _lib.Event()
        a.record()
        for _ in range(nrep):
            fn()
        b.record()
        return a.elapsed_ms(b) / nrep

    if rank == 0:
        for ldir, name in ((1, 'gauss_seidel_x'), (2, 'gauss_seidel_y'), (3, 'gauss_seidel_z')):
            try:
                t = time_call(lambda: _lib.check(lib.emg3d_b200_gauss_seidel(
                    level.handle.ptr, d_e.ptr, d_s.ptr, 2, ldir, order)))
                gbs = bytes_per_cell_sweep * cells * 2 / (t * 1e-3) / 1e9
                kernels[name] = {"ms_nu2": t, "cell_sweeps_per_s": cells * 2 / (t * 1e-3),
                                 "algorithmic_GBs": gbs, "frac": gbs / peak}
            finally:
                level.handle.drop_factors()
        r = level.res_buffer()
        t = time_call(lambda: _lib.check(lib.emg3d_b200_residual(
            level.handle.ptr, d_s.ptr, d_e.ptr, r.ptr, None)))
        gbs = (bytes_per_cell_sweep + 0) * cells / (t * 1e-3) / 1e9
        kernels['residual'] = {"ms": t, "cells_per_s": cells / (t * 1e-3),
                               "algorithmic_GBs": gbs, "frac": gbs / peak}
        # H from E (fields.get_magnetic_field): E read 48 + zeta 8 + H write 48 B/cell
        d_h = _lib.DeviceArray(int(np.prod(grid.shape_cells)) * 3 + sum(
            int(np.prod(grid.shape_cells)) // n for n in grid.shape_cells), sfield.field.dtype)
        t = time_call(lambda: _lib.check(lib.emg3d_b200_magnetic_field(
            level.handle.ptr, d_e.ptr, d_h.ptr, 0.0, -1.0)))
        gbs = 104 * cells / (t * 1e-3) / 1e9
        kernels['magnetic_field'] = {"ms": t, "cells_per_s": cells / (t * 1e-3),
                                     "algorithmic_GBs": gbs, "frac": gbs / peak}
        del d_h

    # --- end to end through the public API with host buffers ----------------------
    # `e2e`: the call a survey makes for one source -- the source is assembled on the host from
    # its coordinates (a handful of non-zero edges: they are the step's host input and cross
    # PCIe as (index, value) pairs), the cycle runs, the responses at a line of receivers are
    # sampled on the device (cubic spline, fields.get_receiver) and only they are read back.
    # `e2e_full_field`: the same call returning the whole field (0.81 GB over PCIe), i.e. what
    # emg3d.solve itself returns.  The model is constant across steps; a Workspace keeps its
    # coefficients and grid hierarchy on the device ("warm"), the first call pays for them.
    e2e = e2e_full = None
    if not args.no_e2e:
        nst = max(1, min(args.steps, 5))
        nbytes_field = d_s.nbytes
        n_prop = sum(getattr(model, k) is not None for k in
                     ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r'))
        del d_e, d_s, level       # the public call allocates its own device buffers
        src, freq = cfg['source'], cfg['frequency']
        nrec = 101
        rec = (np.linspace(0.5 * (grid.nodes_x[0] + src[0]), 0.5 * (grid.nodes_x[-1] + src[0]), nrec),
               float(src[1]), float(src[2]), 0.0, 0.0)
        probe = eb.get_source_field(grid, src, freq)
        n_special = int(probe.sparse[0].size)
        h2d_bytes = n_special * (8 + probe.dtype.itemsize) + 3 * nrec * 8
        base = dict(plain=True, cycle='V', maxit=1, order=args.order, verb=-1)
        # the model does not change between the steps: frozen arrays are trusted by the Workspace
        # without the per-call checksum of every property array (solver.Workspace._digest)
        for k in ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r'):
            if getattr(model, k) is not None:
                getattr(model, k).flags.writeable = False

        def timed(call_kw, reps):
            ws = eb.Workspace(pinned_result=True)
            kw = dict(base, workspace=ws, **call_kw)
            barrier()
            t0 = time.perf_counter()
            out = eb.solve(model, eb.get_source_field(grid, src, freq), **kw)
            _lib.sync()
            cold = time.perf_counter() - t0
            for _ in range(2):
                eb.solve(model, eb.get_source_field(grid, src, freq), **kw)
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = eb.solve(model, eb.get_source_field(grid, src, freq), **kw)
            barrier()
            sec = max_over_ranks(time.perf_counter() - t0)
            ws.clear()
            return out, sec, cold

        resp, sec, cold = timed(dict(receivers=rec, return_field=False), nst)
        assert resp.shape == (nrec,) and np.all(np.isfinite(resp)) and float(np.abs(resp).max()) > 0
        e2e = {"value": world * work * nst / sec, "unit": UNIT, "steps": nst,
               "ms_per_step": 1e3 * sec / nst,
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(nrec * probe.dtype.itemsize),
               "cold_first_call_ms": 1e3 * cold,
               "cold_h2d_bytes": int(n_prop * 8 * cells),
               "note": "per step: the source field is assembled on the host from its coordinates "
                       f"({n_special} non-zero edges) and sent as (index, value) pairs, the cycle runs, "
                       f"{nrec} receiver responses are sampled on the device and read back; model "
                       "coefficients and grid hierarchy stay on the device between steps (Workspace), "
                       "cold_first_call_ms includes their upload and construction",
               "call": "emg3d_b200.solve(model, get_source_field(grid, src, f), plain=True, cycle='V', "
                       "maxit=1, receivers=rec, return_field=False, workspace=ws)"}
        efield, sec, cold = timed({}, nst)
        assert float(np.abs(efield.field[::1000]).max()) > 0
        e2e_full = {"value": world * work * nst / sec, "unit": UNIT, "steps": nst,
                    "ms_per_step": 1e3 * sec / nst, "h2d_bytes_per_step": int(h2d_bytes - 3 * nrec * 8),
                    "d2h_bytes_per_step": int(nbytes_field), "cold_first_call_ms": 1e3 * cold,
                    "call": "emg3d_b200.solve(model, get_source_field(grid, src, f), plain=True, cycle='V', "
                            "maxit=1, workspace=ws)  # the whole field into a pinned host buffer"}
        del efield

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(128 if n >= 128 else n)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(n, 1, args.order),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
            "e2e_full_field": e2e_full,
            "gpu_launches": int(launches), "clocks": clocks,
            "device": _lib.device_name(),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def distributed_parity(comm, rank, world, dist, order, n=128, cycle='F', tol=1e-9):
    """Cheap correctness leg of the N > 1 arm: the stretched triaxial model of BASELINE.json
    configs[4] on n^3 cells, solved by the SAME distributed driver on N GPUs and by the single-GPU
    solver on rank 0 (plain multigrid, multicolour order).  Reports the field difference and the
    number of cycles both need to reach 1e-6; the run fails if the distributed (exact) solve does
    not reach 1e-8, differs from the single-GPU field by more than 1e-7 (both are iterated to
    tol = 1e-9) or needs more than one cycle more to reach 1e-6."""
    import torch
    import emg3d_b200 as eb
    from emg3d_b200 import parallel, recipes
    cfg = recipes.config('config5', n)
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    model = eb.Model(grid, **cfg['model'])
    sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
    out = {}
    for name, exact in (('exact', True), ('relaxed', False)):
        dmg = parallel.DistributedMultigrid(model, sfield, comm, order=order, exact=exact)
        info = dmg.solve(cycle=cycle, tol=tol, maxit=60)
        field = np.zeros(grid.n_edges, dtype=complex)
        dmg.download_owned(field)
        t = torch.from_numpy(field.view(np.float64))
        dist.all_reduce(t)                               # disjoint owned parts: sum = gather (gloo)
        rel = info['error_at_cycle'] / info['ref_error']
        out[name] = dict(field=field, it=int(info['it_mg']), exit=info['exit_message'],
                         final=float(rel[-1]),
                         to_1e6=int(np.argmax(rel < 1e-6)) if (rel < 1e-6).any() else None,
                         err_cycle1=float(rel[1]))
        dmg.close()
        del dmg
        dist.barrier()
    res = None
    if rank == 0:
        e1, i1 = eb.solve(model, sfield, plain=True, cycle=cycle, tol=tol, maxit=60, order=order,
                          return_info=True)
        rel1 = i1['error_at_cycle'] / i1['ref_error']
        single_to = int(np.argmax(rel1 < 1e-6))
        res = {"workload": f"BASELINE.json configs[4] model on {n}^3 cells, plain {cycle}-cycles, tol {tol:g}",
               "single_gpu": {"cycles": int(i1['it_mg']), "cycles_to_1e-6": single_to,
                              "rel_error_after_cycle_1": float(rel1[1])}}
        for name in ('exact', 'relaxed'):
            o = out[name]
            res[name] = {"cycles": o['it'], "cycles_to_1e-6": o['to_1e6'], "exit": o['exit'],
                         "final_rel_error": o['final'],
                         "rel_error_after_cycle_1": o['err_cycle1'],
                         "efield_rel_diff_vs_single_gpu": float(
                             np.linalg.norm(o['field'] - e1.field) / np.linalg.norm(e1.field))}
        ex = res['exact']
        res["ok"] = bool(ex['final_rel_error'] < 1e-8 and ex['efield_rel_diff_vs_single_gpu'] < 1e-7
                         and ex['cycles_to_1e-6'] is not None and ex['cycles_to_1e-6'] <= single_to + 1)
    return res


def run_distributed(args, rank, world, local_rank, dist, barrier, max_over_ranks):
    """N > 1: ONE multigrid solve spread over N GPUs by z-slab decomposition
    (emg3d_b200.parallel), weak scaling: args.size^3 cells per GPU."""
    import emg3d_b200 as eb
    from emg3d_b200 import _lib, parallel, recipes, solver

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = parallel.NcclComm(rank, world, bcast)
    cfg = recipes.bench_grid(world, args.size)
    grid = eb.TensorMesh(cfg['h'], cfg['origin'])
    model = eb.Model(grid, **cfg['model'])
    sfield = eb.get_source_field(grid, cfg['source'], cfg['frequency'])
    shape = tuple(int(v) for v in grid.shape_cells)
    work = vcycle_work(shape)
    dmg = parallel.DistributedMultigrid(model, sfield, comm, order=args.order)
    kw = dict(verb=0, sslsolver=False, semicoarsening=False, linerelaxation=False,
              shape_cells=shape, cycle='V', maxit=1)

    l2_refe = float(np.sqrt(dmg.sum_owned(dmg.levels[0], dmg.s).real))

    def step():
        # same region as the single-GPU step: zero start field, ||r|| = ||b|| known
        var = solver.MGParameters(**kw)
        var.order = args.order
        var.l2_refe = l2_refe
        dmg.e.zero()
        var.e_is_zero, var.s_norm = True, l2_refe
        dmg.multigrid(var)
        return var

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    n0 = _lib.launch_count()
    ev0, ev1 = _lib.Event(), _lib.Event()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_ms(ev1))
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    value = work * args.steps / (ms * 1e-3)

    # dominant kernel: point smoother on this rank's slab of the finest grid
    dl = dmg.levels[0]
    lv = dl.lv
    n_eta = len({id(a) for a in lv.eta})
    bytes_per_cell_sweep = 96 + 48 + 16 * n_eta + 8
    lib = _lib.load()
    order = solver.core.order_id(args.order)
    cells_local = int(np.prod(dl.win.shape))
    reps = 5
    for _ in range(2):
        _lib.check(lib.emg3d_b200_gauss_seidel(dl.win.ptr, dmg.e.ptr, dmg.s.ptr, 2, 0, order))
    k0, k1 = _lib.Event(), _lib.Event()
    l0 = _lib.launch_count()
    k0.record()
    for _ in range(reps):
        _lib.check(lib.emg3d_b200_gauss_seidel(dl.win.ptr, dmg.e.ptr, dmg.s.ptr, 2, 0, order))
    k1.record()
    kms = k0.elapsed_ms(k1)
    klaunch = _lib.launch_count() - l0
    peak, peak_src = peaks()
    bytes_per_launch = bytes_per_cell_sweep * cells_local * 2 * reps / klaunch
    achieved = bytes_per_launch / (kms * 1e-3 / klaunch) / 1e9
    roofline = {"bound": "hbm", "kernel": "gs_point_tile_kernel (rank 0 slab of the finest grid)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": None,
                "bytes_per_cell_sweep": bytes_per_cell_sweep, "launch_ms": kms / klaunch,
                "cell_sweeps_per_s": cells_local * 2 * reps / (kms * 1e-3)}

    # halo-exchange cost on the finest level (one exchange of E), for the record
    for _ in range(2):
        dmg.exchange(dl, dmg.e)
    barrier()          # ranks arrive here at different times; an exchange waits for the neighbours
    x0, x1 = _lib.Event(), _lib.Event()
    x0.record()
    for _ in range(10):
        dmg.exchange(dl, dmg.e)
    x1.record()
    halo_ms = max_over_ranks(x0.elapsed_ms(x1) / 10)
    halo_bytes = sum(cnt for is_send, _, _, cnt in dl.plan if is_send) * dmg.dtype.itemsize

    # end to end through the public API (see the single-GPU arm): per step the source is assembled
    # on the host from its coordinates and every rank sends the non-zero edges of its slab; the
    # cycle runs on the N GPUs; the field is gathered on rank 0's GPU over NVLink, the receiver
    # responses are sampled there and only they come back to the hosts.  `e2e_full_field`: every
    # rank's slab of the field is read back into pinned host memory instead.
    e2e = e2e_full = None
    if not args.no_e2e:
        nloc = lv.n_edges
        nst = max(1, min(args.steps, 5))
        src, freq = cfg['source'], cfg['frequency']
        nrec = 101
        rec = (np.linspace(0.5 * (grid.nodes_x[0] + src[0]), 0.5 * (grid.nodes_x[-1] + src[0]), nrec),
               float(src[1]), float(src[2]), 0.0, 0.0)
        probe = eb.get_source_field(grid, src, freq)
        n_special = int(probe.sparse[0].size)
        call = dict(comm=comm, dist_solver=dmg, plain=True, cycle='V', maxit=1, verb=-1, order=args.order)

        def survey_step():
            return eb.solve(model, eb.get_source_field(grid, src, freq), receivers=rec, return_field=False, **call)

        for _ in range(2):
            resp = survey_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            resp = survey_step()
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0)
        assert resp.shape == (nrec,) and np.all(np.isfinite(resp)) and float(np.abs(resp).max()) > 0
        e2e = {"value": work * nst / sec, "unit": UNIT, "steps": nst, "ms_per_step": 1e3 * sec / nst,
               "h2d_bytes_per_step": int(n_special * (8 + probe.dtype.itemsize) + 3 * nrec * 8),
               "d2h_bytes_per_step": int(world * nrec * 16),
               "call": "emg3d_b200.solve(model, get_source_field(grid, src, f), comm=comm, dist_solver=dmg, "
                       "plain=True, cycle='V', maxit=1, receivers=rec, return_field=False)",
               "note": "per step: source assembled on the host, its non-zero edges sent to the slabs that hold "
                       f"them; cycle on {world} GPUs; field gathered on rank 0's GPU over NVLink, {nrec} receiver "
                       "responses sampled there (cubic spline) and shared; model coefficients, slab hierarchies "
                       "and the NCCL communicator stay alive between steps (dist_solver)"}

        pin_e = _lib.PinnedArray(nloc, dmg.dtype)

        def full_step():
            dmg.upload_source(eb.get_source_field(grid, src, freq))
            info = dmg.solve(cycle='V', maxit=1, verb=-1)
            dmg.e.download(out=pin_e.array)
            return info

        for _ in range(2):
            full_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            full_step()
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0)
        assert float(np.abs(pin_e.array[::1000]).max()) > 0
        import torch
        nb = torch.tensor([float(nloc * dmg.dtype.itemsize)], dtype=torch.float64, device='cuda')
        dist.all_reduce(nb)
        e2e_full = {"value": work * nst / sec, "unit": UNIT, "steps": nst, "ms_per_step": 1e3 * sec / nst,
                    "h2d_bytes_per_step": int(n_special * (8 + probe.dtype.itemsize)),
                    "d2h_bytes_per_step": int(nb.item()),
                    "call": "DistributedMultigrid.upload_source(sparse) + .solve(cycle='V', maxit=1) + the field "
                            "slab of every rank downloaded to pinned host memory"}

    n_dist, dmg_push = dmg.n_dist, bool(dmg.p2p_push)
    dmg.close()
    del dmg
    barrier()
    parity = None if args.no_parity else distributed_parity(comm, rank, world, dist, args.order)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args.size, world, args.order),
            "run": {"distributed_levels": int(n_dist),
                    "halo_transport": ("one peer-memory kernel per exchange (CUDA IPC mapping of the "
                                       "neighbours' slabs, " + ("posted remote stores" if dmg_push else
                                                                "remote loads")
                                       + " over NVLink, flag handshake)"
                                       if comm.p2p else "ncclSend/ncclRecv over NVLink"),
                    "norms": "all-reduced (NCCL)"},
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "e2e_full_field": e2e_full,
            "halo_exchange": {"transport": "peer-memory kernel" if comm.p2p else "nccl", "ms": halo_ms,
                              "bytes_sent_per_rank": int(halo_bytes),
                              "GBs_per_direction": halo_bytes / 2 / (halo_ms * 1e-3) / 1e9 if halo_ms else None},
            "distributed_parity": parity,
            "gpu_launches": int(launches), "clocks": clocks, "device": _lib.device_name(),
        }
        print(json.dumps(line))
    comm.destroy()
    dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        raise SystemExit("distributed parity leg failed: " + json.dumps(parity))


if __name__ == '__main__':
    main()

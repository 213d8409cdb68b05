#!/usr/bin/env python
"""bench.py -- headline benchmark of the ARC hot path on B200 (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W]
    torchrun --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU
    python bench.py --impl reference ...        # CPU reference arm (oracle on host cores)

Metric (BASELINE.json): FP64 eigensolves/s at basis N + radial matrix elements/s.
Primary workload (c3 = BASELINE.json configs[2]): Rb 60S1/2+60S1/2 PairStateInteractions,
dn=5, dl=5, dEmax=25 GHz -> N=2363, k=250 eigenpairs nearest the pair-state energy per R point.
A "step" is one pass of the hot path over one batch of R points (points_per_gpu per rank, weak
scaling: sweep points shard with no data-path collective).

`value`  : device-timed (CUDA events), operators resident in HBM, results left on device.
`e2e`    : the same sweep through the public API (PairStateInteractions.diagonalise) with
           HOST numpy buffers: H2D of matDiagonal/matR/R and D2H of y/highlight/composition
           inside the timed region.
`full_sweeps` : the configurations of BASELINE.json run IN FULL (C1 600 fields, C3 1000 R, C4
           2 x 2000 fields, C5 a fixed 512 of its 4096 R) through defineBasis + diagonalise on
           ALL ranks of the run (strong scaling: the radial table is sharded over the ranks and
           all-gathered over NCCL, the sweep points are sharded and the results all-gathered),
           wall-clock, max over ranks.
Everything that enters a collective runs on every rank; the single-GPU sub-workloads (radial
table C2, documented Shirley example) and the CPU baselines run on rank 0 after the process
group has been destroyed.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # CPU reference arm: one BLAS thread per worker process, set BEFORE numpy loads OpenBLAS
    for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = "1"

import argparse      # noqa: E402
import json          # noqa: E402
import subprocess    # noqa: E402
import threading     # noqa: E402
import time          # noqa: E402

import numpy as np   # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3_WORKLOAD = ("c3: Rb 60S1/2+60S1/2 PairStateInteractions dn=5 dl=5 dEmax=25GHz (BASELINE configs[2]), "
               "N=2363, k=250 eigenpairs nearest the pair state per R point")
L2_BYTES_PER_CLK = 6300.0          # B300_MICROARCH.md "TMA chip-throughput ~6300 B/cyc" (L2 -> SM cap)


# --------------------------------------------------------------------------- helpers
def bench_config(points, world):
    """`config` of the JSON line -- the same dict for this arm and for `--impl reference` (whose bounded
    sample of the workload is described in its `cpu_baseline.sample`)."""
    return {"workload": C3_WORKLOAD, "points_per_step": points,
            "l2": "inputs larger than L2 (the batch workspace of a sweep is tens of GB: top-level workspace_gb)",
            "parallelism": "sweep points sharded, %d rank(s), no data-path collective" % world}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def peaks_raw():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def dist_setup():
    import datetime
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        # a short watchdog: a rank that misses a collective fails the run in minutes, not in
        # the 10 min x N GPUs of the default
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=240))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed_steps(fn, steps, warmup, world):
    """W untimed + exactly K timed steps, barrier+sync on both sides, CUDA events."""
    import torch
    for _ in range(warmup):
        fn()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def free_device_cache():
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()


# --------------------------------------------------------------------------- main workload (weak)
def build_c3(group=None):
    """Rb 60S+60S pair basis through our own defineBasis (GPU radial table; with a process
    group the table is sharded over the ranks and all-gathered over NCCL)."""
    import arc_b200
    atom = arc_b200.Rubidium()
    atom.radial_group = group
    calc = arc_b200.PairStateInteractions(atom, 60, 0, .5, 60, 0, .5, .5, .5)
    calc.defineBasis(0., 0., 5, 5, 25e9)
    return calc


def bench_c3(args, rank, world):
    """Weak-scaling headline: every rank sweeps `points` R points of its own (no collective in
    the data path; the only collectives are the barriers and the max-over-ranks of the timer)."""
    import torch
    from arc_b200 import sweep
    t0 = time.time()
    calc = build_c3(group=None)
    t_define = time.time() - t0
    N = len(calc.basisStates)
    k = 250
    npts = args.points
    allR = np.linspace(1.0, 10.0, 1000)
    # interleaved subsets of the 1000-point grid that each span the whole 1-10 um range
    pick = (np.linspace(0, 999 - (max(world, 1) - 1), npts).astype(int) + rank) % 1000
    myR = np.sort(allR[pick])[::-1].copy()
    ops = calc.device_operators()
    launches = [0]

    def step():
        r = sweep.pair_sweep(ops, myR, k, 0.0, calc.originalPairStateIndex, topk=4,
                             contrib_max=0.95, group=None, to_host=False)
        launches[0] = r.kernel_launches
    with ClockSampler(torch.cuda.current_device()) as clk:
        ms = timed_steps(step, args.steps, args.warmup, world)
    value = npts * world * args.steps / (ms * 1e-3)

    # e2e: public API, host buffers in, host results out, every step
    h2d = d2h = 0

    def step_e2e():
        nonlocal h2d, d2h
        calc._ops = None                       # force re-staging of matDiagonal / matR
        calc.diagonalise(myR, k)               # group=None: this rank's own points
        h2d, d2h = calc.last_sweep.h2d_bytes, calc.last_sweep.d2h_bytes
    for _ in range(max(1, args.warmup // 3)):
        step_e2e()
    barrier(world)
    t1 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier(world)
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t1) * 1e3), world)
    e2e_value = npts * world * args.steps / (ms_e2e * 1e-3)
    batch = sweep._pick_batch(N, npts, k, torch.cuda.current_device())
    from arc_b200 import _lib
    work_bytes = int(_lib.load().arcb200_sweep_workspace_bytes(N, batch, k))
    return {"value": value, "work_bytes": work_bytes, "ms": ms, "e2e_value": e2e_value, "ms_e2e": ms_e2e, "h2d": h2d, "d2h": d2h,
            "launches": launches[0], "clocks": clk.summary(), "N": N, "k": k,
            "t_defineBasis_s": t_define, "batch": batch, "calc": calc, "R": myR}


# --------------------------------------------------------------------------- full sweeps (strong)
def _wall(fn, world):
    """Wall-clock seconds of fn() on all ranks: barrier + sync on both sides, max over ranks."""
    import torch
    barrier(world)
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dt, world)
    return dt, out


def bench_full_sweeps(args, rank, world):
    """BASELINE.json configs in full through the public API on ALL ranks (strong scaling):
    defineBasis (radial table sharded + all-gathered when world > 1) + diagonalise (points
    sharded, results all-gathered).  Every call below is collective for world > 1."""
    import arc_b200
    from arc_b200 import sweep
    G = sweep.WORLD if world > 1 else None
    out = {}

    def record(name, N, npts, t_def, t_diag, launches, extra=None):
        out[name] = {"N": N, "points": npts, "defineBasis_s": round(t_def, 4), "diagonalise_s": round(t_diag, 4),
                     "full_sweep_s": round(t_def + t_diag, 4), "solves_per_s": npts / t_diag,
                     "gpu_launches": launches}
        if extra:
            out[name].update(extra)

    def run(name, fn):
        try:
            fn()
        except Exception as ex:          # same exception on every rank (deterministic inputs)
            out[name] = {"error": repr(ex)[:300]}
        free_device_cache()

    # ---- C1: Rb StarkMap, 600 fields
    def c1():
        atom = arc_b200.Rubidium()
        atom.radial_group = G
        calc = arc_b200.StarkMap(atom)
        t_def, _ = _wall(lambda: calc.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20), world)
        F = np.linspace(0, 600, 600)
        calc.diagonalise(F, group=G)                                    # warm-up
        # a 25-50 ms sweep: one host hiccup is several times the sweep (a 2-GPU run once showed 0.115 s where
        # seven back-to-back calls take 0.026-0.031 s, profiles/r02_c1_2gpu_per_call.log) -> median of three
        ts = [_wall(lambda: calc.diagonalise(F, group=G), world)[0] for _ in range(3)]
        t_diag = sorted(ts)[1]
        record("c1", len(calc.basisStates), len(F), t_def, t_diag, calc.last_sweep.kernel_launches,
               {"diagonalise_s_calls": [round(t, 4) for t in ts]})
        out["_c1_calc"] = calc
    run("c1", c1)

    # ---- C3: Rb 60S+60S, 1000 R, k=250 (C6 extraction from the level diagram on the host)
    def c3():
        t_def, calc = _wall(lambda: build_c3(group=G), world)
        R = np.linspace(1.0, 10.0, 1000)
        t_diag, _ = _wall(lambda: calc.diagonalise(R, 250, group=G), world)
        c6 = None
        try:
            import contextlib
            import io
            # like the reference, the fit uses every distance of a diagonalise call (its range test
            # assumes ascending r), so the van der Waals region is swept on its own
            calc.diagonalise(np.linspace(4.0, 10.0, 74 * max(world, 1)), 250, group=G)
            with contextlib.redirect_stdout(io.StringIO()):
                c6 = float(calc.getC6fromLevelDiagram(4.0, 10.0))
        except Exception:
            c6 = None
        record("c3", len(calc.basisStates), len(R), t_def, t_diag, calc.last_sweep.kernel_launches,
               {"k": 250, "c6_from_level_diagram_GHz_um6": c6})
        calc._ops = None
    run("c3", c3)

    # ---- C4: Sr88 singlet / triplet, 2000 fields each
    def c4():
        for name, j, s_ in (("c4_singlet", 0, 0), ("c4_triplet", 1, 1)):
            atom = arc_b200.Strontium88()
            atom.radial_group = G
            calc = arc_b200.StarkMap(atom)
            t_def, _ = _wall(lambda: calc.defineBasis(60, 0, j, 0, 50, 70, 30, s=s_), world)
            F = np.linspace(0, 200, 2000)
            # warm-up on a slice, like C1's: the first use of a kernel variant in a process pays its lazy module
            # load (0.2 s here, once) -- not part of a sweep's cost
            calc.diagonalise(F[: min(len(F), 300 * max(world, 1))], group=G)      # two batches per rank: same kernel variants
            reps = 3 if s_ == 0 else 1                                            # singlet: 0.1-0.5 s per sweep
            ts = [_wall(lambda: calc.diagonalise(F, group=G), world)[0] for _ in range(reps)]
            t_diag = sorted(ts)[len(ts) // 2]
            record(name, len(calc.basisStates), len(F), t_def, t_diag, calc.last_sweep.kernel_launches,
                   {"diagonalise_s_calls": [round(t, 4) for t in ts]})
            out["_" + name + "_calc"] = calc
            free_device_cache()
    run("c4", c4)

    # ---- C5: Rb 80D5/2+80D5/2, theta=pi/4, Bz=1e-4 T, N=10384; a fixed subset of the 4096 R
    def c5():
        atom = arc_b200.Rubidium()
        atom.radial_group = G
        calc = arc_b200.PairStateInteractions(atom, 80, 2, 2.5, 80, 2, 2.5, 2.5, 2.5)
        t_def, _ = _wall(lambda: calc.defineBasis(np.pi / 4, 0, 3, 4, 10e9, Bz=1e-4), world)
        npts = args.c5_points
        R = np.linspace(2, 20, 4096)[:: 4096 // npts][:npts]
        t_diag, _ = _wall(lambda: calc.diagonalise(R, 250, group=G), world)
        record("c5", len(calc.basisStates), len(R), t_def, t_diag, calc.last_sweep.kernel_launches,
               {"k": 250, "of_points": 4096})
        calc._ops = None
        out["_c5_calc"] = calc
    if args.c5_points > 0:
        run("c5", c5)
    return out


# --------------------------------------------------------------------------- roofline (rank 0)
def load_traffic(tag):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(tag)
        except Exception:
            return None
    return None


def fp64_peaks(dev):
    """DMMA (mma.sync.m8n8k4.f64) and DFMA register-only peaks measured in this run."""
    import ctypes
    import torch
    from arc_b200 import _lib
    lib = _lib.load()
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    tf = ctypes.c_double(0.0)
    _lib.check(lib.arcb200_dmma_peak(dev.index, _lib.stream_ptr(dev.index), _lib.ptr(scratch),
                                     ctypes.addressof(tf)), "dmma_peak")
    ff = ctypes.c_double(0.0)
    try:
        _lib.check(lib.arcb200_dfma_peak(dev.index, _lib.stream_ptr(dev.index), _lib.ptr(scratch),
                                         ctypes.addressof(ff)), "dfma_peak")
    except AttributeError:
        ff.value = float("nan")
    return tf.value, ff.value


def roofline_c3(N, k, npts):
    """The dominant kernel (band reduction, FP64 tensor cores) timed alone on one batch."""
    import torch
    from arc_b200 import sweep, _lib
    hbm, hbm_src = measured_peaks()
    dev = torch.device("cuda", torch.cuda.current_device())
    batch = sweep._pick_batch(N, npts, k, dev.index)
    rng = np.random.default_rng(0)
    nb = min(batch, npts)
    A = rng.standard_normal((nb, N, N))
    A = A + A.transpose(0, 2, 1)
    lib = _lib.load()
    d_A = torch.as_tensor(A).to(dev)
    del A
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    band = torch.zeros((nb, N, 33), dtype=torch.float64, device=dev)

    def sy():
        rc = lib.arcb200_sy2sb_batched(dev.index, _lib.stream_ptr(dev.index), N, nb, _lib.ptr(d_A),
                                       _lib.ptr(band), _lib.ptr(work), wbytes, 0)
        _lib.check(rc, "sy2sb")
    for _ in range(2):
        sy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        sy()
    e1.record()
    torch.cuda.synchronize()
    sy_ms = e0.elapsed_time(e1) / reps
    dmma, dfma = fp64_peaks(dev)
    flops = (4.0 / 3.0) * float(N) ** 3 * nb
    achieved = flops / (sy_ms * 1e-3) / 1e12
    traffic = load_traffic("sy2sb_c3_per_matrix")
    roof = {"bound": "tensor",
            "kernel": "band reduction dense -> band (per panel: sy2sb_qr_kernel + sy2sb_phase_kernel<1,2,3>: panel QR, symmetric product, W, rank-2k update; mma.sync.m8n8k4.f64 fed by mbarrier-guarded cp.async rings), all launches of one arcb200_sy2sb_batched call",
            "achieved": achieved, "peak": dmma, "unit": "TFLOP/s", "frac": achieved / dmma,
            "traffic": traffic * nb if traffic else None,
            "peak_source": "measured in this run: register-only mma.sync.m8n8k4.f64 loop (arcb200_dmma_peak); "
                           "MEASURED_PEAKS.json has no FP64 entry",
            "dfma_tflops": dfma,
            "algorithmic_flop_per_launch": flops, "launch_ms": sy_ms, "matrices_per_launch": nb,
            "hbm_view": {"algorithmic_bytes_per_launch": float(N) ** 3 / 6.0 * nb,
                         "achieved_gbs": float(N) ** 3 / 6.0 * nb / (sy_ms * 1e-3) / 1e9, "peak_gbs": hbm,
                         "peak_source": hbm_src},
            "note": "timed via arcb200_sy2sb_batched (includes one N^2 load pass and the band copy)"}
    del d_A, work, band
    free_device_cache()
    return roof


# --------------------------------------------------------------------------- CPU baselines
def cpu_baseline_c3(calc, nR=4):
    """Reference CPU path for the pair sweep: scipy eigsh shift-invert per R
    (calculations_atom_pairstate.py:3425-3450) via the oracle, on a bounded sample, one core."""
    from oracle import oracle
    R = np.linspace(1.0, 10.0, 1000)[:: max(1, 1000 // max(nR, 1))][:nR]
    with _one_thread():
        t0 = time.perf_counter()
        oracle.pair_diagonalise(calc.matDiagonal, calc.matR, R, 250, calc.originalPairStateIndex)
        dt = time.perf_counter() - t0
    return len(R) / dt, dt


def _one_thread():
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def cpu_legs(full, args):
    """CPU baselines of the full-sweep configurations: the reference's own call per sweep point
    (numpy.linalg.eigh + highlight + composition for Stark maps, scipy eigsh k=250 for pair
    states) on a bounded sample, ONE core (BLAS threads limited to 1)."""
    from oracle import oracle
    res = {}
    with _one_thread():
        calc = full.get("_c1_calc")
        if calc is not None:
            F = np.linspace(0, 600, 600)
            nf = args.cpu_stark_fields
            idx = np.arange(0, 600, 600 // nf)[:nf]
            t0 = time.perf_counter()
            y_o, hl_o, _ = oracle.stark_diagonalise(calc.mat1, calc.mat2, F[idx], calc.indexOfCoupledState)
            dt = time.perf_counter() - t0
            err = float(np.abs(np.array(calc.y)[idx] - y_o).max() / np.abs(y_o).max())
            res["c1"] = {"value": nf / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                         "sample": "%d of the 600 fields: numpy.linalg.eigh + highlight + _stateComposition2" % nf,
                         "max_rel_err_gpu_vs_cpu": err}
        for name, nf in (("c4_singlet", 8), ("c4_triplet", 2)):
            calc = full.get("_" + name + "_calc")
            if calc is None:
                continue
            F = np.linspace(0, 200, 2000)
            idx = np.linspace(0, 1999, nf).astype(int)
            t0 = time.perf_counter()
            y_o, hl_o, _ = oracle.stark_diagonalise(calc.mat1, calc.mat2, F[idx], calc.indexOfCoupledState)
            dt = time.perf_counter() - t0
            err = float(np.abs(np.array(calc.y)[idx] - y_o).max() / np.abs(y_o).max())
            res[name] = {"value": nf / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                         "sample": "%d of the 2000 fields: numpy.linalg.eigh + highlight + _stateComposition2" % nf,
                         "max_rel_err_gpu_vs_cpu": err}
        calc = full.get("_c5_calc")
        if calc is not None and args.cpu_c5_points > 0:
            # the reference's call (eigsh k=250, shift-invert at the pair-state energy) on 1 R point
            r = np.asarray(calc.r)
            pick = [len(r) // 2]
            t0 = time.perf_counter()
            _r, y_o, _hl, _v = oracle.pair_diagonalise(calc.matDiagonal, calc.matR, r[pick], 250,
                                                       calc.originalPairStateIndex)
            dt = time.perf_counter() - t0
            got = np.asarray(calc.y[pick[0]])
            err = float(np.abs(got - y_o[0]).max() / max(np.abs(y_o[0]).max(), 1e-300))
            res["c5"] = {"value": 1.0 / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                         "sample": "1 R point: scipy.sparse.linalg.eigsh(k=250, sigma=0, tol=1e-6) "
                                   "(calculations_atom_pairstate.py:3444)",
                         "max_abs_diff_gpu_vs_eigsh_over_scale": err}
    return res


# --------------------------------------------------------------------------- single-GPU sub-workloads
def bench_c2(args):
    """BASELINE.json configs[1]: Rb radial dipole+quadrupole table, n=20-120, l<=20."""
    import torch
    import arc_b200
    from arc_b200 import radial
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from dev_radial_bench import c2_workload, c2_blocks
    atom = arc_b200.Rubidium()
    states, pairs = c2_workload(atom)
    recs = np.array([atom._state_record(*s) for s in states])
    times = []
    for it in range(2 + 2):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        rb = radial.RadialBatch(recs)
        e1.record()
        out = rb.integrals_device(pairs)
        e2.record()
        torch.cuda.synchronize()
        times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    t_num = float(np.mean([t[0] for t in times[2:]]))
    t_int = float(np.mean([t[1] for t in times[2:]]))
    lens = rb.lengths.astype(np.int64)
    up = np.minimum(lens[pairs[:, 0]], lens[pairs[:, 1]])
    ptpairs = up.sum()
    l2_bytes = 24.0 * ptpairs          # two psi + one r grid per grid point and pair, served by L1 / L2
    # tensor-core Gram form of the same table (opt-in atom.fast_radial)
    gram = None
    try:
        _st, blocks, nel = c2_blocks(atom)
        tg = []
        for it in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rb.gram_refined_device(blocks, thresh=radial.GRAM_REFINE_THRESH)
            e1.record()
            torch.cuda.synchronize()
            tg.append(e0.elapsed_time(e1))
        gram = {"ms_table": float(np.mean(tg[1:])), "elements": int(nel),
                "value": nel / ((t_num + float(np.mean(tg[1:]))) * 1e-3), "unit": "elements/s"}
    except Exception as ex:
        gram = {"error": repr(ex)[:200]}
    sm_mhz = peaks_raw().get("sm_max_mhz", 1965.0)
    l2_peak = L2_BYTES_PER_CLK * sm_mhz * 1e6 / 1e9
    ach = l2_bytes / (t_int * 1e-3) / 1e9
    # cpu baseline: oracle on a random sample of pairs
    from oracle import oracle
    rng = np.random.default_rng(0)
    sample = rng.choice(len(pairs), size=args.cpu_radial_pairs, replace=False)

    def sp(i):
        n, l, j = states[i]
        d = dict(n=n, l=l, j=j, E_au=atom.getEnergy(n, l, j) / 27.211386245981, alphaC=atom.alphaC,
                 alpha=atom.alpha, Z=atom.Z, mu=(atom.mass - 9.1093837139e-31) / atom.mass)
        for kk in ("a1", "a2", "a3", "a4", "rc"):
            d[kk] = getattr(atom, kk)[l] if l < 4 else 0.0
        return d
    t0 = time.perf_counter()
    worst = 0.0
    got = out[torch.as_tensor(sample, device=out.device)].cpu().numpy()
    for q, pi in enumerate(sample):
        a, b, p = pairs[pi]
        ref = oracle.radial_integral(sp(a), sp(b), int(p), use_ref=oracle.have_ref())
        worst = max(worst, abs(got[q] - ref) / max(abs(ref), 1e-300))
    dt = time.perf_counter() - t0
    return {"workload": "c2: Rb radial dipole+quadrupole table n=20-120 l<=20 (4139 states, %d elements)" % len(pairs),
            "metric": "radial_matrix_elements_per_sec", "value": len(pairs) / ((t_num + t_int) * 1e-3),
            "unit": "elements/s", "ms_numerov": t_num, "ms_integrals": t_int, "gpu_launches": 2,
            "max_rel_err_vs_cpu_sample": worst, "gram_path": gram,
            "roofline": {"bound": "l2", "kernel": "radial_integral_kernel", "achieved": ach, "peak": l2_peak,
                         "unit": "GB/s", "frac": ach / l2_peak, "traffic": load_traffic("radial_integral_c2"),
                         "peak_source": "L2->SM cap 6300 B/clk (B300_MICROARCH.md, measured on B300) x %.0f MHz; the "
                                        "wavefunctions (7.1 GB) are re-read ~450x from L2, HBM sees each once" % sm_mhz,
                         "note": "24 B per grid point per pair (two psi + one r grid); the table is sorted by its "
                                 "lower-energy state, so the warps of a CTA share psi_a / r_a through L1"},
            "cpu_baseline": {"value": len(sample) / dt, "unit": "elements/s", "cores": 1,
                             "kind": "reference" if oracle.have_ref() else "port",
                             "sample": "%d random pairs of the table, 2 Numerov runs + trapezoid each" % len(sample)}}


def bench_shirley(args):
    """The one workload with a published reference timing (doc/AC_Stark_primer.ipynb:597-644,
    SURVEY.md section 6): Rb85 56D5/2 mj=1/2, n=46..66, l<=20 (861 states as printed in the
    notebook), Floquet rank 1 (3 x 861 = 2583),
    100 field amplitudes at 12.3 GHz -- 2 min 13 s wall in the notebook (hardware unstated)."""
    import torch
    import arc_b200
    t0 = time.perf_counter()
    calc = arc_b200.ShirleyMethod(arc_b200.Rubidium85())
    calc.defineBasis(56, 2, 2.5, 0.5, 0, 46, 66, 20)
    calc.defineShirleyHamiltonian(fn=1)
    t_def = time.perf_counter() - t0
    E = np.geomspace(1e0, 1.2e1, 100)
    freq = 12.3e9
    ts = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        calc.diagonalise(E, freq)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    sec = float(np.mean(ts[1:]))
    dim0 = len(calc.basisStates)
    N = 3 * dim0
    # CPU sample: the reference's loop body (dense eigh + transition probabilities) on 2 points
    tar = calc.indexOfCoupledState + dim0
    t0 = time.perf_counter()
    err = 0.0
    for ei in (0, 99):
        Hf = (calc.H0 + calc.dT * freq + calc.B * E[ei]).toarray()
        ev, vec = np.linalg.eigh(Hf)
        tp = (np.abs(vec * vec[tar]) ** 2).reshape((3, dim0, N)).sum(axis=(0, -1))
        err = max(err, float(np.abs(calc.eigs[ei] - ev).max() / np.abs(ev).max()),
                  float(np.abs(calc.transProbs[ei] - tp).max()))
    dt = (time.perf_counter() - t0) / 2
    return {"workload": "shirley: Rb85 56D5/2 mj=1/2 ShirleyMethod fn=1, dim %d x 3 = %d, 100 fields x 1 frequency "
                        "(doc/AC_Stark_primer.ipynb)" % (dim0, N),
            "metric": "fp64_eigensolves_per_sec", "value": 100 / sec, "unit": "solves/s",
            "s_per_sweep_e2e": sec, "defineBasis_s": t_def, "gpu_launches": calc.gpu_kernel_launches,
            "max_err_vs_cpu_sample": err,
            "published_reference": {"value": 100 / 133.0, "unit": "solves/s",
                                    "source": "doc/AC_Stark_primer.ipynb:643-644 (wall 2 min 13 s, hardware unstated)"},
            "cpu_baseline": {"value": 1.0 / dt, "unit": "solves/s", "cores": "default BLAS threads", "kind": "port",
                             "sample": "2 of the 100 points: numpy.linalg.eigh + transition probabilities"}}


# --------------------------------------------------------------------------- reference arm
_REF = {}


def _ref_worker(Rs):
    """Runs in a forked worker: the operators are inherited from the parent (no pickling)."""
    from oracle import oracle
    with _one_thread():
        oracle.pair_diagonalise(_REF["diag"], _REF["matR"], np.asarray(Rs), 250, _REF["opi"])
    return len(Rs)


def run_reference(args):
    """The reference's CPU implementation of the path (scipy eigsh shift-invert per R point,
    calculations_atom_pairstate.py:3444) on the box's host cores: persistent pools of forked
    workers created OUTSIDE the timed region, operators shipped once (fork inheritance), one BLAS
    thread per worker (set before numpy was imported), several R points per worker and step.
    The process count is swept once (untimed) and the best one is used."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    # UNTIMED set-up: the operators of the workload, built WITHOUT the GPU: host logic of
    # defineBasis + radial elements from the oracle (the reference's own C Numerov when
    # oracle/_ref is present), parallel over the host cores.
    import arc_b200
    from oracle.host_radial import patch_atom_with_oracle
    t0 = time.time()
    calc = arc_b200.PairStateInteractions(patch_atom_with_oracle(arc_b200.Rubidium(), procs=cores),
                                          60, 0, .5, 60, 0, .5, .5, .5)
    calc.defineBasis(0., 0., 5, 5, 25e9)
    sys.stderr.write("reference arm: operators built on the host in %.1f s (N=%d)\n"
                     % (time.time() - t0, len(calc.basisStates)))
    _REF.update(diag=calc.matDiagonal, matR=calc.matR, opi=calc.originalPairStateIndex)
    allR = np.linspace(1.0, 10.0, 1000)
    ctx = mp.get_context("fork")

    def one_pass(pool, procs, per):
        pts = allR[np.linspace(0, 999, procs * per).astype(int)]
        chunks = [pts[i::procs] for i in range(procs)]
        t0 = time.perf_counter()
        n = sum(pool.map(_ref_worker, chunks, chunksize=1))
        return n / (time.perf_counter() - t0), n

    cands = sorted({p for p in (8, 16, 32, 64, cores // 2, cores) if 1 <= p <= cores})
    if args.ref_procs:
        cands = [min(cores, args.ref_procs)]
    sweep_log, best, best_rate, pools = {}, None, -1.0, {}
    for p in cands:                                    # ascending
        pools[p] = ctx.Pool(p)
        one_pass(pools[p], p, 1)                       # spawn + first-touch warm-up
        rate, _ = one_pass(pools[p], p, 1)
        sweep_log[str(p)] = round(rate, 3)
        if rate > best_rate:
            best, best_rate = p, rate
        elif rate < 0.9 * best_rate:
            break                                      # past the knee (memory-bound SuperLU solves): stop sweeping
    for p in pools:
        if p != best:
            pools[p].close()
    pool = pools[best]
    per = max(1, args.ref_points_per_proc)
    # bounded sample: the K timed steps are projected from the swept rate and the points per worker are
    # reduced until they fit --ref-max-seconds (the arm has to end within a few minutes for any K)
    while per > 1 and args.steps * best * per / max(best_rate, 1e-9) > args.ref_max_seconds:
        per -= 1
    nR = best * per
    for _ in range(min(args.warmup, 1)):
        one_pass(pool, best, per)
    ts = []
    for _ in range(args.steps):
        rate, n = one_pass(pool, best, per)
        ts.append(n / rate)
    pool.close()
    dt = float(np.mean(ts))
    value = nR / dt
    line = {"metric": "fp64_eigensolves_per_sec", "value": value, "unit": "solves/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (Rb 60S+60S pair Hamiltonian built from quantum defects; no dataset)",
            "config": bench_config(args.points, max(1, args.gpus)), "ref_points_per_step": nR,
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": best, "kind": "port",
                             "sample": "%d R points per step (%d per worker), scipy eigsh shift-invert k=250, %d forked "
                                       "workers x 1 BLAS thread, persistent pool, operators inherited"
                                       % (nR, per, best),
                             "host_cores": cores, "cpu_model": cpu_model(),
                             "procs_sweep_solves_per_s": sweep_log},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3"])
    ap.add_argument("--points", type=int, default=296, help="R points per GPU per step")
    ap.add_argument("--skip-sub", action="store_true", help="skip the full sweeps and the sub-workloads")
    ap.add_argument("--c5-points", type=int, default=512, help="R points of the C5 full-sweep subset (0 = skip)")
    ap.add_argument("--cpu-pair-points", type=int, default=6)
    ap.add_argument("--cpu-radial-pairs", type=int, default=150)
    ap.add_argument("--cpu-stark-fields", type=int, default=100)
    ap.add_argument("--cpu-c5-points", type=int, default=1)
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-points-per-proc", type=int, default=2)
    ap.add_argument("--ref-max-seconds", type=float, default=240.0,
                    help="reference arm: projected wall time of the timed steps above which the sample shrinks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank, world, local = dist_setup()
    r = bench_c3(args, rank, world)
    calc = r["calc"]
    # the full sweeps need workspaces of very different sizes (141 GB for C5): drop the device
    # operators of the main workload, which would pin its cached 68 GB segment
    calc._ops = None
    calc.last_sweep = None
    free_device_cache()
    full = {} if args.skip_sub else bench_full_sweeps(args, rank, world)
    barrier(world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- rank 0 only from here on: no collective is entered below
    roof = roofline_c3(r["N"], r["k"], args.points)
    v, dt = cpu_baseline_c3(calc, nR=args.cpu_pair_points)
    cpu = {"value": v, "unit": "solves/s", "cores": 1, "kind": "port",
           "sample": "%d of the R points, scipy.sparse.linalg.eigsh(k=250, sigma=0, tol=1e-6) per point "
                     "(the reference's call, calculations_atom_pairstate.py:3444), one BLAS thread"
                     % args.cpu_pair_points,
           "host_cores": os.cpu_count(), "cpu_model": cpu_model()}
    sub = {}
    if not args.skip_sub:
        for name, fn in (("radial_c2", bench_c2), ("shirley_doc", bench_shirley)):
            try:
                sub[name] = fn(args)
            except Exception as ex:
                sub[name] = {"error": repr(ex)[:300]}
            free_device_cache()
        try:
            legs = cpu_legs(full, args)
            for name, leg in legs.items():
                if name in full and "error" not in full[name]:
                    full[name]["cpu_baseline"] = leg
        except Exception as ex:
            sub["cpu_legs_error"] = repr(ex)[:300]
    full = {k: v for k, v in full.items() if not k.startswith("_")}

    def g(d, *ks):
        for k in ks:
            d = d.get(k) if isinstance(d, dict) else None
        return None if d is None else (round(d, 4) if isinstance(d, float) else d)
    # short scalars at the END of the line, so that they survive a truncated stdout tail
    summary = {"n_gpus": world,
               "c3_weak_solves_per_s": round(r["value"], 2), "c3_e2e_solves_per_s": round(r["e2e_value"], 2),
               "roofline_frac": round(roof["frac"], 4),
               "c1_solves_per_s": g(full, "c1", "solves_per_s"), "c1_full_sweep_s": g(full, "c1", "full_sweep_s"),
               "c2_elements_per_s": g(sub, "radial_c2", "value"),
               "c2_gram_elements_per_s": g(sub, "radial_c2", "gram_path", "value"),
               "c3_full_sweep_s": g(full, "c3", "full_sweep_s"), "c3_full_solves_per_s": g(full, "c3", "solves_per_s"),
               "c4_singlet_solves_per_s": g(full, "c4_singlet", "solves_per_s"),
               "c4_triplet_solves_per_s": g(full, "c4_triplet", "solves_per_s"),
               "c4_full_sweep_s": (g(full, "c4_singlet", "full_sweep_s") or 0) + (g(full, "c4_triplet", "full_sweep_s") or 0),
               "c5_solves_per_s": g(full, "c5", "solves_per_s"), "c5_full_sweep_s": g(full, "c5", "full_sweep_s"),
               "c5_points": g(full, "c5", "points"),
               "shirley_solves_per_s": g(sub, "shirley_doc", "value"),
               "cpu_c3_1core_solves_per_s": round(v, 3)}
    line = {
        "metric": "fp64_eigensolves_per_sec", "value": r["value"], "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (Rb 60S+60S pair Hamiltonian built from quantum defects by our own defineBasis; no dataset)",
        "config": bench_config(args.points, world),
        "batch_in_flight": r["batch"], "workspace_gb": round(r["work_bytes"] / 1e9, 1),
        "e2e": {"value": r["e2e_value"], "unit": "solves/s", "h2d_bytes_per_step": r["h2d"],
                "d2h_bytes_per_step": r["d2h"], "ms_per_step": r["ms_e2e"] / args.steps,
                "api": "PairStateInteractions.diagonalise(R, 250) with host buffers"},
        "gpu_launches": r["launches"] * args.steps,
        "roofline": roof, "cpu_baseline": cpu, "clocks": r["clocks"],
        "defineBasis_s": r["t_defineBasis_s"],
        "full_sweeps": {"scaling": "strong", "n_gpus": world, **full},
        "sub": sub,
        "summary": summary,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()

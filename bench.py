#!/usr/bin/env python
"""bench.py -- headline benchmark of the ARC hot path on B200 (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2]
    python bench.py --impl reference ...        # CPU reference arm (oracle on host cores)

Metric (BASELINE.json): FP64 eigensolves/s at basis N + radial matrix elements/s.
Primary workload (default c3 = BASELINE.json configs[2]): Rb 60S1/2+60S1/2
PairStateInteractions, dn=5, dl=5, dEmax=25 GHz -> N=2363, k=250 eigenpairs nearest the
pair-state energy per R point.  A "step" is one pass of the hot path over one batch of
R points (points_per_gpu per rank, weak scaling: sweep points shard with no data-path
collective).  The line also carries the radial-table workload (c2 =
configs[1], elements/s) and the Stark-map workload (c1 = configs[0], solves/s) as
sub-objects, each with its own roofline and cpu_baseline.

`value`  : device-timed (CUDA events), operators resident in HBM, results left on device.
`e2e`    : the same sweep through the public API (PairStateInteractions.diagonalise) with
           HOST numpy buffers: H2D of matDiagonal/matR/R and D2H of y/highlight/composition
           inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------- helpers
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


def dist_setup(ngpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
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


# --------------------------------------------------------------------------- workloads
def build_c3(shard_radial=False):
    """Rb 60S+60S pair basis through our own defineBasis (GPU radial table; with several
    ranks the table is sharded and all-gathered over NCCL instead of rebuilt per GPU)."""
    import arc_b200
    atom = arc_b200.Rubidium()
    atom.shard_radial = bool(shard_radial)
    calc = arc_b200.PairStateInteractions(atom, 60, 0, .5, 60, 0, .5, .5, .5)
    calc.defineBasis(0., 0., 5, 5, 25e9)
    return calc


def bench_c3(args, rank, world):
    import torch
    from arc_b200 import sweep, _lib
    t0 = time.time()
    calc = build_c3(shard_radial=world > 1)
    t_define = time.time() - t0
    N = len(calc.basisStates)
    k = 250
    npts = args.points
    allR = np.linspace(1.0, 10.0, 1000)
    # weak scaling: every rank sweeps `npts` R points of the 1-10 um range (its own shard)
    # (interleaved subsets of the 1000-point grid that each span the whole 1-10 um range)
    pick = (np.linspace(0, 999 - (max(world, 1) - 1), npts).astype(int) + rank) % 1000
    myR = np.sort(allR[pick])[::-1].copy()
    ops = calc.device_operators()
    launches = [0]

    def step():
        r = sweep.pair_sweep(ops, myR, k, 0.0, calc.originalPairStateIndex, topk=4,
                             contrib_max=0.95, distributed=False, to_host=False)
        launches[0] = r.kernel_launches
    with ClockSampler(torch.cuda.current_device()) as clk:
        ms = timed_steps(step, args.steps, args.warmup, world)
    value = npts * world * args.steps / (ms * 1e-3)

    # e2e: public API, host buffers in, host results out, every step
    h2d = d2h = 0

    def step_e2e():
        nonlocal h2d, d2h
        calc._ops = None                       # force re-staging of matDiagonal / matR
        calc.diagonalise(myR, k)
        h2d, d2h = calc.last_sweep.h2d_bytes, calc.last_sweep.d2h_bytes
    import arc_b200.sweep as sw
    old = sw.dist_info
    sw.dist_info = lambda group=None: (0, 1)   # every rank sweeps its own points (weak scaling)
    try:
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
    finally:
        sw.dist_info = old
    e2e_value = npts * world * args.steps / (ms_e2e * 1e-3)
    # the use the configuration names ("C6 extraction"): fit of the tracked level of the last
    # end-to-end sweep (host post-pass, calculations_atom_pairstate.py:3978-4112)
    c6_fit = None
    try:
        import contextlib
        import io
        # like the reference, the fit uses every distance of the last diagonalise (its range
        # test assumes ascending r while diagonalise stores r descending), so the van der
        # Waals region is swept on its own
        old = sw.dist_info
        sw.dist_info = lambda group=None: (0, 1)
        try:
            calc.diagonalise(np.linspace(4.0, 10.0, 74), k)
        finally:
            sw.dist_info = old
        with contextlib.redirect_stdout(io.StringIO()):
            c6_fit = float(calc.getC6fromLevelDiagram(4.0, 10.0))
    except Exception:
        c6_fit = None

    # roofline of the dominant kernel: time it alone on one batch
    hbm, hbm_src = measured_peaks()
    batch = sweep._pick_batch(N, npts, k, torch.cuda.current_device())
    rng = np.random.default_rng(0)
    A = rng.standard_normal((min(batch, npts), N, N))
    A = A + A.transpose(0, 2, 1)
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    d_A = torch.as_tensor(A).to(dev)
    nb = d_A.shape[0]
    outs = [torch.zeros((nb, N), dtype=torch.float64, device=dev) for _ in range(3)]
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    del A
    two_stage = os.environ.get("ARCB200_TWO_STAGE", "1") != "0" and 1024 <= N <= 16384

    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    if two_stage:
        # sy2sb_kernel (dense -> band, DMMA): (4/3) N^3 flop per matrix, tensor bound.  The
        # entry point = load pass + sy2sb kernel + band copy (two N^2 passes, < 1 % of the time).
        band = torch.zeros((nb, N, 33), dtype=torch.float64, device=dev)

        def sy():
            rc = lib.arcb200_sy2sb_batched(dev.index, _lib.stream_ptr(dev.index), N, nb, _lib.ptr(d_A),
                                           _lib.ptr(band), _lib.ptr(work), wbytes)
            _lib.check(rc, "sy2sb")
        sy_ms = timed(sy)
        scratch = torch.zeros(8, dtype=torch.float64, device=dev)
        import ctypes
        tf = ctypes.c_double(0.0)
        _lib.check(lib.arcb200_dmma_peak(dev.index, _lib.stream_ptr(dev.index), _lib.ptr(scratch),
                                         ctypes.addressof(tf)), "dmma_peak")
        flops = (4.0 / 3.0) * float(N) ** 3 * nb
        achieved = flops / (sy_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "sy2sb_kernel (dense -> band reduction: panel QR + symmetric product "
                "+ rank-2k update on FP64 tensor cores, mma.sync.m8n8k4.f64)",
                "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
                "traffic": (load_traffic("sy2sb_c3_per_matrix") or 0.0) * nb or None,
                "peak_source": "measured in this run: register-only mma.sync.m8n8k4.f64 loop, 8 warps/SM "
                               "(arcb200_dmma_peak); MEASURED_PEAKS.json has no FP64 figure (bf16 %s TFLOP/s, HBM %s GB/s)"
                               % (peaks_raw().get("bf16_tflops"), hbm),
                "algorithmic_flop_per_launch": flops, "launch_ms": sy_ms, "matrices_per_launch": nb,
                "hbm_view": {"algorithmic_bytes_per_launch": float(N) ** 3 / 6.0 * nb,
                             "achieved_gbs": float(N) ** 3 / 6.0 * nb / (sy_ms * 1e-3) / 1e9, "peak_gbs": hbm,
                             "note": "two reads of the trailing lower triangle (symmetric product) + one "
                                     "read-modify-write (rank-2k update) per 32-column panel = N^3/6 bytes"},
                "note": "timed via arcb200_sy2sb_batched (includes one N^2 load pass and the band copy)"}
        del band
    else:
        def sy():
            rc = lib.arcb200_sytrd_batched(dev.index, _lib.stream_ptr(dev.index), N, nb, _lib.ptr(d_A),
                                           _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]),
                                           _lib.ptr(work), wbytes, 0)
            _lib.check(rc, "sytrd")
        sy_ms = timed(sy)
        alg_bytes = (4.0 / 3.0) * float(N) ** 3 * nb          # (4/3) N^3 bytes per solve
        achieved = alg_bytes / (sy_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "sytrd_kernel (blocked Householder tridiagonalisation, symv sweeps)",
                "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": load_traffic("sytrd_c3"), "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": sy_ms, "matrices_per_launch": nb,
                "note": "timed via arcb200_sytrd_batched (includes one N^2 load pass per matrix)"}
    del d_A, work
    out = {"value": value, "ms": ms, "e2e_value": e2e_value, "ms_e2e": ms_e2e, "h2d": h2d, "d2h": d2h,
           "launches": launches[0], "roofline": roof, "clocks": clk.summary(), "N": N, "k": k,
           "t_defineBasis_s": t_define, "batch": batch, "calc": calc, "R": myR, "c6_fit": c6_fit}
    return out


def load_traffic(tag):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(tag)
        except Exception:
            return None
    return None


def cpu_baseline_c3(calc, nR=4, procs=1):
    """Reference CPU path for the pair sweep: scipy eigsh shift-invert per R
    (calculations_atom_pairstate.py:3425-3450) via the oracle, on a bounded sample."""
    from oracle import oracle
    R = np.linspace(1.0, 10.0, 1000)[:: max(1, 1000 // max(nR, 1))][:nR]
    t0 = time.perf_counter()
    if procs <= 1:
        oracle.pair_diagonalise(calc.matDiagonal, calc.matR, R, 250, calc.originalPairStateIndex)
    else:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        chunks = [c for c in np.array_split(R, procs) if len(c)]
        with ctx.Pool(len(chunks), initializer=_cpu_init) as pool:
            pool.map(_cpu_pair_worker, [(calc.matDiagonal, calc.matR, c, calc.originalPairStateIndex)
                                        for c in chunks])
    dt = time.perf_counter() - t0
    return len(R) / dt, dt


def _cpu_init():
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"


def _cpu_pair_worker(a):
    from oracle import oracle
    oracle.pair_diagonalise(a[0], a[1], a[2], 250, a[3])
    return len(a[2])


def bench_c2(args):
    """BASELINE.json configs[1]: Rb radial dipole+quadrupole table, n=20-120, l<=20."""
    import torch
    import arc_b200
    from arc_b200 import radial
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from dev_radial_bench import c2_workload
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
    ptpairs = np.minimum(lens[pairs[:, 0]], lens[pairs[:, 1]]).sum()
    hbm, src = measured_peaks()
    ach = ptpairs * 24.0 / (t_int * 1e-3) / 1e9
    # cpu baseline: oracle on a stratified sample of pairs
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
            "max_rel_err_vs_cpu_sample": worst,
            "roofline": {"bound": "hbm", "kernel": "radial_integral_kernel", "achieved": ach, "peak": hbm,
                         "unit": "GB/s", "frac": ach / hbm, "traffic": load_traffic("radial_integral_c2"),
                         "peak_source": src,
                         "note": "24 B per grid point per pair (two psi + one r grid); >1 means L2 reuse"},
            "cpu_baseline": {"value": len(sample) / dt, "unit": "elements/s", "cores": 1,
                             "kind": "reference" if oracle.have_ref() else "port",
                             "sample": "%d random pairs of the table, 2 Numerov runs + trapezoid each" % len(sample)}}


def bench_c1(args):
    """BASELINE.json configs[0]: Rb StarkMap 28S, n=23-32, l<=20, 600 fields."""
    import torch
    import arc_b200
    from arc_b200 import sweep
    calc = arc_b200.StarkMap(arc_b200.Rubidium())
    calc.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20)
    F = np.linspace(0, 600, 600)
    h0 = np.diag(calc.mat1).copy()
    N = len(h0)
    res = None
    ts = []
    for it in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        res = sweep.stark_sweep(h0, calc.mat2, F, calc.indexOfCoupledState, distributed=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    ms = float(np.mean(ts[2:]))
    from oracle import oracle
    nf = args.cpu_stark_fields
    t0 = time.perf_counter()
    y_o, hl_o, _ = oracle.stark_diagonalise(calc.mat1, calc.mat2, F[:: 600 // nf][:nf], calc.indexOfCoupledState)
    dt = time.perf_counter() - t0
    err = float(np.abs(res.y[:: 600 // nf][:nf] - y_o).max() / np.abs(y_o).max())
    return {"workload": "c1: Rb StarkMap defineBasis(28,0,0.5,0.5,23,32,20), 600 fields 0-600 V/m, N=%d" % N,
            "metric": "fp64_eigensolves_per_sec", "value": 600 / (ms * 1e-3), "unit": "solves/s",
            "ms_per_sweep_e2e": ms, "gpu_launches": res.kernel_launches, "max_rel_err_vs_cpu_sample": err,
            "note": "timed through sweep.stark_sweep with host buffers (H2D + D2H inside)",
            "cpu_baseline": {"value": nf / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                             "sample": "%d of the 600 fields: numpy.linalg.eigh + highlight + _stateComposition2" % nf}}


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


def bench_c4(args):
    """BASELINE.json configs[3]: Sr88 divalent StarkMap, singlet (s=0) and triplet (s=1) bases,
    n=50-70, l<=30; a bounded number of the 2000 fields per GPU."""
    import torch
    import arc_b200
    out = {}
    nF = 148
    for name, j, s_ in (("singlet", 0, 0), ("triplet", 1, 1)):
        calc = arc_b200.StarkMap(arc_b200.Strontium88())
        t0 = time.perf_counter()
        calc.defineBasis(60, 0, j, 0, 50, 70, 30, s=s_)
        t_def = time.perf_counter() - t0
        F = np.linspace(0, 200, 2000)[:: 2000 // nF][:nF]
        ts = []
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            calc.diagonalise(F)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        sec = float(np.mean(ts[1:]))
        i = nF // 2
        t0 = time.perf_counter()
        w = np.linalg.eigh(calc.mat1 + calc.mat2 * F[i])[0]
        dt = time.perf_counter() - t0
        out[name] = {"N": len(calc.basisStates), "value": nF / sec, "unit": "solves/s", "fields": nF,
                     "s_per_sweep_e2e": sec, "defineBasis_s": t_def,
                     "max_rel_err_vs_cpu_sample": float(np.abs(calc.y[i] - w).max() / np.abs(w).max()),
                     "cpu_baseline": {"value": 1.0 / dt, "unit": "solves/s", "cores": "default BLAS threads",
                                      "kind": "port", "sample": "1 field: numpy.linalg.eigh"}}
    return {"workload": "c4: Sr88 StarkMap defineBasis(60,0,J,0,50,70,30,s=S), %d of 2000 fields 0-200 V/m" % nF,
            "metric": "fp64_eigensolves_per_sec", **out}


def bench_c5(args):
    """BASELINE.json configs[4]: Rb 80D5/2+80D5/2, theta=pi/4, Bz=1e-4 T, N=10384; 32 of the 4096
    R points (one batch: 4.4 GB of workspace per matrix), k=250."""
    import gc
    import torch
    import arc_b200
    gc.collect()
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    calc = arc_b200.PairStateInteractions(arc_b200.Rubidium(), 80, 2, 2.5, 80, 2, 2.5, 2.5, 2.5)
    calc.defineBasis(np.pi / 4, 0, 3, 4, 10e9, Bz=1e-4)
    t_def = time.perf_counter() - t0
    N = len(calc.basisStates)
    R = np.linspace(2, 20, 4096)[:: 4096 // 32][:32]
    from arc_b200 import sweep as _sw
    free_gb = torch.cuda.mem_get_info()[0] / 1e9
    batch = _sw._pick_batch(N, len(R), 250, torch.cuda.current_device())
    calc.diagonalise(R, 250)              # warm-up: maps the 141 GB workspace
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    calc.diagonalise(R, 250)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    H = calc.matDiagonal.toarray() + calc.matR[0].toarray() / (calc.r[0] * 1e-6) ** 3
    t0 = time.perf_counter()
    w = np.linalg.eigvalsh(H)
    dt = time.perf_counter() - t0
    sel = np.sort(np.argsort(np.abs(w), kind="stable")[:250])
    return {"workload": "c5: Rb 80D5/2+80D5/2 theta=pi/4 Bz=1e-4 T, N=%d, 32 of 4096 R points, k=250" % N,
            "metric": "fp64_eigensolves_per_sec", "value": 32 / sec, "unit": "solves/s", "s_per_point": sec / 32,
            "defineBasis_s": t_def, "gpu_launches": calc.last_sweep.kernel_launches,
            "batch_in_flight": batch, "free_hbm_GB_before": free_gb,
            "max_rel_err_vs_cpu_sample": float(np.abs(calc.y[0] - w[sel]).max() / np.abs(w).max()),
            "cpu_baseline": {"value": 1.0 / dt, "unit": "solves/s", "cores": "default BLAS threads", "kind": "port",
                             "sample": "1 R point: numpy.linalg.eigvalsh of the dense matrix (eigenvalues only)"}}


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # matrices of the workload: golden-equivalent operators rebuilt on the host from the
    # committed fixture when there is no GPU; with a GPU, from our own defineBasis
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
    procs = min(cores, args.ref_procs or cores)
    nR = procs * max(1, args.ref_points_per_proc)
    ts = []
    for it in range(args.warmup + args.steps):
        if it < args.warmup and it >= 1:
            continue                # one warm-up pass is enough for a CPU path
        _v, dt = cpu_baseline_c3(calc, nR=nR, procs=procs)
        ts.append(dt)
    dt = float(np.mean(ts[-args.steps:])) if len(ts) >= args.steps else float(np.mean(ts))
    value = nR / dt
    line = {"metric": "fp64_eigensolves_per_sec", "value": value, "unit": "solves/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (Rb 60S+60S pair Hamiltonian built from quantum defects; no dataset)",
            "config": {"workload": "c3: Rb 60S1/2+60S1/2 PairStateInteractions dn=5 dl=5 dEmax=25GHz (BASELINE configs[2]), "
                                   "N=2363, k=250 eigenpairs nearest the pair state per R point",
                       "points_per_step": nR},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": procs, "kind": "port",
                             "sample": "%d R points per step, scipy eigsh shift-invert k=250, %d processes x 1 BLAS thread"
                                       % (nR, procs)},
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
    ap.add_argument("--skip-sub", action="store_true", help="skip the c1/c2 sub-workloads")
    ap.add_argument("--cpu-pair-points", type=int, default=6)
    ap.add_argument("--cpu-radial-pairs", type=int, default=150)
    ap.add_argument("--cpu-stark-fields", type=int, default=100)
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-points-per-proc", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank, world, local = dist_setup(args.gpus)
    r = bench_c3(args, rank, world)
    sub = {}
    cpu = None
    if rank == 0:
        v, dt = cpu_baseline_c3(r["calc"], nR=args.cpu_pair_points, procs=1)
        cpu = {"value": v, "unit": "solves/s", "cores": 1, "kind": "port",
               "sample": "%d of the R points, scipy.sparse.linalg.eigsh(k=250, sigma=0, tol=1e-6) per point "
                         "(the reference's call, calculations_atom_pairstate.py:3444), default BLAS threads"
                         % args.cpu_pair_points}
        if not args.skip_sub:
            # the sub-workloads need workspaces of very different sizes (141 GB for C5): drop the
            # device operators of the main workload, which would pin its cached 68 GB segment
            import gc
            import torch
            r["calc"]._ops = None
            r["calc"].last_sweep = None
            gc.collect()
            torch.cuda.empty_cache()
            try:
                sub["radial_c2"] = bench_c2(args)
            except Exception as ex:
                sub["radial_c2"] = {"error": repr(ex)}
            try:
                sub["stark_c1"] = bench_c1(args)
            except Exception as ex:
                sub["stark_c1"] = {"error": repr(ex)}
            try:
                sub["shirley_doc"] = bench_shirley(args)
            except Exception as ex:
                sub["shirley_doc"] = {"error": repr(ex)}
            try:
                sub["stark_c4"] = bench_c4(args)
            except Exception as ex:
                sub["stark_c4"] = {"error": repr(ex)}
            try:
                sub["pair_c5"] = bench_c5(args)
            except Exception as ex:
                sub["pair_c5"] = {"error": repr(ex)}
    barrier(world)
    if rank == 0:
        line = {
            "metric": "fp64_eigensolves_per_sec", "value": r["value"], "unit": "solves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (Rb 60S+60S pair Hamiltonian built from quantum defects by our own defineBasis; no dataset)",
            "config": {"workload": "c3: Rb 60S1/2+60S1/2 PairStateInteractions dn=5 dl=5 dEmax=25GHz (BASELINE configs[2]), "
                                   "N=%d, k=%d eigenpairs nearest the pair state per R point" % (r["N"], r["k"]),
                       "points_per_gpu_per_step": args.points, "batch_in_flight": r["batch"],
                       "l2": "inputs larger than L2 (batch workspace %.1f GB)" % (r["batch"] * 4 * (r["N"] ** 2) * 8 / 1e9),
                       "parallelism": "sweep points sharded, %d rank(s), no data-path collective" % world},
            "e2e": {"value": r["e2e_value"], "unit": "solves/s", "h2d_bytes_per_step": r["h2d"],
                    "d2h_bytes_per_step": r["d2h"], "ms_per_step": r["ms_e2e"] / args.steps,
                    "api": "PairStateInteractions.diagonalise(R, 250) with host buffers"},
            "gpu_launches": r["launches"] * args.steps,
            "roofline": r["roofline"], "cpu_baseline": cpu, "clocks": r["clocks"],
            "defineBasis_s": r["t_defineBasis_s"],
            "c6_from_level_diagram_GHz_um6": r["c6_fit"],
            "sub": sub,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

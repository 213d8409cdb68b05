"""world_size-2 gloo test of the sweep sharding + result gather (the N>1 host path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arc_b200 import sweep


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _collect(procs, q, timeout):
    """One result per process; a worker that died without reporting fails the test at once
    (instead of after the queue timeout)."""
    import queue
    import time
    outs, t0 = [], time.time()
    while len(outs) < len(procs):
        try:
            outs.append(q.get(timeout=1.0))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            assert not dead, "worker exited with %r before reporting" % dead
            assert time.time() - t0 < timeout, "timed out waiting for the workers"
    return outs


def _worker(rank, world, port, npoints, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sweep.dist_info() == (0, 1)              # sharding is opt-in
        assert sweep.dist_info(sweep.WORLD) == (rank, world)
        lo, hi = sweep.shard_bounds(npoints, rank, world)
        # every rank "solves" its own shard: row p holds f(p)
        local = torch.stack([torch.arange(4, dtype=torch.float64) + 10.0 * p for p in range(lo, hi)]) \
            if hi > lo else torch.zeros((0, 4), dtype=torch.float64)
        full = sweep.gather_rows(local, npoints)
        li = torch.arange(lo, hi, dtype=torch.int32).reshape(-1, 1, 1).expand(-1, 2, 3).contiguous()
        fi = sweep.gather_rows(li, npoints)
        q.put((rank, full.numpy(), fi.numpy()))
    finally:
        dist.destroy_process_group()


def _run(npoints):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, npoints, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = _collect(procs, q, 120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return outs


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 600, 1000, 4096):
        for w in (1, 2, 3, 8):
            b = [sweep.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_gather_rows_gloo_world2_ragged():
    for npoints in (7, 1):
        want = np.stack([np.arange(4.0) + 10.0 * p for p in range(npoints)])
        for rank, full, fi in _run(npoints):
            assert full.shape == (npoints, 4)
            assert np.array_equal(full, want)
            assert np.array_equal(fi[:, 0, 0], np.arange(npoints))


def _table_worker(rank, world, port, npairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arc_b200 import radial
        pairs = np.stack([np.arange(npairs), np.arange(npairs) + 1, 1 + np.arange(npairs) % 2], 1).astype(np.int32)
        seen = []

        def compute(shard):                      # stand-in for RadialBatch.integrals_device
            seen.append(len(shard))
            return torch.as_tensor(shard[:, 0] * 100.0 + shard[:, 1] + 0.5 * shard[:, 2], dtype=torch.float64)
        full = radial.sharded_table(compute, pairs)
        q.put((rank, full.numpy(), seen[0]))
    finally:
        dist.destroy_process_group()


def test_radial_table_sharded_allgather_gloo_world2():
    """Matrix-element table computed once per box: each rank evaluates its shard of the pair
    list, the shards are all-gathered (NCCL on the GPUs)."""
    npairs = 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_table_worker, args=(r, 2, port, npairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = _collect(procs, q, 120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    i = np.arange(npairs)
    want = i * 100.0 + (i + 1) + 0.5 * (1 + i % 2)
    assert sorted(o[2] for o in outs) == [500, 501]          # each rank computed only its shard
    for rank, full, _n in outs:
        assert np.array_equal(full, want)


class _FakeRadialBatch:
    """CPU stand-in for radial.RadialBatch (the real one runs the Numerov kernel): element values
    are a deterministic function of the two state records and the power, so a sharded table has
    to equal the unsharded one entry by entry."""
    kernel_launches = 2
    shard_sizes = []

    def __init__(self, states, normalise=True, device=None):
        self.states = np.asarray(states)

    def _values(self, rows):
        rows = np.asarray(rows).reshape(-1, 3)
        a, b = self.states[rows[:, 0]], self.states[rows[:, 1]]
        v = np.sin(37.0 * a["stateEnergy"] * 1e3 + a["l"] + 2.0 * a["j"]) * np.cos(91.0 * b["stateEnergy"] * 1e3 + b["l"]) \
            * (10.0 ** rows[:, 2]) + 3.0
        return torch.as_tensor(v, dtype=torch.float64)

    def raise_on_failure(self):
        pass

    def integrals(self, rows):
        return self._values(rows).numpy()

    def integrals_sharded(self, rows, group="world"):
        from arc_b200 import radial

        def compute(shard):
            _FakeRadialBatch.shard_sizes.append(len(shard))
            return self._values(shard)
        return radial.sharded_table(compute, np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 3), group)


def _definebasis_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import arc_b200
        from arc_b200 import radial
        radial.RadialBatch = _FakeRadialBatch

        def stark(group):
            atom = arc_b200.Rubidium()
            atom.radial_group = group
            calc = arc_b200.StarkMap(atom)
            calc.defineBasis(30, 0, 0.5, 0.5, 26, 34, 12)
            return calc

        def pair(group):
            atom = arc_b200.Rubidium()
            atom.radial_group = group
            calc = arc_b200.PairStateInteractions(atom, 50, 0, .5, 50, 0, .5, .5, .5, interactionsUpTo=2)
            calc.defineBasis(0., 0., 4, 4, 25e9)
            return calc
        _FakeRadialBatch.shard_sizes = []
        s_sh, p_sh = stark(sweep.WORLD), pair(sweep.WORLD)          # collective: both ranks call them
        sizes = list(_FakeRadialBatch.shard_sizes)
        s_1, p_1 = stark(None), pair(None)                          # every rank on its own, no collective
        ok = np.array_equal(s_sh.mat2, s_1.mat2) and s_sh.atom._dipole == s_1.atom._dipole
        for a, b in zip(p_sh.matR, p_1.matR):
            ok = ok and (abs(a - b)).nnz == 0 and a.nnz > 0
        ok = ok and p_sh.atom1._quadrupole == p_1.atom1._quadrupole and len(p_1.atom1._quadrupole) > 0
        q.put((rank, bool(ok), sizes, len(s_1.atom._dipole), len(p_1.atom1._dipole) + len(p_1.atom1._quadrupole)))
    finally:
        dist.destroy_process_group()


def test_definebasis_with_sharded_radial_table_gloo_world2():
    """atom.radial_group: StarkMap / PairStateInteractions.defineBasis on two ranks build the same
    matrices as an unsharded run, each rank evaluating half of every radial table (>= 1024 new
    elements) -- the host path the 2/4/8-GPU bench runs, with the Numerov kernel stubbed."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_definebasis_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = _collect(procs, q, 240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, sizes, n_stark, n_pair in outs:
        assert ok
        assert n_stark >= 1024 and n_pair >= 1024
        assert len(sizes) == 2                                       # one sharded table per defineBasis
        assert abs(sizes[0] - n_stark / 2) <= 1 and abs(sizes[1] - n_pair / 2) <= 1

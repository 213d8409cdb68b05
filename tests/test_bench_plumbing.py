"""CPU tests of bench.py's plumbing (no GPU): the JSON line carries every key of the bench
contract, both arms share one `config`, and the full-sweep block records what it should.  The
measurement functions are replaced by stubs; nothing here times anything."""
import contextlib
import io
import json
import sys
import time
import types

import numpy as np
import pytest


@pytest.fixture()
def bench(monkeypatch):
    import bench as b
    monkeypatch.setattr(b, "barrier", lambda world: None)
    monkeypatch.setattr(b, "max_over_ranks", lambda x, world: x)
    monkeypatch.setattr(b, "free_device_cache", lambda: None)
    return b


def test_json_line_carries_the_contract_keys(bench, monkeypatch):
    class Calc:
        _ops = 1
        last_sweep = 1
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2"])
    monkeypatch.setattr(bench, "dist_setup", lambda: (0, 1, 0))
    monkeypatch.setattr(bench, "bench_c3", lambda args, rank, world: {
        "value": 620.0, "ms": 950.0, "e2e_value": 610.0, "ms_e2e": 970.0, "h2d": 10, "d2h": 20, "launches": 100,
        "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []}, "N": 2363, "k": 250,
        "t_defineBasis_s": 0.12, "batch": 296, "calc": Calc(), "R": None, "work_bytes": 37.7e9})
    monkeypatch.setattr(bench, "bench_full_sweeps", lambda args, rank, world: {
        "c1": {"solves_per_s": 12000.0, "full_sweep_s": 0.085}, "_c1_calc": None})
    monkeypatch.setattr(bench, "roofline_c3", lambda N, k, npts: {
        "bound": "tensor", "achieved": 22.0, "peak": 37.0, "unit": "TFLOP/s", "frac": 0.59, "traffic": None})
    monkeypatch.setattr(bench, "cpu_baseline_c3", lambda calc, nR=4: (0.7, 1.0))
    monkeypatch.setattr(bench, "bench_c2", lambda args: {"value": 4.9e6, "gram_path": {"value": 5.1e6}})
    monkeypatch.setattr(bench, "bench_shirley", lambda args: {"value": 126.0})
    monkeypatch.setattr(bench, "cpu_legs", lambda full, args: {"c1": {"value": 45.0}})
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1                                            # ONE JSON line
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline",
                "cpu_baseline", "clocks", "full_sweeps", "sub", "summary"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] >= 3 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["ms_per_step"] == 950.0 / 2 and d["gpu_launches"] == 200
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["config"] == bench.bench_config(296, 1) and "workload" in d["config"]
    assert "model" not in d["config"]
    assert d["full_sweeps"]["c1"]["cpu_baseline"] == {"value": 45.0}
    assert not any(k.startswith("_") for k in d["full_sweeps"])
    assert d["summary"]["c1_solves_per_s"] == 12000.0 and d["summary"]["c2_gram_elements_per_s"] == 5.1e6


def test_full_sweeps_block_with_stub_calculations(bench, monkeypatch):
    import arc_b200

    def wall(fn, world):
        t0 = time.perf_counter()
        out = fn()
        return time.perf_counter() - t0, out
    monkeypatch.setattr(bench, "_wall", wall)

    class LS:
        kernel_launches = 7

    class FakeStark:
        calls = 0

        def __init__(self, atom):
            self.basisStates, self.last_sweep = [0] * 5, LS()

        def defineBasis(self, *a, **k):
            pass

        def diagonalise(self, F, group=None):
            FakeStark.calls += 1

    class FakePair(FakeStark):
        def __init__(self, atom, *a):
            super().__init__(atom)
            self._ops = 1

        def diagonalise(self, R, k, group=None):
            self.r = R

        def getC6fromLevelDiagram(self, a, b):
            return 1.5
    monkeypatch.setattr(arc_b200, "StarkMap", FakeStark)
    monkeypatch.setattr(arc_b200, "PairStateInteractions", FakePair)
    out = bench.bench_full_sweeps(types.SimpleNamespace(c5_points=16), 0, 1)
    assert {"c1", "c3", "c4_singlet", "c4_triplet", "c5"} <= set(out)
    assert all("error" not in out[k] for k in ("c1", "c3", "c4_singlet", "c4_triplet", "c5"))
    assert (out["c1"]["points"], out["c3"]["points"], out["c4_triplet"]["points"], out["c5"]["points"]) == (600, 1000, 2000, 16)
    assert len(out["c1"]["diagonalise_s_calls"]) == 3 and len(out["c4_singlet"]["diagonalise_s_calls"]) == 3
    assert len(out["c4_triplet"]["diagonalise_s_calls"]) == 1
    assert out["c3"]["c6_from_level_diagram_GHz_um6"] == 1.5 and out["c5"]["of_points"] == 4096
    assert np.isfinite(out["c1"]["solves_per_s"])


def test_reference_arm_uses_the_same_config(bench):
    a, b = bench.bench_config(296, 1), bench.bench_config(296, 1)
    assert a == b and a["points_per_step"] == 296
    src = open(bench.__file__).read()
    assert src.count("bench_config(args.points") == 2                # this arm and --impl reference

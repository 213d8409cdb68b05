#!/bin/bash
# Round 2, last GPU call (5.9 GPU-minutes left): validates the vector defineBasis host path and the C-ABI argument
# checks on the real GPU path.  Prioritised and time-boxed: smoke, the parity tests that go through defineBasis /
# the radial tables, one short bench.py run (all legs), then the rest of the suite as far as the budget reaches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))s"; tail -n 2 gpurun_out/r02_final_smoke.log
timeout -k 5 150 python -m pytest -m gpu -x -v -p no:cacheprovider --durations=8 \
    tests/test_radial_gpu.py tests/test_stark_gpu.py tests/test_shirley_gpu.py tests/test_c6_gpu.py \
    tests/test_resonances_gpu.py tests/test_sweep_semantics_gpu.py \
    "tests/test_pair_gpu.py::test_pair_small" "tests/test_pair_gpu.py::test_pair_c3" \
    "tests/test_pair_gpu.py::test_pair_sort_eigenvectors_matches_reference" \
    "tests/test_pair_gpu.py::test_level_diagram_fits_against_reference" \
    "tests/test_round2_gpu.py::test_c4_sr88_full_size_matches_reference" \
    "tests/test_round2_gpu.py::test_stark_get_state" "tests/test_round2_gpu.py::test_c6_angle_pairs_degenerate_batched" \
    "tests/test_round2_gpu.py::test_radial_wavefunction_python_route_matches_reference" \
    > gpurun_out/r02_final_pytest_a.log 2>&1; echo "pytest A rc=$? t=$(( $(date +%s) - T0 ))s"; grep -c PASSED gpurun_out/r02_final_pytest_a.log; tail -n 14 gpurun_out/r02_final_pytest_a.log
LEFT=$(( 290 - ( $(date +%s) - T0 ) ))
if [ "$LEFT" -gt 60 ]; then
# short bench run: every leg of bench.py once (C5 full sweep skipped, small CPU samples); defineBasis_s of C1/C3/C4 in full_sweeps
timeout -k 5 $LEFT python bench.py --steps 2 --warmup 3 --c5-points 0 --cpu-pair-points 1 --cpu-stark-fields 10 \
    --cpu-radial-pairs 10 > gpurun_out/r02_final_bench_short.json 2> gpurun_out/r02_final_bench_short.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"; tail -c 1500 gpurun_out/r02_final_bench_short.json; tail -n 3 gpurun_out/r02_final_bench_short.err
fi
LEFT=$(( 300 - ( $(date +%s) - T0 ) ))
if [ "$LEFT" -gt 20 ]; then
timeout -k 5 $LEFT python -m pytest -m gpu -x -v -p no:cacheprovider --durations=8 \
    tests/test_twostage_gpu.py tests/test_stein_gpu.py tests/test_round2_gpu.py tests/test_pair_gpu.py \
    > gpurun_out/r02_final_pytest_b.log 2>&1; echo "pytest B rc=$? (124 = budget ran out) t=$(( $(date +%s) - T0 ))s"; grep -c PASSED gpurun_out/r02_final_pytest_b.log; tail -n 6 gpurun_out/r02_final_pytest_b.log
fi

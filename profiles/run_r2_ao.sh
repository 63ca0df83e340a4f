#!/bin/bash
# compute-sanitizer initcheck (reads of uninitialised device memory) and racecheck (shared-memory hazards), bounded
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool initcheck --error-exitcode 7 --log-file gpurun_out/ao_initcheck.log python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not cold_start" > gpurun_out/ao_pytest_init.log 2>&1; echo "initcheck rc=$?"
tail -2 gpurun_out/ao_pytest_init.log; grep -c "Uninitialized" gpurun_out/ao_initcheck.log; grep -A12 "Uninitialized" gpurun_out/ao_initcheck.log | grep -E "Uninitialized|at .*\(|by thread" | head -20; tail -2 gpurun_out/ao_initcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/ao_racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_pointops.py -q -m gpu -x -k "rulebook or indice_pairs or mask_sorted or three_nn or voxel_query or voxeliz" > gpurun_out/ao_pytest_race.log 2>&1; echo "racecheck rc=$?"
tail -2 gpurun_out/ao_pytest_race.log; tail -3 gpurun_out/ao_racecheck.log

#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in the round (voxelizer rank kernel + level-0 table, BatchNorm)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/ah_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_bench_parity.py -x -q -m gpu -k "level0_coordinate_table or ragged_batch or training_batchnorm_matches_torch" > gpurun_out/ah_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/ah_pytest.log; tail -5 gpurun_out/ah_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/ah_racecheck.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "level0_coordinate_table" > gpurun_out/ah_pytest2.log 2>&1; echo "racecheck rc=$?"
tail -2 gpurun_out/ah_pytest2.log; tail -5 gpurun_out/ah_racecheck.log

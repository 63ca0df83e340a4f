#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cold_start or run_stream" > gpurun_out/an_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/an_pytest.log
# the same test against the old allocation (zeros) must fail: proves the test sees the race
sed -i 's/dev_off=torch.empty((batch + 1,), dtype=torch.int32, device=device))/dev_off=torch.zeros((batch + 1,), dtype=torch.int32, device=device))/' from-voxel-to-point_b200/pipeline.py
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cold_start" > gpurun_out/an_pytest_old.log 2>&1; echo "with the old allocation rc=$? (expected 1)"; tail -3 gpurun_out/an_pytest_old.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py::test_capacity_overflow_grows_the_arena_and_reruns tests/test_pointops.py -m gpu -q -rA > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|PASSED|FAILED|Error|error" gpurun_out/c_pytest.log | tail -30

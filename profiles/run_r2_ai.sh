#!/bin/bash
# compute-sanitizer memcheck over the GPU parity suite (bounded: whatever finishes in 9 minutes)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 540 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/ai_memcheck.log python -m pytest tests/test_bench_parity.py tests/test_sharding.py -q -m gpu > gpurun_out/ai_pytest.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/ai_pytest.log; grep -c "Invalid\|Error" gpurun_out/ai_memcheck.log; tail -4 gpurun_out/ai_memcheck.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/ak_mem.log python -m pytest tests/test_bench_parity.py -q -m gpu -k "run_stream_matches_reference" > gpurun_out/ak_pytest_$i.log 2>&1; echo "run $i rc=$?"; grep -E "passed|failed" gpurun_out/ak_pytest_$i.log | tail -1
done
grep -B30 "AssertionError" gpurun_out/ak_pytest_*.log | grep -E "assert|Error|>" | head -20

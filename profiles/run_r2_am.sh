#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/am_mem.log python profiles/debug_stream2.py kitti_b8 waymo_b4 > gpurun_out/am_1.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/am_1.log
timeout 540 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/am_memcheck.log python -m pytest tests/test_bench_parity.py tests/test_sharding.py -q -m gpu > gpurun_out/am_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/am_pytest.log; tail -3 gpurun_out/am_memcheck.log

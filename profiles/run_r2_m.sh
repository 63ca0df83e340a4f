#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_parity.py -x -q -m gpu -k "batchnorm or training_step" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/m_pytest.log

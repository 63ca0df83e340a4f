#!/bin/bash
# Round-2: full GPU suite + default bench line + micro + reference arm (outputs are small text files only).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/v_pytest.log | tail -12
timeout 900 python bench.py > gpurun_out/v_bench_default.json 2> gpurun_out/v_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --precision bf16 --no-extras --no-cpu-baseline > gpurun_out/v_bench_bf16.json 2> gpurun_out/v_bench_bf16.err; echo "bench bf16 rc=$?"
timeout 600 python bench.py --workload micro > gpurun_out/v_micro.json 2> gpurun_out/v_micro.err; echo "micro rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v_ref.json 2> gpurun_out/v_ref.err; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/v_smoke.log

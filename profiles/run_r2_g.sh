#!/bin/bash
# Round-2: two-GPU sanity of the sharded 64-frame workload + the new module-path tests.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_parity.py -m gpu -q -k "module_api or unrecognised" > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g_bench_2gpu.json 2> gpurun_out/g_bench_2gpu.err; echo "2gpu rc=$?"
timeout 600 python bench.py --workload waymo_64 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/g_bench_w64_1gpu.json 2> gpurun_out/g_bench_w64_1gpu.err; echo "w64 1gpu rc=$?"

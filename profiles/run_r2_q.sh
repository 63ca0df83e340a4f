#!/bin/bash
# multi-launch steps with two lanes: waymo_64 on one GPU, then the same 64 frames over two GPUs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python bench.py --workload waymo_64 --no-extras --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/q_w64_1gpu.json 2> gpurun_out/q_w64_1gpu.err; echo "1gpu rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_2gpu.json 2> gpurun_out/q_2gpu.err; echo "2gpu rc=$?"
python - <<P
import json
for f in ("q_w64_1gpu","q_2gpu"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"], d["config"]["workload"], d["scaling"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 gpurun_out/q_2gpu.err

#!/bin/bash
# the driver's scaling run at the largest N: 64 Waymo frames sharded over the GPUs of one box
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus visible: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/af_${N}gpu.json 2> gpurun_out/af_${N}gpu.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/af_${N}gpu_ref.json 2> gpurun_out/af_${N}gpu_ref.err; echo "ref rc=$?"
python - <<P
import json
for f in ("af_${N}gpu","af_${N}gpu_ref"):
    try:
        lines=[l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")]
        print(f, "json lines:", len(lines))
        d=json.loads(lines[-1])
        print(d.get("impl"), d["value"], d.get("ms_per_step"), d["e2e"]["value"], d["n_gpus"], d["config"]["workload"], d.get("scaling"), d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", e)
P
tail -3 gpurun_out/af_${N}gpu.err

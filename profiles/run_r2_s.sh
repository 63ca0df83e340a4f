#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for depth in 2 3 4 6; do
  for prec in fp32 bf16; do
    FV2P_STREAM_DEPTH=$depth timeout 300 python bench.py --workload waymo_b4 --precision $prec --no-extras --no-cpu-baseline --steps 40 > gpurun_out/s_${prec}_d$depth.json 2> gpurun_out/s_${prec}_d$depth.err
    echo "depth $depth $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/s_${prec}_d$depth.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e: print("ERR", e)
P
  done
done

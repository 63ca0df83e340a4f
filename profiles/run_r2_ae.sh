#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for pri in 0 1; do
  for prec in fp32 bf16; do
    for wlx in waymo_b4 kitti_b8; do
    FV2P_FIRST_GEO_PRIORITY=$pri timeout 300 python bench.py --workload $wlx --precision $prec --no-extras --no-cpu-baseline > gpurun_out/ae_${wlx}_${prec}_$pri.json 2> gpurun_out/ae_${wlx}_${prec}_$pri.err
    echo "first-geo priority $pri $wlx $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/ae_${wlx}_${prec}_$pri.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e: print("ERR", e)
P
    done
  done
done

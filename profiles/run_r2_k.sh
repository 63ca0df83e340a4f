#!/bin/bash
# which rulebooks are worth grouping: all / submanifold only / none
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for g in all subm none; do
  for prec in fp32 bf16; do
    for wlx in waymo_b4 kitti_b8; do
      timeout 300 python bench.py --workload $wlx --precision $prec --group-rows $g --no-extras --no-cpu-baseline > gpurun_out/k_${wlx}_${prec}_$g.json 2> gpurun_out/k_${wlx}_${prec}_$g.err
      echo "$wlx $prec group=$g rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/k_${wlx}_${prec}_$g.json").read().strip().splitlines()[-1])
    st=d.get("stages",{})
    print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), "geo", st.get("rulebooks_ms"), st.get("rulebook_launches"), "conv", st.get("conv_ms_sum"), [ (l["c"], l["ms"]) for l in st.get("layers",[]) if l["l"] in (5,10,15,20)])
except Exception as e: print("ERR", e)
P
    done
  done
done
du -sh gpurun_out

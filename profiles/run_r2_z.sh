#!/bin/bash
# A-tile producer: cp.async vs TMA gather4, after the fp32 split change
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for g in lsu tma; do
  for prec in fp32 bf16; do
    timeout 300 python bench.py --workload waymo_b4 --precision $prec --gather $g --no-extras --no-cpu-baseline > gpurun_out/z_${prec}_$g.json 2> gpurun_out/z_${prec}_$g.err
    echo "gather $g $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/z_${prec}_$g.json").read().strip().splitlines()[-1])
    st=d.get("stages",{})
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv", st.get("conv_ms_sum"), [(l["c"], l["ms"]) for l in st.get("layers",[]) if l["l"] in (6,7,10,11,12,15,16,17,20)])
except Exception as e: print("ERR", e)
P
  done
done

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/aa_pytest.log
for prec in fp32 bf16; do
  for wlx in waymo_b4 kitti_b8; do
    timeout 300 python bench.py --workload $wlx --precision $prec --no-extras --no-cpu-baseline > gpurun_out/aa_${wlx}_${prec}.json 2> gpurun_out/aa_${wlx}_${prec}.err
    echo "$wlx $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/aa_${wlx}_${prec}.json").read().strip().splitlines()[-1])
    st=d.get("stages",{})
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv", st.get("conv_ms_sum"))
except Exception as e: print("ERR", e)
P
  done
done

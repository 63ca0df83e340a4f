#!/bin/bash
# voxelizer: one rank kernel + level-0 table built in it.  Full GPU suite, then the bench.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/n_pytest.log
for prec in fp32 bf16; do
  for wlx in waymo_b4 kitti_b8; do
    timeout 300 python bench.py --workload $wlx --precision $prec --no-extras --no-cpu-baseline > gpurun_out/n_${wlx}_${prec}.json 2> gpurun_out/n_${wlx}_${prec}.err
    echo "$wlx $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/n_${wlx}_${prec}.json").read().strip().splitlines()[-1])
    st=d.get("stages",{})
    print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), "launches", d["gpu_launches"]/d["steps"], "vox", st.get("voxelize_ms"), "geo", st.get("rulebooks_ms"), st.get("rulebook_launches"), "conv", st.get("conv_ms_sum"))
except Exception as e: print("ERR", e)
P
  done
done
du -sh gpurun_out

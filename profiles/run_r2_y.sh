#!/bin/bash
# role timers of the fp32 layers after the split-bf16 change (timers build made on the box; the product build is untouched in the repo)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
FV2P_EXTRA_NVCC_FLAGS=-DFV2P_TC_TIMERS timeout 300 python from-voxel-to-point_b200/build.py --force > gpurun_out/y_build.log 2>&1; echo "build rc=$?"
for layer in 2 7 12 17; do
  timeout 200 python profiles/run_layer.py --workload waymo_b4 --precision fp32 --layer $layer --debug 0 3 4 6 > gpurun_out/y_timers_fp32_l$layer.log 2>&1
  echo "== layer $layer"; grep -A15 "role timers" gpurun_out/y_timers_fp32_l$layer.log | head -16; grep -i "debug\|ms" gpurun_out/y_timers_fp32_l$layer.log | tail -6
done

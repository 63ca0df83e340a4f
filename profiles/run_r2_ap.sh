#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for wlx in waymo_b4 kitti_b8; do for prec in fp32 bf16; do
  timeout 400 python profiles/contention.py --workload $wlx --precision $prec > gpurun_out/ap_${wlx}_$prec.log 2>&1; echo "== $wlx $prec rc=$?"; grep "concurrent=True\|alone" gpurun_out/ap_${wlx}_$prec.log
done; done

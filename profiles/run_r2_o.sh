#!/bin/bash
# per-kernel durations of the voxelizer (ncu launch list, third eager step) on both workloads
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for wlx in waymo_b4 kitti_b8; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vox_|table_insert|fill_ranges|subm_probe" -s 20 -c 10 --csv --log-file gpurun_out/o_vox_$wlx.csv python profiles/run_geo.py --workload $wlx > gpurun_out/o_ncu_$wlx.log 2>&1; echo "ncu $wlx rc=$?"
  python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/o_vox_$wlx.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:40], r[-1])
P
done

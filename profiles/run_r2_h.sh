#!/bin/bash
# Round-2: smoke(), ncu --set full of the voxelizer + level-0 table + prefill kernels (HBM evidence for the hash stage).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/ncu
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/h_smoke.log
timeout 600 ncu --set full --clock-control none -k regex:"vox_|table_insert|fill_ranges|scan_chunk" -s 22 -c 11 -o /tmp/ncu/vox python profiles/run_geo.py --workload waymo_b4 > gpurun_out/h_ncu.log 2>&1; echo "ncu vox rc=$?"
python profiles/extract_ncu.py /tmp/ncu/vox.ncu-rep "ncu --set full: voxelizer, prefill and level-0 table kernels of one waymo_b4 step (726k points -> 325k voxels; eager, cold L2)" > gpurun_out/r2_ncu_vox_waymo.md 2>> gpurun_out/h_ncu.log
cat gpurun_out/r2_ncu_vox_waymo.md | cut -c1-150

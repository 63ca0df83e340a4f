#!/bin/bash
# Final build of round 2: ncu launch list of two eager steps + --set full of the geometry / voxelizer kernels and of four
# conv layers per precision, condensed to text on the box (the .ncu-rep files are too large to bring back).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/ncu
rm -f gpurun_out/r2_ncu_tc_waymo.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_waymo_fp32.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/u_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"subm_probe|conv_insert|conv_rank|conv_nbr|group_|vox_|table_insert|fill_ranges|conv_small" -s 78 -c 39 -o /tmp/ncu/geo python profiles/run_geo.py --workload waymo_b4 > gpurun_out/u_ncu2.log 2>&1; echo "ncu geo rc=$?"
python profiles/extract_ncu.py /tmp/ncu/geo.ncu-rep "ncu --set full: voxelizer, prefill and every geometry kernel of one waymo_b4 step (fp32 build; eager, cold L2)" > gpurun_out/r2_ncu_geo_waymo_fp32.md 2>> gpurun_out/u_ncu2.log
for prec in fp32 bf16; do
  for layer in 2 7 12 17; do
    timeout 300 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 41 -c 1 -o /tmp/ncu/tc_${prec}_l$layer python profiles/run_layer.py --workload waymo_b4 --precision $prec --layer $layer > gpurun_out/u_ncu_${prec}_l$layer.log 2>&1
    python profiles/extract_ncu.py /tmp/ncu/tc_${prec}_l$layer.ncu-rep "conv layer $layer, $prec, waymo_b4" >> gpurun_out/r2_ncu_tc_waymo.md 2>> gpurun_out/u_ncu_${prec}_l$layer.log
  done
done
du -sh gpurun_out; wc -l gpurun_out/r2_ncu_geo_waymo_fp32.md gpurun_out/r2_ncu_tc_waymo.md gpurun_out/r2_launches_waymo_fp32.csv

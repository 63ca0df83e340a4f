#!/bin/bash
# Round-2 call: tests, bench (4 configs), per-kernel launch list, role timers.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rel err" gpurun_out/b_pytest.log | tail -12
for wl in waymo_b4 kitti_b8; do
  for prec in fp32 bf16; do
    timeout 300 python bench.py --workload $wl --precision $prec --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_${wl}_$prec.json 2> gpurun_out/b_${wl}_$prec.err
    echo "$wl $prec rc=$?"
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b_launches_waymo_fp32.csv python bench.py --workload waymo_b4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$?"
FV2P_EXTRA_NVCC_FLAGS=-DFV2P_TC_TIMERS timeout 300 python from-voxel-to-point_b200/build.py --force > gpurun_out/b_build.log 2>&1
for prec in fp32 bf16; do
  for layer in 2 7 12 17; do
    timeout 200 python profiles/run_layer.py --workload waymo_b4 --precision $prec --layer $layer --debug 0 3 4 6 > gpurun_out/b_timers_${prec}_l$layer.log 2>&1
  done
done
echo done

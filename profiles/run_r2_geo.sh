#!/bin/bash
# Round-2 call 2: validate the second-generation geometry pass (tests), then bench both workloads.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/g_pytest.log
for wl in waymo_b4 kitti_b8; do
  for prec in fp32 bf16; do
    timeout 300 python bench.py --workload $wl --precision $prec --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_${wl}_$prec.json 2> gpurun_out/g_${wl}_$prec.err
    echo "$wl $prec rc=$?"; tail -c 600 gpurun_out/g_${wl}_$prec.json | head -c 300; echo
  done
done

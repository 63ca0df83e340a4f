#!/bin/bash
# Round-2 exploratory call: baseline sanity, reference CUDA arm, role timers + ablations on Waymo-sized layers.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python - > gpurun_out/x_refgpu.log 2>&1 <<'P'
import json, torch, bench
dev = torch.device("cuda", 0)
for w in ("kitti_b8", "waymo_b4"):
    print(w, json.dumps(bench.reference_gpu(bench.WORKLOADS[w], dev)))
P
echo "refgpu rc=$?"
for prec in fp32 bf16; do
  timeout 300 python bench.py --workload waymo_b4 --precision $prec --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/x_waymo_$prec.json 2> gpurun_out/x_waymo_$prec.err
done
FV2P_EXTRA_NVCC_FLAGS=-DFV2P_TC_TIMERS timeout 300 python from-voxel-to-point_b200/build.py --force > gpurun_out/x_build.log 2>&1
for prec in fp32 bf16; do
  for layer in 2 7 12 17; do
    timeout 200 python profiles/run_layer.py --workload waymo_b4 --precision $prec --layer $layer --debug 0 3 4 6 > gpurun_out/x_timers_${prec}_l$layer.log 2>&1
  done
done
echo done

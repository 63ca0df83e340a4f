#!/bin/bash
# indirect rows A/B: parity tests first, then the bench with and without the grouped copy
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bench_parity.py -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/j_pytest.log
for ind in 1 0; do
  for prec in fp32 bf16; do
    for wlx in waymo_b4 kitti_b8; do
      timeout 300 python bench.py --workload $wlx --precision $prec --indirect-rows $ind --no-extras --no-cpu-baseline > gpurun_out/j_${wlx}_${prec}_ind$ind.json 2> gpurun_out/j_${wlx}_${prec}_ind$ind.err
      echo "$wlx $prec indirect=$ind rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/j_${wlx}_${prec}_ind$ind.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("gpu_launches"), {k:v for k,v in d.get("stages",{}).items()} if isinstance(d.get("stages"),dict) else None)
except Exception as e: print("ERR", e)
P
    done
  done
done
du -sh gpurun_out

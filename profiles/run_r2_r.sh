#!/bin/bash
# run_stream with staging one batch ahead: parity, then e2e numbers
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_parity.py -x -q -m gpu > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r_pytest.log
for prec in fp32 bf16; do
  for wlx in waymo_b4 kitti_b8; do
    timeout 300 python bench.py --workload $wlx --precision $prec --no-extras --no-cpu-baseline > gpurun_out/r_${wlx}_${prec}.json 2> gpurun_out/r_${wlx}_${prec}.err
    echo "$wlx $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r_${wlx}_${prec}.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("serial_call_ms"))
except Exception as e: print("ERR", e)
P
  done
done

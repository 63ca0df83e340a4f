#!/bin/bash
# epilogue look-ahead: parity, then the bench on both workloads / precisions
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_parity.py -x -q -m gpu > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
for prec in fp32 bf16; do
  for wlx in waymo_b4 kitti_b8; do
    timeout 300 python bench.py --workload $wlx --precision $prec --no-extras --no-cpu-baseline > gpurun_out/l_${wlx}_${prec}.json 2> gpurun_out/l_${wlx}_${prec}.err
    echo "$wlx $prec rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/l_${wlx}_${prec}.json").read().strip().splitlines()[-1])
    st=d.get("stages",{})
    print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), "geo", st.get("rulebooks_ms"), "conv", st.get("conv_ms_sum"), [ (l["c"], l["ms"]) for l in st.get("layers",[]) if l["l"] in (1,2,6,7,11,12,16,17)])
except Exception as e: print("ERR", e)
P
  done
done
du -sh gpurun_out

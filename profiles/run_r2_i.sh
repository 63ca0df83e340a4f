#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for prec in fp32 bf16; do
  timeout 600 python profiles/contention.py --workload waymo_b4 --precision $prec > gpurun_out/i_contention_$prec.log 2>&1; echo "rc=$?"; cat gpurun_out/i_contention_$prec.log | tail -12
  timeout 300 python profiles/timeline.py --workload waymo_b4 --precision $prec > gpurun_out/i_timeline_$prec.log 2>&1; tail -26 gpurun_out/i_timeline_$prec.log
done

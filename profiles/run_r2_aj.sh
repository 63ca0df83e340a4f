#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfgx in "2 2 1" "1 2 1" "2 2 0"; do
  set -- $cfgx
  timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/aj_mem.log python profiles/debug_stream.py waymo_b4 $1 $2 $3 > gpurun_out/aj_$1_$2_$3.log 2>&1; echo "lanes $1 depth $2 graph $3 rc=$?"; tail -10 gpurun_out/aj_$1_$2_$3.log
done

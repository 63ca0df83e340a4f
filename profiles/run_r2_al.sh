#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/al_mem.log python profiles/debug_stream2.py kitti_b8 waymo_b4 > gpurun_out/al_1.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/al_1.log
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/al_mem.log python profiles/debug_stream2.py waymo_b4 > gpurun_out/al_2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/al_2.log
timeout 300 python profiles/debug_stream2.py kitti_b8 waymo_b4 > gpurun_out/al_3.log 2>&1; echo "no sanitizer rc=$?"; tail -8 gpurun_out/al_3.log

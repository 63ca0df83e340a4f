"""Profiling helper: two eager warm-up steps of the bench workload, then ONE more eager step whose geometry kernels
(`ncu -k regex:... -s <launches of the warm-up> -c <n>`) are captured.  Usage (under gpurun):
    ncu --set full --clock-control none -k regex:"subm_probe|conv_insert|conv_rank|conv_nbr|group_|vox_|table_insert|fill_ranges" \
        -s 94 -c 47 -o gpurun_out/geo python profiles/run_geo.py --workload waymo_b4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--workload", default="waymo_b4")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
net, hp, state, cfg = bench.build_model(wl, dev, a.precision)
frames = bench.make_frames(wl, 0, wl["batch"])
pts, off, mfp, _ = hp.upload(frames, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(a.steps):
    flush.zero_()
    h = hp.launch_resident(pts, off, mfp)
    outs, info = hp.finish(h)
print("rows per level", info["counts"], "launches per step", hp.engine.launch_count(table0_built=True) + hp.voxelizer.LAUNCHES)

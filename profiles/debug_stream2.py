"""Debug helper: two workloads streamed one after the other in one process (like tests/test_bench_parity.py), row counts per batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = "cuda"
for wlname in sys.argv[1:] or ["kitti_b8", "waymo_b4"]:
    wl = bench.WORKLOADS[wlname]
    net, hp, state, cfg = bench.build_model(wl, torch.device(dev), "fp32", use_graph=True)
    frames = bench.make_frames(wl, 0, wl["batch"])
    for i, res in enumerate(hp.run_stream((frames for _ in range(4)), dev)):
        print(wlname, "batch", i, "counts", res["counts"], "rows", tuple(res["encoded_indices"].shape), flush=True)

"""Debug helper: HotPath.run_stream on the same waymo_b4 batch N times; reports, per yielded batch, whether the stride-8
result equals the first batch's plain-call result (run it under compute-sanitizer to shake out timing-dependent races)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

wlname = sys.argv[1] if len(sys.argv) > 1 else "waymo_b4"
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
use_graph = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
wl = bench.WORKLOADS[wlname]
dev = torch.device("cuda", 0)
net, hp, state, cfg = bench.build_model(wl, dev, "fp32", use_graph=use_graph)
hp.lanes = lanes
frames = bench.make_frames(wl, 0, wl["batch"])
# the reference result comes from a second, independent HotPath: the streamed one starts cold, like a fresh caller
net2, hp2, _, _ = bench.build_model(wl, dev, "fp32", use_graph=False)
bd, info = hp2(frames, dev, fetch="encoded")
ref_ind = info["encoded_indices_host"].clone().numpy()
ref_feat = info["encoded_features_host"].clone().numpy()
ref_counts = list(info["counts"])
print("plain call counts", ref_counts)
bad = 0
for i, res in enumerate(hp.run_stream((frames for _ in range(8)), dev, depth)):
    ind = res["encoded_indices"].numpy()
    feat = res["encoded_features"].numpy()
    same_counts = res["counts"] == ref_counts
    same_ind = ind.shape == ref_ind.shape and np.array_equal(ind, ref_ind)
    nd = int((ind != ref_ind).any(axis=1).sum()) if ind.shape == ref_ind.shape else -1
    same_feat = feat.shape == ref_feat.shape and np.array_equal(feat, ref_feat)
    fd = int((feat != ref_feat).any(axis=1).sum()) if feat.shape == ref_feat.shape else -1
    print("batch", i, "counts", "ok" if same_counts else res["counts"], "indices", "ok" if same_ind else "DIFF rows=%d" % nd,
          "features", "ok" if same_feat else "DIFF rows=%d" % fd)
    bad += (not same_counts) or (not same_ind) or (not same_feat)
print("lanes", lanes, "depth", depth, "graph", use_graph, "bad batches:", bad)

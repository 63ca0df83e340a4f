"""How much of a graph-replayed step is the conv pass, and what the concurrent geometry costs it.
Prints the step time (CUDA events around one graph replay, L2 flushed first, median of 9) for the persistent geometry
grids at 1 / 2 / 4 / 8 CTAs per SM, with the geometry and the feature pass on forked streams (the product) and on ONE
stream (no overlap), and the feature pass alone (graph of the 21 conv launches over an already built geometry).
    python profiles/contention.py --workload waymo_b4 --precision fp32
"""
import argparse
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--workload", default="waymo_b4")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
from fv2p_b200 import _lib  # noqa: E402
lib = _lib.load()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
frames = bench.make_frames(wl, 0, wl["batch"])


def step_ms(hp, n=9):
    ts = []
    for _ in range(3):
        hp.finish(hp.launch_graph(0))
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h = hp.launch_graph(0)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    hp.finish(h)
    return statistics.median(ts)


for concurrent in (True, False):
    for ctas in (0, 1, 2, 8):
        lib.fv2p_debug_geo_ctas(ctypes.c_int(ctas))
        net, hp, state, cfg = bench.build_model(wl, dev, a.precision, use_graph=True)
        hp.engine.concurrent = concurrent
        hp.upload(frames, dev)
        print("concurrent=%s geometry CTAs/SM=%s: step %.3f ms" % (concurrent, ctas or "default(4)", step_ms(hp)))
        del hp, net
lib.fv2p_debug_geo_ctas(ctypes.c_int(0))
# the feature pass alone, as a graph
net, hp, state, cfg = bench.build_model(wl, dev, a.precision, use_graph=False)
pts, off, mfp, _ = hp.upload(frames, dev)
h = hp.launch_resident(pts, off, mfp)
hp.finish(h)
eng, arena = hp.engine, h["arena"]
prm = eng._prepare_params(dev)
vf, cap0 = h["vox"]["voxel_features"], h["vox"]["cap"]
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    for i, p in enumerate(prm):
        eng.run_conv_step(arena, i, p, vf, cap0)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i, p in enumerate(prm):
        eng.run_conv_step(arena, i, p, vf, cap0)
ts = []
for _ in range(9):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print("feature pass alone (21 conv launches, graph): %.3f ms" % statistics.median(ts))

"""Where the conv kernels sit inside one graph-replayed step: CTA 0 of every tensor-core conv launch stamps
%globaltimer at entry and exit (debug mode 7).  Prints start/end relative to the first conv of the step, the gap to
the previous conv (time the feature pass waited for geometry or for the launch), and the step's total.
    python profiles/timeline.py --precision fp32
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--workload", default="kitti_b8")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
net, hp, state, cfg = bench.build_model(wl, dev, a.precision)
frames = bench.make_frames(wl, 0, wl["batch"])
from fv2p_b200 import _lib  # noqa: E402
lib = _lib.load()
slot = hp.upload(frames, dev)
for _ in range(5):
    h = hp.launch_graph(0)
    hp.finish(h)
lib.fv2p_debug_set(ctypes.c_int(7))
buf = (ctypes.c_ulonglong * 512)()
lib.fv2p_debug_stamps(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
h = hp.launch_graph(0)
e1.record()
hp.finish(h)
n = lib.fv2p_debug_stamps(buf)
lib.fv2p_debug_set(ctypes.c_int(0))
st = np.array(list(buf), dtype=np.uint64).reshape(256, 2)[:n].astype(np.int64)
order = np.argsort(st[:, 0])
st = st[order]
t0 = st[0, 0]
tc_steps = [s for s, p in zip(hp.engine.steps, hp.engine._prepare_params(dev)) if p["mode"] in (1, 2)]
print("step (events) %.3f ms, %d stamped conv launches" % (e0.elapsed_time(e1), n))
prev_end = t0
for i, (b, e) in enumerate(st):
    s = tc_steps[i] if i < len(tc_steps) else None
    print("%2d %-8s %3d>%-3d start %8.1f us  dur %7.1f us  gap %6.1f us" %
          (i, s.key if s else "?", s.cin if s else 0, s.cout if s else 0, (b - t0) / 1e3, (e - b) / 1e3, (b - prev_end) / 1e3))
    prev_end = e
print("first conv start -> last conv end: %.1f us" % ((st[-1, 1] - t0) / 1e3))

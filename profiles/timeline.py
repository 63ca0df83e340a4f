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
ap.add_argument("--stream", type=int, default=0, help="N > 0: stamp the last conv launches of an N-batch run_stream")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
net, hp, state, cfg = bench.build_model(wl, dev, a.precision, use_graph=True)
frames = bench.make_frames(wl, 0, wl["batch"])
from fv2p_b200 import _lib  # noqa: E402
lib = _lib.load()
slot = hp.upload(frames, dev)
for _ in range(5):
    h = hp.launch_graph(0)
    hp.finish(h)
if a.stream > 0:
    # two batches in flight (engine lanes): which lane a launch belongs to shows in its output pointer
    for _ in hp.run_stream((frames for _ in range(400)), dev):  # long enough for clocks and host to settle
        pass
    lib.fv2p_debug_set(ctypes.c_int(7))
    buf = (ctypes.c_ulonglong * 1024)()
    lib.fv2p_debug_stamps(buf)
    import time
    t0 = time.perf_counter()
    for _ in hp.run_stream((frames for _ in range(a.stream)), dev):
        pass
    wall = time.perf_counter() - t0
    n = lib.fv2p_debug_stamps(buf)
    lib.fv2p_debug_set(ctypes.c_int(0))
    st = np.array(list(buf), dtype=np.uint64).reshape(256, 4)[:min(n, 256)].astype(np.int64)
    st = st[np.argsort(st[:, 0])]
    lanes = {p: i for i, p in enumerate(sorted(set((st[:, 3] >> 28).tolist())))}
    print("run_stream: %d batches in %.2f ms (%.3f ms/batch); %d conv launches stamped, last %d shown" %
          (a.stream, wall * 1e3, wall * 1e3 / a.stream, n, len(st)))
    t0 = st[0, 0]
    prev_end = t0
    busy = 0
    for b, e, shape, ptr in st:
        print("lane~%d %3d>%-3d start %9.1f us  dur %7.1f us  gap %7.1f us" %
              (lanes[ptr >> 28], shape // 1000, shape % 1000, (b - t0) / 1e3, (e - b) / 1e3, (b - prev_end) / 1e3))
        busy += e - b
        prev_end = max(prev_end, e)
    print("span %.1f us, sum of conv durations %.1f us" % ((prev_end - t0) / 1e3, busy / 1e3))
    sys.exit(0)
lib.fv2p_debug_set(ctypes.c_int(7))
buf = (ctypes.c_ulonglong * 1024)()
lib.fv2p_debug_stamps(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
h = hp.launch_graph(0)
e1.record()
hp.finish(h)
n = lib.fv2p_debug_stamps(buf)
lib.fv2p_debug_set(ctypes.c_int(0))
st = np.array(list(buf), dtype=np.uint64).reshape(256, 4)[:n].astype(np.int64)
order = np.argsort(st[:, 0])
st = st[order]
t0 = st[0, 0]
tc_steps = [s for s, p in zip(hp.engine.steps, hp.engine._prepare_params(dev)) if p["mode"] in (1, 2)]
print("step (events) %.3f ms, %d stamped conv launches" % (e0.elapsed_time(e1), n))
prev_end = t0
for i, (b, e, _, _) in enumerate(st):
    s = tc_steps[i] if i < len(tc_steps) else None
    print("%2d %-8s %3d>%-3d start %8.1f us  dur %7.1f us  gap %6.1f us" %
          (i, s.key if s else "?", s.cin if s else 0, s.cout if s else 0, (b - t0) / 1e3, (e - b) / 1e3, (b - prev_end) / 1e3))
    prev_end = e
print("first conv start -> last conv end: %.1f us" % ((st[-1, 1] - t0) / 1e3))

"""Profiling helper: builds the bench workload, runs two full steps, then re-launches one conv layer a few
times so that `ncu -k regex:<kernel> -s <skip> -c <n>` can capture it.  Usage (under gpurun):
    ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 2 \
        -o gpurun_out/prof python profiles/run_layer.py --precision bf16 --layer 12
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--workload", default="kitti_b8")
ap.add_argument("--layer", type=int, default=12)
ap.add_argument("--repeat", type=int, default=3)
ap.add_argument("--tma", type=int, default=-1)
ap.add_argument("--debug", type=int, nargs="*", default=[])
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
net, hp, state, cfg = bench.build_model(wl, dev, a.precision)
frames = bench.make_frames(wl, 0, wl["batch"])
pts, off, mfp, _ = hp.upload(frames, dev)
from fv2p_b200 import _lib as _L  # noqa: E402
_L.load().fv2p_tc_gather_mode(a.tma)
for _ in range(2):
    h = hp.launch_resident(pts, off, mfp)
    hp.finish(h)
from fv2p_b200 import _lib  # noqa: E402
eng, arena = hp.engine, h["arena"]
prm = eng._prepare_params(dev)
st, p = eng.steps[a.layer], prm[a.layer]
vf, cap0 = h["vox"]["voxel_features"], h["vox"]["cap"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(a.repeat):
    flush.zero_()
    eng.run_conv_step(arena, a.layer, p, vf, cap0)
torch.cuda.synchronize()
import ctypes  # noqa: E402
for mode in a.debug:
    lib = _lib.load()
    lib.fv2p_debug_set(ctypes.c_int(mode))
    ts = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_conv_step(arena, a.layer, p, vf, cap0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("debug mode", mode, "ms", min(ts))
    lib.fv2p_debug_set(ctypes.c_int(0))
if hasattr(_lib.load(), "fv2p_debug_timers"):
    import numpy as np  # noqa: E402
    buf = (ctypes.c_ulonglong * (160 * 16))()
    eng.run_conv_step(arena, a.layer, p, vf, cap0)
    _lib.load().fv2p_debug_timers(buf)
    t = np.array(list(buf), dtype=np.float64).reshape(160, 16)[:148]
    names = ["total", "stages", "tiles", "prod wait(empty)", "prod issue", "prod tile prologue", "mma wait(full)",
             "mma wait(tmem)", "mma issue", "epi wait", "epi work", "xform wait", "xform work"]
    busiest = int(np.argmax(t[:, 0]))
    print("role timers (cycles): mean over CTAs | busiest CTA %d" % busiest)
    for i, nm in enumerate(names):
        print("  %-20s %12.0f %12.0f" % (nm, t[:, i].mean(), t[busiest, i]))
    print("  cycles per stage (busiest CTA): %.0f; producer warp 0 handled 1/%d of the stages" %
          (t[busiest, 0] / max(t[busiest, 1], 1), 8))
print("layer", a.layer, st.key, st.cin, st.cout, "mode", p["mode"], "rows", hp.finish(h)[1]["counts"])

#!/usr/bin/env python
"""Benchmark of the hot path: voxelize + MeanVFE + sparse 3D backbone, frames/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload kitti_b8|waymo_b4]

One "step" = one pass of the hot path over one batch of synthetic frames per GPU (BASELINE.json configs[1]:
KITTI fv2p.yaml backbone = VoxelResBackBone8x, batch 8, fp32).  Frames are independent, so N GPUs process N
independent batches (weak scaling, no data-path collective); only the elapsed time is reduced (MAX) over ranks.

`value`   device-timed throughput with the points already resident in HBM (CUDA events per step on the launching
          stream, L2 flushed between steps, untimed).
`e2e`     the same metric through HotPath(frames) with HOST buffers: pinned-memory H2D of the points, all kernels,
          D2H of the row counts and of the stride-8 output features inside the timed region.
`roofline` the dominant kernel (slowest launch of the conv layer shape with the largest share of the step), timed
          alone with CUDA events.
`cpu_baseline` the reference's compiled CPU path (oracle/_ref sparse_conv_ext) + the C port of its numba voxelizer,
          timed on this box's host cores on a bounded sample.  Only this leg and --impl reference execute oracle/.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "kitti_b8": dict(dataset="kitti", backbone="VoxelResBackBone8x", batch=8, split="test",
                     desc="KITTI fv2p.yaml (VoxelResBackBone8x), 8 synthetic frames/GPU, ~19k pts/frame"),
    "waymo_b4": dict(dataset="waymo", backbone="VoxelResBackBone8x", batch=4, split="test",
                     desc="Waymo waymo_fv2p_e30 (VoxelResBackBone8x), 4 synthetic frames/GPU, ~180k pts/frame"),
    "kitti_b8_plain": dict(dataset="kitti", backbone="VoxelBackBone8x", batch=8, split="test",
                           desc="KITTI VoxelBackBone8x, 8 synthetic frames/GPU"),
}
METRIC = "frames/sec voxelize+MeanVFE+VoxelResBackBone8x (fv2p.yaml backbone)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts + [time.perf_counter()])
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=3)
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[-1] <= t1]
            self.rows = inside if inside else self.rows[-3:]
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.rows))


def make_frames(wl, rank, n):
    from fv2p_b200 import synth
    return [synth.lidar_frame(wl["dataset"], seed=rank * 64 + i) for i in range(n)]


def build_model(wl, device, precision, use_graph=False):
    import torch
    import fv2p_b200
    from fv2p_b200 import synth
    cfg = synth.DATASETS[wl["dataset"]]
    net = getattr(fv2p_b200, wl["backbone"])({"PRECISION": precision}, cfg["num_point_features"],
                                             np.array(synth.grid_size(cfg))).eval()
    state = synth.randomize_state(net.state_dict(), seed=0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=False)
    net = net.to(device)
    hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points_per_voxel"],
                           cfg["max_voxels"][wl["split"]], use_graph=use_graph)
    return net, hp, state, cfg


# ------------------------------------------------------------------------------------------ ours
def layer_profile(hp, handle, flush):
    """Times every conv layer alone (CUDA events, L2 flushed before each) and returns per-layer records with
    algorithmic flops/bytes: flops = 2*P*Cin*Cout, bytes = Nin*Cin*e + Nout*Cout*e
    + K*Cin*Cout*e + 8*P (+ Nout*Cout*e when a residual is read)."""
    import torch
    from fv2p_b200 import _lib
    eng, a = hp.engine, handle["arena"]
    outs, n = eng.views(a, handle["vox"]["voxel_coords"], handle["batch"])
    prm = eng._prepare_params(a["device"])
    recs = []
    for i, (st, p) in enumerate(zip(eng.steps, prm)):
        nbr = eng.conv_operands(a, st, p)[0]
        pairs_total = int((nbr[:, :n[st.out_level]] >= 0).sum().item())
        src = handle["vox"]["voxel_features"] if st.in_buf < 0 else a["bufs"][st.in_buf]
        res = a["bufs"][st.res_buf] if st.res_buf is not None else None
        out = a["bufs"][st.out_buf]
        times = []
        for _ in range(3):
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run_conv_step(a, i, p, handle["vox"]["voxel_features"], handle["vox"]["cap"])
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        e_in = 4 if src.dtype == torch.float32 else 2
        e_out = 4 if out.dtype == torch.float32 else 2
        n_in, n_out = n[st.in_level], n[st.out_level]
        flops = 2.0 * pairs_total * st.cin * st.cout
        bytes_ = n_in * st.cin * e_in + n_out * st.cout * e_out + st.kvol * st.cin * st.cout * e_in + 8 * pairs_total \
            + (n_out * st.cout * e_out if res is not None else 0)
        recs.append(dict(layer=i, key=st.key, cin=st.cin, cout=st.cout, n_in=n_in, n_out=n_out, pairs=pairs_total,
                         mode=p["mode"], ms=min(times), flops=flops, bytes=bytes_))
    return recs


def run_ours(args, wl, rank, world, device):
    import torch
    import torch.distributed as dist
    from fv2p_b200 import _lib
    precision = args.precision
    _lib.load().fv2p_tc_gather_mode({"auto": -1, "lsu": 0, "tma": 1}[args.gather])
    sampler = ClockSampler(torch.cuda.current_device() if device.index is None else device.index)
    sampler.start()  # nvidia-smi takes a second to start streaming; rows are filtered to the timed region later
    net, hp, state, cfg = build_model(wl, device, precision, use_graph=not args.no_graph)
    frames = make_frames(wl, rank, wl["batch"])
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=device)

    def flush():
        flush_buf.zero_()

    pts, off, mfp, h2d_bytes = hp.upload(frames, device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, one event pair per step, L2 flushed (untimed) between steps
    def step():
        return hp.launch_graph() if hp.use_graph else hp.launch_resident(pts, off, mfp)

    for _ in range(args.warmup):
        handle = step()
        hp.finish(handle)
    barrier()
    step_ms = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        handle = step()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_s = time.perf_counter() - t_wall0
    outs, info = hp.finish(handle)
    counts = info["counts"]
    total_ms = sum(step_ms)

    # ---- end to end through the public API with host buffers (H2D + D2H inside the timed region):
    # HotPath.run_stream = the pipelined form of HotPath(frames, fetch="encoded"): per step one pinned H2D of the
    # points, one graph replay, D2H of the row counts + stride-8 features and indices; copies of step i overlap
    # the kernels of step i+1 (double buffered).  Every step is synchronised on the host when its result lands.
    for _ in hp.run_stream((frames for _ in range(max(2, args.warmup))), device):
        pass
    barrier()
    t0 = time.perf_counter()
    d2h_bytes = 0
    for res in hp.run_stream((frames for _ in range(args.steps)), device):
        d2h_bytes = res["d2h_bytes"]
        assert res["encoded_features"].shape[0] == res["counts"][-1]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the unpipelined call, for reference (one step fully serialised: upload -> kernels -> download)
    t1 = time.perf_counter()
    for _ in range(max(3, args.steps // 5)):
        hp(frames, device, fetch="encoded")
    torch.cuda.synchronize()
    e2e_serial_ms = (time.perf_counter() - t1) * 1000.0 / max(3, args.steps // 5)
    clocks = sampler.stop(t_wall0, time.perf_counter())

    # ---- reduce over ranks (MAX of elapsed)
    from fv2p_b200 import sharding
    total_ms = sharding.max_over_ranks(total_ms, device)
    e2e_ms = sharding.max_over_ranks(e2e_s * 1000.0, device)
    frames_total = wl["batch"] * args.steps * world
    value = frames_total / (total_ms / 1000.0)
    e2e_value = frames_total / (e2e_ms / 1000.0)

    if rank != 0:
        return None
    # ---- per-kernel evidence on rank 0
    handle = step()
    hp.finish(handle)
    recs = layer_profile(hp, handle, flush)
    pk = peaks()
    # dominant kernel = the layer shape (cin, cout, rulebook) that takes the largest share of the step, represented
    # by its slowest launch (stable from run to run, unlike the single slowest launch)
    fam = {}
    for r in recs:
        fam.setdefault((r["cin"], r["cout"], r["key"]), []).append(r)
    dom = max(max(fam.values(), key=lambda rs: sum(r["ms"] for r in rs)), key=lambda r: r["ms"])
    conv_ms = sum(r["ms"] for r in recs)
    tflops = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
    gbs = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
    ai = dom["flops"] / max(dom["bytes"], 1)
    ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    if ai >= ridge / 8:  # contraction-dominated layer: judge it against the tensor pipe
        roof = dict(bound="tensor", achieved=round(tflops, 3), peak=pk["bf16_tflops"], unit="TFLOP/s",
                    frac=round(tflops / pk["bf16_tflops"], 5))
    else:
        roof = dict(bound="hbm", achieved=round(gbs, 1), peak=pk["hbm_gbs"], unit="GB/s",
                    frac=round(gbs / pk["hbm_gbs"], 5))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):  # dram bytes per launch of the same kernel shape from the committed ncu --set full capture
        traffic = json.load(open(tpath)).get("%s:%s:%d:%d" % (args.workload, precision, dom["cin"], dom["cout"]))
    roof.update(traffic=traffic, peak_source=pk["source"],
                kernel="conv_fwd layer %d (%s, %d->%d, N_out=%d, pairs=%d, mode=%d)" % (
                    dom["layer"], dom["key"], dom["cin"], dom["cout"], dom["n_out"], dom["pairs"], dom["mode"]),
                kernel_ms=round(dom["ms"], 4), kernel_share_of_conv=round(dom["ms"] / conv_ms, 3),
                algorithmic_flops=dom["flops"], algorithmic_bytes=dom["bytes"])
    # ---- the HBM-bound stages, timed as groups with CUDA events (L2 flushed first)
    def timed(fn):
        best = 1e9
        for _ in range(3):
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    vox_ms = timed(lambda: hp.voxelizer(pts, off, mfp))
    vox = hp.voxelizer(pts, off, mfp)
    eng = hp.engine
    n0 = vox["voxel_offsets"][wl["batch"]:wl["batch"] + 1]
    geo_conv_ms = timed(lambda: eng.launch(vox["voxel_features"], vox["voxel_coords"], wl["batch"], n0_dev=n0,
                                           cap0=vox["cap"]))
    F = frames[0].shape[1]
    P = int(sum(f.shape[0] for f in frames))
    vox_bytes = P * F * 4 + counts[0] * (16 + F * 4 + 4)
    rb_bytes = 0
    for bk in eng.books:  # N*16 + K*2*N*4 + K*4 (+ Nout*16), plus the output-major map K*Nout*4 this design adds
        n_in, n_o = counts[bk.in_level], counts[bk.out_level]
        rb_bytes += n_in * 16 + bk.kvol * 2 * n_in * 4 + bk.kvol * 4 + (0 if bk.subm else n_o * 16) + bk.kvol * n_o * 4
    rb_ms = max(geo_conv_ms - conv_ms, 1e-3)
    # ---- the step right after the path (not part of `value`): HeightCompression of the stride-8
    # output, zero fill + scatter; algorithmic bytes = rows read + indices + the whole BEV map written
    from fv2p_b200.height_compression import height_compression
    enc = outs["out"]
    bev_shape = [int(v) for v in enc.spatial_shape]
    bev_buf = torch.empty((wl["batch"], enc.features.shape[1] * bev_shape[0], bev_shape[1], bev_shape[2]),
                          dtype=enc.features.dtype, device=device)
    bev_ms = timed(lambda: height_compression(enc.features, enc.indices, bev_shape, wl["batch"], out=bev_buf))
    bev_bytes = enc.features.numel() * enc.features.element_size() + enc.indices.numel() * 4 + \
        bev_buf.numel() * bev_buf.element_size()
    launches = hp.engine.launch_count() + 8
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if precision == "fp32" else "bf16",
        "data": "synthetic (seeded LiDAR-like frames, random-init weights)",
        "config": {"workload": args.workload, "description": wl["desc"], "frames_per_gpu_per_step": wl["batch"],
                   "precision": precision, "l2": "flushed between timed steps (512 MiB memset, untimed)",
                   "launch": "one CUDA graph replay per step" if hp.use_graph else "eager launches",
                   "parallelism": "frames sharded per GPU, no collective on the data path",
                   "rows_per_level": counts, "points_per_step": int(sum(f.shape[0] for f in frames))},
        "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": round(e2e_ms / args.steps, 4),
                "api": "HotPath.run_stream (double-buffered copies)", "serial_call_ms": round(e2e_serial_ms, 4)},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        "roofline": roof,
        "stages": {"voxelize_ms": round(vox_ms, 4), "voxelize_gbs": round(vox_bytes / vox_ms / 1e6, 1),
                   "voxelize_frac_of_hbm": round(vox_bytes / vox_ms / 1e6 / pk["hbm_gbs"], 4),
                   "rulebooks_ms": round(rb_ms, 4), "rulebooks_gbs": round(rb_bytes / rb_ms / 1e6, 1),
                   "rulebooks_frac_of_hbm": round(rb_bytes / rb_ms / 1e6 / pk["hbm_gbs"], 4),
                   "height_compression_ms": round(bev_ms, 4),
                   "height_compression_gbs": round(bev_bytes / bev_ms / 1e6, 1),
                   "height_compression_frac_of_hbm": round(bev_bytes / bev_ms / 1e6 / pk["hbm_gbs"], 4),
                   "conv_ms_sum": round(conv_ms, 4), "step_ms_median": round(statistics.median(step_ms), 4),
                   "host_wall_ms_per_step": round(wall_s * 1000 / args.steps, 3),
                   "total_gflop": round(sum(r["flops"] for r in recs) / 1e9, 3),
                   "total_compulsory_mb": round(sum(r["bytes"] for r in recs) / 1e6, 2),
                   "layers": [dict(l=r["layer"], c="%d>%d" % (r["cin"], r["cout"]), ms=round(r["ms"], 4),
                                   tf=round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 3),
                                   gbs=round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1)) for r in recs]},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(wl, steps=3, warmup=1, frames_per_step=1)
    return line


# ------------------------------------------------------------------------------------- reference arm
def cpu_reference(wl, steps, warmup, frames_per_step):
    """The reference's CPU implementation of the path on this box's host cores: its compiled extension
    (oracle/_ref, built from /root/reference by oracle/build_ref.py) driven with the call sequence of
    conv.py/spconv_backbone.py, torch CPU BatchNorm/ReLU, and the C port of the numba voxelizer + MeanVFE."""
    import torch
    from oracle import oracle as O
    from oracle import ref as R
    import fv2p_b200
    from fv2p_b200 import synth
    cfg = synth.DATASETS[wl["dataset"]]
    gs = synth.grid_size(cfg)
    shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
    net = getattr(fv2p_b200, wl["backbone"])({}, cfg["num_point_features"], np.array(gs))
    state = synth.randomize_state(net.state_dict(), seed=0)
    kind = "reference" if R.have_ext() else "port"
    params_t = {k: torch.from_numpy(v) for k, v in state.items()}
    frames = make_frames(wl, 0, frames_per_step)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    timers = {}

    def one_step():
        t0 = time.perf_counter()
        feats, coords = [], []
        for b, f in enumerate(frames):
            v, c, n = O.voxelize(f, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points_per_voxel"],
                                 cfg["max_voxels"][wl["split"]])
            feats.append(O.mean_vfe(v, n))
            coords.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], 1))
        t1 = time.perf_counter()
        feats, coords = np.concatenate(feats), np.concatenate(coords)
        if kind == "reference":
            R.ext_backbone_forward(wl["backbone"], params_t, feats, coords, len(frames), shape, timers=timers)
        else:
            O.backbone_forward(wl["backbone"], state, feats, coords, len(frames), shape)
        timers["voxelize_s"] = timers.get("voxelize_s", 0.0) + (t1 - t0)
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    timers.clear()
    times = [one_step() for _ in range(steps)]
    total = sum(times)
    return dict(value=round(frames_per_step * steps / total, 4), unit="frames/s", cores=cores if kind == "reference" else 1,
                kind=kind, ms_per_frame=round(1000 * total / (frames_per_step * steps), 1),
                sample="%d step(s) x %d frame(s) of %s, batch %d per call; backbone via the reference's compiled "
                       "sparse_conv_ext CPU path (MKL GEMMs use %d threads, rulebook/gather/scatter are serial there), "
                       "voxelizer+MeanVFE via the C port of the numba kernel (1 thread)" % (
                           steps, frames_per_step, wl["desc"], frames_per_step, torch.get_num_threads()),
                breakdown_ms_per_frame={k[:-2]: round(1000 * v / (frames_per_step * steps), 1)
                                        for k, v in timers.items()})


def reference_gpu(wl, device, steps=5, warmup=2, frames=None):
    """The reference's OWN CUDA path on this GPU: oracle/_ref/sparse_conv_ext was
    built from /root/reference with -DWITH_CUDA for sm_100, so the call sequence of conv.py / spconv_backbone.py on
    CUDA tensors runs its kernels (src/indice_cuda.cu:30-135 rulebooks on a dense int32 grid + torch::_unique,
    src/reordering_cuda.cu:31-140 gather / scatter-add, cuBLAS sgemm through torch::mm_out, one host sync per conv)
    with torch CUDA BatchNorm/ReLU.  Voxels are made on the CPU beforehand (the reference voxelizes in DataLoader
    workers) and are resident when the timed region starts, so this is the backbone only - the like-for-like bar for
    our feature + geometry pass.  Timed with CUDA events around each forward."""
    import torch
    from oracle import oracle as O
    from oracle import ref as R
    import fv2p_b200
    from fv2p_b200 import synth
    if not R.have_ext():
        return {"unavailable": "oracle/_ref/sparse_conv_ext.so not present"}
    cfg = synth.DATASETS[wl["dataset"]]
    gs = synth.grid_size(cfg)
    shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
    net = getattr(fv2p_b200, wl["backbone"])({}, cfg["num_point_features"], np.array(gs))
    state = synth.randomize_state(net.state_dict(), seed=0)
    params = {k: torch.from_numpy(v).to(device) for k, v in state.items()}
    frames = frames if frames is not None else make_frames(wl, 0, wl["batch"])
    feats, coords = [], []
    for b, f in enumerate(frames):
        v, c, n = O.voxelize(f, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points_per_voxel"],
                             cfg["max_voxels"][wl["split"]])
        feats.append(O.mean_vfe(v, n))
        coords.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], 1))
    feats = torch.from_numpy(np.concatenate(feats)).to(device)
    coords = torch.from_numpy(np.concatenate(coords)).to(device)
    try:
        times = []
        for i in range(warmup + steps):
            timers = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            R.ext_backbone_forward(wl["backbone"], params, feats, coords, len(frames), shape, timers=timers,
                                   device=device)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                times.append(e0.elapsed_time(e1))
    except Exception as e:  # the reference's kernels are not ours to fix
        return {"unavailable": "reference CUDA path failed on this device: %s" % (str(e).splitlines()[0][:200])}
    ms = statistics.median(times)
    return dict(value=round(len(frames) / (ms * 1e-3), 2), unit="frames/s", ms_per_step=round(ms, 3),
                frames_per_step=len(frames), steps=steps,
                scope="MeanVFE output resident -> %s forward (rulebooks + convs + BN/ReLU), voxelization not included "
                      "(the reference voxelizes on the CPU)" % wl["backbone"],
                kind="reference sparse_conv_ext CUDA kernels (oracle/_ref, sm_100) + cuBLAS + torch CUDA BN/ReLU",
                host_rulebook_ms=round(1000 * timers.get("rulebook_s", 0.0), 2),
                host_conv_ms=round(1000 * timers.get("conv_s", 0.0), 2))


def run_reference(args, wl, rank, world):
    if rank != 0:
        return None
    frames_per_step = 2
    res = cpu_reference(wl, steps=args.steps, warmup=min(args.warmup, 1), frames_per_step=frames_per_step)
    total_s = frames_per_step * args.steps / res["value"]
    return {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000 * total_s / args.steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"],
                   "note": "reference CPU path, bounded sample of %d frames per step" % frames_per_step},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti_b8", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "lsu", "tma"],
                    help="A-tile producer of the tensor-core conv (fv2p_tc_gather_mode); auto = measured best per shape")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, for one) goes
    # to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        line = run_reference(args, wl, rank, world)
        if line is not None:
            print(json.dumps(line), file=json_out, flush=True)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    line = run_ours(args, wl, rank, world, device)
    if line is not None:
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

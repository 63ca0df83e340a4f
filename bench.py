#!/usr/bin/env python
"""Benchmark of the hot path: voxelize + MeanVFE + sparse 3D backbone, frames/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]
                    [--workload auto|waymo_b4|waymo_64|kitti_b8|kitti_b8_plain|waymo_b16|kitti_b64|micro]

Workload (BASELINE.json `configs`): one GPU -> `waymo_b4` (configs[2]: Waymo-shaped frames, waymo_fv2p_e30 backbone =
VoxelResBackBone8x, 4 frames per step, fp32): the north-star shape and the largest single-GPU configuration.  N > 1
(torchrun) -> `waymo_64` (configs[3]): a FIXED batch of 64 Waymo-shaped frames sharded per frame over the ranks
(fv2p_b200.sharding), every rank running its 64/N frames in the same 4-frame launches, so the one-GPU default is the
N = 1 point of the same curve; `scaling: strong`, no collective on the data path, only the elapsed time is reduced
(MAX over ranks).  `kitti_b8` (configs[1]) fp32/bf16, the other precision of the headline workload and two larger
launches (`waymo_b16`, `kitti_b64`) ride along as `other_workloads` at N = 1; `--workload micro` is configs[4].

One "step" = one pass of the hot path over the workload's frames.
`value`   device-timed throughput with the points already resident in HBM (CUDA events per step on the launching
          stream, one CUDA-graph replay per launch, L2 flushed between steps, untimed).
`e2e`     the same metric through HotPath.run_stream with HOST buffers: pinned-memory H2D of the points, all kernels,
          D2H of the row counts and of the stride-8 output features inside the timed region; `serial_call_ms` is the
          plain synchronous call hp(frames, fetch="encoded").
`roofline` the dominant kernel (slowest launch of the conv layer shape with the largest share of the step), timed
          alone with CUDA events; `traffic` = dram bytes of that very layer from the committed ncu capture.
`stages`  voxelizer, the whole geometry pass (engine.launch(run_convs=False)) and HeightCompression, each timed
          directly, with achieved GB/s against the measured HBM peak; per-layer conv times.
`reference_gpu` the reference's OWN CUDA kernels (oracle/_ref, built from /root/reference for sm_100) on the same
          frames and GPU - the like-for-like bar.
`cpu_baseline` the reference's compiled CPU path (oracle/_ref sparse_conv_ext) + the C port of its numba voxelizer,
          timed on this box's host cores on a bounded sample.  Only these two legs and --impl reference execute oracle/.
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
    # BASELINE.json configs[2]: the north-star shape and the default at one GPU
    "waymo_b4": dict(dataset="waymo", backbone="VoxelResBackBone8x", batch=4, frames=4, split="test",
                     desc="Waymo waymo_fv2p_e30 (VoxelResBackBone8x), 4 synthetic frames per step, ~180k pts/frame"),
    # configs[3]: a fixed batch of 64 Waymo-shaped frames sharded per frame over the ranks (strong scaling); every rank
    # runs its shard in the same 4-frame launches as waymo_b4, so the one-GPU default is the N = 1 point of this curve
    "waymo_64": dict(dataset="waymo", backbone="VoxelResBackBone8x", batch=4, frames=64, split="test",
                     desc="Waymo-shaped batch of 64 frames sharded per frame over the GPUs, 4-frame launches "
                          "(VoxelResBackBone8x)"),
    # configs[1]
    "kitti_b8": dict(dataset="kitti", backbone="VoxelResBackBone8x", batch=8, frames=8, split="test",
                     desc="KITTI fv2p.yaml (VoxelResBackBone8x), 8 synthetic frames per step, ~19k pts/frame"),
    "kitti_b8_plain": dict(dataset="kitti", backbone="VoxelBackBone8x", batch=8, frames=8, split="test",
                           desc="KITTI VoxelBackBone8x, 8 synthetic frames per step"),
    # larger launches (latency amortised): what the arena holds in one step
    "waymo_b16": dict(dataset="waymo", backbone="VoxelResBackBone8x", batch=16, frames=16, split="test",
                      desc="Waymo-shaped, 16 frames per step"),
    "kitti_b64": dict(dataset="kitti", backbone="VoxelResBackBone8x", batch=64, frames=64, split="test",
                      desc="KITTI-shaped, 64 frames per step"),
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


def make_frames(wl, first, n):
    """Frames first .. first+n-1 of the workload's fixed synthetic sequence (seed = frame number)."""
    from fv2p_b200 import synth
    return [synth.lidar_frame(wl["dataset"], seed=first + i) for i in range(n)]


SORT_ROWS = True  # --group-rows: which rulebooks get a grouped row order (engine.BackboneEngine sort_rows)


def build_model(wl, device, precision, use_graph=False):
    import torch
    import fv2p_b200
    from fv2p_b200 import synth
    cfg = synth.DATASETS[wl["dataset"]]
    net = getattr(fv2p_b200, wl["backbone"])({"PRECISION": precision, "SORT_ROWS": SORT_ROWS},
                                             cfg["num_point_features"],
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


def _batches_for_rank(wl, rank, world):
    """This rank's share of the workload's frames, cut into launches of wl['batch'] frames."""
    from fv2p_b200 import sharding
    lo, hi = sharding.shard_range(wl["frames"], rank, world)
    frames = make_frames(wl, lo, hi - lo)
    b = wl["batch"]
    return [frames[i:i + b] for i in range(0, len(frames), b)], (lo, hi)


def measure(hp, batches, device, steps, warmup, flush, barrier):
    """Device-timed (points resident, one CUDA-graph replay per launch, L2 flushed before every step) and end-to-end
    (HotPath.run_stream from host buffers) over `steps` passes of the rank's batches.  Returns a dict of raw timings."""
    import torch
    # the largest batch first, so that every staging slot / the arena is sized once
    order = sorted(range(len(batches)), key=lambda j: -sum(f.shape[0] for f in batches[j]))
    h2d = 0
    for j in order:
        h2d += hp.upload(batches[j], device, slot=j)[3]
    torch.cuda.synchronize()

    # a step of several launches alternates them between two engine lanes on two streams, like HotPath.run_stream: the
    # head of launch j+1 (voxelizer, first rulebooks - small dependent kernels) runs under the convolutions of launch j
    n_lanes = min(2, len(batches), hp.lanes)
    main = torch.cuda.current_stream(device)
    lane_streams = [main] + [torch.cuda.Stream(device=device) for _ in range(n_lanes - 1)]

    def one_pass():
        handle = None
        if n_lanes > 1:
            fork = torch.cuda.Event()
            fork.record(main)
            for st in lane_streams[1:]:
                st.wait_event(fork)
        for j in range(len(batches)):
            lane = j % n_lanes
            with torch.cuda.stream(lane_streams[lane]):
                handle = hp.launch_graph(slot=j, lane=lane) if hp.use_graph else \
                    hp.launch_resident(*hp.staged(j), lane=lane)
        for st in lane_streams[1:]:
            main.wait_stream(st)
        return handle

    for _ in range(warmup):
        hp.finish(one_pass())
    barrier()
    step_ms = []
    t0 = time.perf_counter()
    for _ in range(steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        handle = one_pass()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_s = time.perf_counter() - t0
    outs, info = hp.finish(handle)
    # end to end through the public API with host buffers: per launch one pinned H2D of the points, one graph replay,
    # D2H of the row counts + stride-8 features and indices; copies of launch i overlap the kernels of launch i+1.
    # Every launch is synchronised on the host when its result lands.
    depth = int(os.environ.get("FV2P_STREAM_DEPTH", "2"))
    for _ in hp.run_stream((b for _ in range(max(2 * depth, warmup)) for b in batches), device, depth):  # slots x lanes of graphs
        pass
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for res in hp.run_stream((b for _ in range(steps) for b in batches), device, depth):
        d2h += res["d2h_bytes"]
        assert res["encoded_features"].shape[0] == res["counts"][-1]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t1
    return dict(step_ms=step_ms, total_ms=sum(step_ms), wall_s=wall_s, e2e_s=e2e_s, h2d_bytes_per_step=h2d,
                d2h_bytes_per_step=d2h // max(steps, 1), handle=handle, counts=info["counts"])


def serial_call_ms(hp, frames, device, n=5):
    """The plain call a detector makes, fully serialised: upload -> kernels -> download of the stride-8 result.
    Returns (ms per call, breakdown in ms of one call taken apart with a synchronise after every phase)."""
    import torch
    hp(frames, device, fetch="encoded")
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        hp(frames, device, fetch="encoded")
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) * 1000.0 / n
    t0 = time.perf_counter()
    pts, off, mfp, _ = hp.upload(frames, device)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    handle = hp.launch_graph() if hp.use_graph else hp.launch_resident(pts, off, mfp)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    hp.finish(handle, "encoded")
    t5 = time.perf_counter()
    parts = dict(pack_and_enqueue_h2d=t1 - t0, h2d_wait=t2 - t1, launch=t3 - t2, kernels_wait=t4 - t3,
                 counts_and_d2h=t5 - t4)
    return ms, {k: round(v * 1e3, 3) for k, v in parts.items()}


def quick_value(name, precision, device, steps, flush):
    """frames/s (device-timed and end to end) of another workload on this GPU, for the extra keys of the JSON line."""
    import torch
    wl = WORKLOADS[name]
    net, hp, state, cfg = build_model(wl, device, precision, use_graph=True)
    batches, _ = _batches_for_rank(wl, 0, 1)
    m = measure(hp, batches, device, steps, 3, flush, torch.cuda.synchronize)
    n = wl["frames"] * steps
    out = dict(value=round(n / (m["total_ms"] / 1e3), 2), e2e=round(n / m["e2e_s"], 2),
               ms_per_step=round(m["total_ms"] / steps, 4), frames_per_step=wl["frames"], precision=precision,
               arena_mb=round(hp.engine.arena_bytes() / 1e6, 1), rows_per_level=m["counts"])
    del hp, net
    torch.cuda.empty_cache()
    return out


def run_ours(args, wl, rank, world, device):
    import torch
    import torch.distributed as dist
    from fv2p_b200 import _lib, sharding
    precision = args.precision
    _lib.load().fv2p_tc_gather_mode({"auto": -1, "lsu": 0, "tma": 1}[args.gather])
    sampler = ClockSampler(torch.cuda.current_device() if device.index is None else device.index)
    sampler.start()  # nvidia-smi takes a second to start streaming; rows are filtered to the timed region later
    net, hp, state, cfg = build_model(wl, device, precision, use_graph=not args.no_graph)
    batches, (lo, hi) = _batches_for_rank(wl, rank, world)
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=device)

    def flush():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_wall0 = time.perf_counter()
    m = measure(hp, batches, device, args.steps, args.warmup, flush, barrier)
    clocks = sampler.stop(t_wall0, time.perf_counter())
    handle, counts, step_ms = m["handle"], m["counts"], m["step_ms"]

    # ---- reduce over ranks (MAX of elapsed); value = frames of ALL ranks / that time
    total_ms = sharding.max_over_ranks(m["total_ms"], device)
    e2e_ms = sharding.max_over_ranks(m["e2e_s"] * 1000.0, device)
    frames_total = wl["frames"] * args.steps
    value = frames_total / (total_ms / 1000.0)
    e2e_value = frames_total / (e2e_ms / 1000.0)
    if rank != 0:
        return None

    # ---- per-kernel evidence on rank 0 (last launch of the pass: its arena is still in place)
    frames = batches[-1]
    pts, off, mfp = hp.staged(len(batches) - 1)
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
    # roofline model: attainable = min(tensor peak, AI x HBM peak); the layer is judged against whichever bounds it
    if ai >= ridge:
        roof = dict(bound="tensor", achieved=round(tflops, 3), peak=pk["bf16_tflops"], unit="TFLOP/s",
                    frac=round(tflops / pk["bf16_tflops"], 5))
    else:
        roof = dict(bound="hbm", achieved=round(gbs, 1), peak=pk["hbm_gbs"], unit="GB/s",
                    frac=round(gbs / pk["hbm_gbs"], 5))
    # dram bytes of THIS layer of THIS workload from the committed `ncu --set full` capture (profiles/r2_traffic.json,
    # key workload:precision:layer), per launch like `achieved`; null when that layer was not captured
    traffic = None
    wkey = "waymo_b4" if wl["dataset"] == "waymo" and wl["batch"] == 4 else args.workload
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s:%s:l%d" % (wkey, precision, dom["layer"]))
    roof.update(traffic=traffic, peak_source=pk["source"], arithmetic_intensity=round(ai, 1), ridge=round(ridge, 1),
                achieved_tflops=round(tflops, 3), achieved_gbs=round(gbs, 1),
                kernel="conv_fwd layer %d (%s, %d->%d, N_out=%d, pairs=%d, mode=%d)" % (
                    dom["layer"], dom["key"], dom["cin"], dom["cout"], dom["n_out"], dom["pairs"], dom["mode"]),
                kernel_ms=round(dom["ms"], 4), kernel_share_of_conv=round(dom["ms"] / conv_ms, 3),
                algorithmic_flops=dom["flops"], algorithmic_bytes=dom["bytes"],
                tensor_flops_note="fp32 mode issues 3 bf16 tensor products per algorithmic product (hi*hi + hi*lo + "
                                  "lo*hi of the split operands): the tensor pipe does three times the algorithmic "
                                  "flops" if precision == "fp32" else None)

    # ---- the HBM-class stages, each timed directly with CUDA events (L2 flushed first, best of 3)
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
    nb = len(frames)
    eng = hp.engine
    # as in HotPath.launch_resident: the voxelizer also fills the convolutions' level-0 coordinate table
    table0 = eng.level0_table(pts.device, hp.voxelizer.capacity(pts.device, pts.shape[0], nb, pts.shape[1]), nb)
    vox_ms = timed(lambda: hp.voxelizer(pts, off, mfp, level0_table=table0))
    vox = hp.voxelizer(pts, off, mfp, level0_table=table0)
    n0 = vox["voxel_offsets"][nb:nb + 1]
    geo_ms = timed(lambda: eng.launch(vox["voxel_features"], vox["voxel_coords"], nb, n0_dev=n0, cap0=vox["cap"],
                                      run_convs=False, table0_built=True))
    F = frames[0].shape[1]
    P = int(sum(f.shape[0] for f in frames))
    vox_bytes = P * F * 4 + counts[0] * (16 + F * 4 + 4)
    rb_bytes = 0
    for d in hp.engine.arena["geo"]:  # SURVEY 8d per rulebook: N*16 read + (Nout*16) + the output-major map K*Nout*4
        bk = d["book"]                # written, read and written again grouped (+ Nout*4 of row order) when grouped
        n_in, n_o = counts[bk.in_level], counts[bk.out_level]
        rb_bytes += n_in * 16 + (0 if bk.subm else n_o * 16) + bk.kvol * n_o * 4 * (3 if d["sorted"] else 1) + \
            (n_o * 4 if d["sorted"] else 0)
    # ---- the step right after the path (not part of `value`): HeightCompression of the stride-8
    # output; algorithmic bytes = rows read + indices + the whole BEV map written
    from fv2p_b200.height_compression import height_compression
    outs, _ = hp.finish(handle)
    enc = outs["out"]
    bev_shape = [int(v) for v in enc.spatial_shape]
    bev_buf = torch.empty((nb, enc.features.shape[1] * bev_shape[0], bev_shape[1], bev_shape[2]),
                          dtype=enc.features.dtype, device=device)
    bev_ms = timed(lambda: height_compression(enc.features, enc.indices, bev_shape, nb, out=bev_buf))
    bev_bytes = enc.features.numel() * enc.features.element_size() + enc.indices.numel() * 4 + \
        bev_buf.numel() * bev_buf.element_size()
    del bev_buf
    launches_per_pass = (hp.engine.launch_count(table0_built=True) + hp.voxelizer.LAUNCHES) * len(batches)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if wl["frames"] > wl["batch"] or world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32" if precision == "fp32" else "bf16",
        "data": "synthetic (seeded LiDAR-like frames, random-init weights)",
        "config": {"workload": args.workload, "description": wl["desc"], "frames_per_step": wl["frames"],
                   "frames_per_launch": wl["batch"], "launches_per_gpu_per_step": len(batches),
                   "frames_of_rank0": [lo, hi], "precision": precision,
                   "l2": "flushed between timed steps (512 MiB memset, untimed)",
                   "launch": ("one CUDA graph replay per launch" if hp.use_graph else "eager launches") +
                             ("; launches alternate between two engine lanes on two streams" if len(batches) > 1 else ""),
                   "parallelism": "frames sharded per GPU (strong scaling of the fixed batch), no collective on the "
                                  "data path" if world > 1 else "one GPU",
                   "rows_per_level_last_launch": counts, "points_last_launch": P},
        "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": int(m["h2d_bytes_per_step"]),
                "d2h_bytes_per_step": int(m["d2h_bytes_per_step"]), "ms_per_step": round(e2e_ms / args.steps, 4),
                "api": "HotPath.run_stream (double-buffered copies, two engine lanes)"},
        "gpu_launches": launches_per_pass * args.steps,
        "clocks": clocks,
        "roofline": roof,
        "stages": {"note": "per launch of %d frames; every entry timed on its own with CUDA events" % nb,
                   "voxelize_ms": round(vox_ms, 4), "voxelize_gbs": round(vox_bytes / vox_ms / 1e6, 1),
                   "voxelize_frac_of_hbm": round(vox_bytes / vox_ms / 1e6 / pk["hbm_gbs"], 4),
                   "rulebooks_ms": round(geo_ms, 4), "rulebooks_gbs": round(rb_bytes / geo_ms / 1e6, 1),
                   "rulebooks_frac_of_hbm": round(rb_bytes / geo_ms / 1e6 / pk["hbm_gbs"], 4),
                   "rulebook_launches": hp.engine.launch_count(table0_built=True) - len(hp.engine.steps),
                   "voxelize_launches": hp.voxelizer.LAUNCHES,
                   "height_compression_ms": round(bev_ms, 4),
                   "height_compression_gbs": round(bev_bytes / bev_ms / 1e6, 1),
                   "height_compression_frac_of_hbm": round(bev_bytes / bev_ms / 1e6 / pk["hbm_gbs"], 4),
                   "conv_ms_sum": round(conv_ms, 4), "step_ms_median": round(statistics.median(step_ms), 4),
                   "host_wall_ms_per_step": round(m["wall_s"] * 1000 / args.steps, 3),
                   "total_gflop": round(sum(r["flops"] for r in recs) / 1e9, 3),
                   "total_compulsory_mb": round(sum(r["bytes"] for r in recs) / 1e6, 2),
                   "arena_mb_per_lane": round(hp.engine.arena_bytes() / 1e6, 1),
                   "layers": [dict(l=r["layer"], c="%d>%d" % (r["cin"], r["cout"]), ms=round(r["ms"], 4),
                                   tf=round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 3),
                                   gbs=round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1)) for r in recs]},
    }
    if world == 1:
        ms, parts = serial_call_ms(hp, frames, device)
        line["e2e"]["serial_call_ms"] = round(ms, 4)
        line["e2e"]["serial_call_frames_per_s"] = round(nb / (ms / 1e3), 2)
        line["e2e"]["serial_call_breakdown_ms"] = parts
    if world == 1 and not args.no_extras:
        del hp, net
        torch.cuda.empty_cache()
        # the reference's own CUDA path on the same frames and GPU (the like-for-like bar; backbone only)
        line["reference_gpu"] = reference_gpu(wl, device, frames=frames)
        # the other configurations of BASELINE.json on this GPU (same code path, fewer steps)
        others = {}
        other = "bf16" if precision == "fp32" else "fp32"
        for name, prec in ((args.workload, other), ("kitti_b8", "fp32"), ("kitti_b8", "bf16"), ("waymo_b16", precision),
                           ("kitti_b64", precision)):
            if (name, prec) == (args.workload, precision):
                continue
            try:
                others["%s:%s" % (name, prec)] = quick_value(name, prec, device, max(5, args.steps // 4), flush)
            except Exception as e:  # an extra key must not take the headline down with it
                others["%s:%s" % (name, prec)] = {"error": str(e).splitlines()[0][:200]}
        line["other_workloads"] = others
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(wl, steps=3, warmup=1, frames_per_step=1)
    return line


# --------------------------------------------------------------------------------------------- micro (configs[4])
def run_micro(args, device):
    """Rulebook + fused conv microbenchmark (BASELINE.json configs[4]): 10k - 1M active voxels (surface-like: voxels of
    1..13 synthetic Waymo frames, or the first 10k of one), 3x3x3 submanifold and stride-2 rulebooks, channels 16-128,
    fp32 and bf16.  Every number is one stream-ordered call timed with CUDA events, best of 3, L2 flushed first."""
    import torch
    from fv2p_b200 import _lib, spconv, synth
    from oracle import oracle as O  # only to voxelize the synthetic frames on the host (not timed)
    cfg = synth.DATASETS["waymo"]
    gs = synth.grid_size(cfg)
    shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
    per_frame = []
    for i in range(13):
        _, c, _ = O.voxelize(synth.lidar_frame("waymo", seed=i), cfg["voxel_size"], cfg["point_cloud_range"], 5, 90000)
        per_frame.append(np.concatenate([np.full((c.shape[0], 1), i, np.int32), c], 1))
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=device)
    pk = peaks()

    def timed(fn):
        best = 1e9
        for _ in range(3):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    rows = []
    for target in (10000, 100000, 1000000):
        ind, nf = [], 0
        while sum(x.shape[0] for x in ind) < target:
            ind.append(per_frame[nf])
            nf += 1
        ind = np.concatenate(ind)[:target]
        batch = int(ind[:, 0].max()) + 1
        ind_t = torch.from_numpy(np.ascontiguousarray(ind)).to(device)
        for kind, stride in (("subm", 1), ("conv_s2", 2)):
            subm = kind == "subm"
            call = lambda: spconv.ops.get_indice_pairs(ind_t, batch, shape, 3, stride, 1, 1, 0, subm, False,
                                                       return_nbr=True, want_pairs=False)
            outids, _, _, nbr = call()
            rb_ms = timed(call)
            n_out = outids.shape[0]
            nbr = nbr.contiguous()
            grp = lambda: spconv.ops.sort_rows_by_mask(nbr, n_out, return_tile_order=True)
            perm, nbr_sorted, order = grp()
            grp_ms = timed(grp)
            pairs = int((nbr >= 0).sum().item())
            rb_bytes = ind.shape[0] * 16 + (0 if subm else n_out * 16) + 27 * n_out * 4
            rec = dict(n_in=int(ind.shape[0]), kind=kind, n_out=int(n_out), pairs=pairs, rulebook_ms=round(rb_ms, 4),
                       rulebook_gbs=round(rb_bytes / rb_ms / 1e6, 1), group_rows_ms=round(grp_ms, 4), conv={})
            for ch in (16, 32, 64, 128):
                for prec, mode, dt in (("fp32", _lib.MODE_TF32X3_TC, torch.float32),
                                       ("bf16", _lib.MODE_BF16_TC, torch.bfloat16)):
                    feats = torch.randn(ind.shape[0], ch, device=device).to(dt)
                    w = (torch.randn(27, ch, ch, device=device) / (27 * ch) ** 0.5)
                    packed = spconv.ops.pack_weight(w, mode)
                    out = torch.empty((n_out, ch), dtype=dt, device=device)
                    fn = lambda: spconv.ops.conv_forward(feats, packed, nbr_sorted.contiguous(), n_out, relu=True,
                                                         mode=mode, row_perm=perm.contiguous(),
                                                         tile_order=order.contiguous(), out=out)
                    fn()
                    ms = timed(fn)
                    e = 4 if prec == "fp32" else 2
                    by = ind.shape[0] * ch * e + n_out * ch * e + 27 * ch * ch * e + 8 * pairs
                    rec["conv"]["%d:%s" % (ch, prec)] = dict(ms=round(ms, 4),
                                                            tflops=round(2.0 * pairs * ch * ch / ms / 1e9, 2),
                                                            gbs=round(by / ms / 1e6, 1))
            rows.append(rec)
    return {"metric": "rulebook + fused sparse conv microbenchmark (BASELINE.json configs[4])", "unit": "ms per call",
            "value": rows[-2]["rulebook_ms"], "n_gpus": 1, "higher_is_better": False, "dtype": "int32 / f32 / bf16",
            "data": "synthetic (voxels of 1-13 Waymo-shaped frames)", "config": {"workload": "micro"},
            "peaks": {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops": pk["bf16_tflops"], "source": pk["source"]},
            "rows": rows}


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
        "higher_is_better": True, "scaling": "strong" if wl["frames"] > wl["batch"] or world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"],
                   "note": "reference CPU path, bounded sample of %d frames per step" % frames_per_step},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "micro"] + sorted(WORKLOADS),
                    help="auto = waymo_b4 on one GPU (BASELINE.json configs[2]), waymo_64 sharded over the ranks "
                         "otherwise (configs[3]); micro = rulebook + conv microbenchmark (configs[4])")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip reference_gpu and the other workloads' extra keys (profiling runs)")
    ap.add_argument("--gather", default="auto", choices=["auto", "lsu", "tma"],
                    help="A-tile producer of the tensor-core conv (fv2p_tc_gather_mode); auto = TMA gather4, cp.async for packed stages")
    ap.add_argument("--group-rows", default="all", choices=["all", "subm", "none"],
                    help="rulebooks whose rows are grouped by neighbour-mask digest for the tensor-core conv")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    global SORT_ROWS
    SORT_ROWS = {"all": True, "subm": "subm", "none": False}[args.group_rows]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "auto":
        args.workload = "waymo_b4" if world == 1 else "waymo_64"

    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, for one) goes
    # to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        if args.workload == "micro":
            args.workload = "waymo_b4"
        line = run_reference(args, WORKLOADS[args.workload], rank, world)
        if line is not None:
            print(json.dumps(line), file=json_out, flush=True)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if args.workload == "micro":
        if rank == 0:
            print(json.dumps(run_micro(args, device)), file=json_out, flush=True)
        return 0
    wl = WORKLOADS[args.workload]
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

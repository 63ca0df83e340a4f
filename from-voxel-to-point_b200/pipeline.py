"""HotPath: raw points -> voxelize + MeanVFE -> sparse backbone, as one stream-ordered step.

This is the call a user of the path makes (the reference spreads it over DataLoader workers running the
numba voxelizer, collate_batch, load_data_to_gpu, MeanVFE.forward and the backbone forward:
pcdet/datasets/dataset.py:137-183, pcdet/models/__init__.py:15-21, detectors/detector3d_template.py:22-25).

    hp = HotPath(backbone, voxel_size, point_cloud_range, max_points_per_voxel, max_voxels)
    batch_dict, info = hp(list_of_host_point_arrays)    # H2D of the points, all kernels, D2H of counts (+ features)
    for result in hp.run_stream(iterable_of_batches):   # same, pipelined: copies of batch i overlap batch i+1

`upload` / `launch_resident` / `launch_graph` / `finish` split the step for callers that keep the points in HBM
and for the benchmark.  Between the H2D copy and the final D2H copy nothing synchronises with the host.

`run_stream` keeps `lanes` batches in flight on the device, each with its own engine state (arena, rulebooks,
voxel buffers) and stream: the geometry of batch i+1 (voxelize, rulebooks, sorts - small latency-bound kernels)
runs under the feature pass of batch i instead of in front of its own.
"""
import numpy as np
import torch

from . import _lib
from .engine import BackboneEngine, CapacityOverflow
from .voxel_generator import BatchVoxelizer


class HotPath(object):
    def __init__(self, backbone, voxel_size, point_cloud_range, max_points_per_voxel, max_voxels,
                 precision=None, use_graph=False, lanes=2, bev=False):
        self.backbone = backbone
        if precision is not None:
            try:
                backbone.model_cfg['PRECISION'] = precision
            except TypeError:
                setattr(backbone.model_cfg, 'PRECISION', precision)
        # own engine(s): the pair tensors of indice_dict are built only when somebody reads them (the module API keeps
        # building them with every step unless MATERIALIZE_PAIRS says otherwise)
        self.engine = backbone.make_engine(materialize_pairs='lazy')
        self.voxelizer = BatchVoxelizer(voxel_size, point_cloud_range, max_points_per_voxel, max_voxels)
        # lane 0 is the engine the backbone module itself uses; further lanes are built on first use by run_stream
        self._vox_args = (voxel_size, point_cloud_range, max_points_per_voxel, max_voxels)
        self._lanes = [(self.engine, self.voxelizer)]
        self._lane_streams = {}
        self.lanes = max(1, int(lanes))
        self.bev = bool(bev)  # append HeightCompression (the BEV map of the stride-8 output) to the step
        self._bev_bufs = {}
        self.use_graph = use_graph
        self._slots = {}
        self._graphs = {}
        self._copy_stream = None
        self._copy_in_stream = None
        self._enc_host = None

    # ------------------------------------------------------------------ host -> device
    def _slot(self, device, total_points, batch, f, index=0):
        """Staging slot `index`: pinned host + device buffers for the points of one batch (grow-only)."""
        s = self._slots.get(index)
        if s is None or s["pcap"] < total_points or s["batch"] != batch or s["f"] != f or s["device"] != device:
            # all slots share one capacity so that they share the voxelizer / arena sizing (and graph shapes)
            pcap = max([int(total_points * 1.1) + 1024] + [x["pcap"] for k, x in self._slots.items()
                                                           if isinstance(k, int) and x["batch"] == batch and x["f"] == f])
            s = dict(pcap=pcap, batch=batch, f=f, device=device, index=index,
                     host_pts=torch.empty((pcap, f), dtype=torch.float32).pin_memory(),
                     host_off=torch.zeros((batch + 1,), dtype=torch.int32).pin_memory(),
                     dev_pts=torch.empty((pcap, f), dtype=torch.float32, device=device),
                     # empty, not zeros: a fill kernel would run on whatever stream is current here, unordered with
                     # the H2D copy that upload() enqueues on the copy stream right afterwards (the copy writes every
                     # entry); with the fill landing second a batch saw all-zero frame offsets = no points
                     dev_off=torch.empty((batch + 1,), dtype=torch.int32, device=device))
            self._slots[index] = s
        return s

    @property
    def _stage(self):
        return self._slots.get(0)

    def upload(self, frames, device="cuda", slot=0, stream=None, after=None):
        """frames: list of [P_b,F] float32 numpy arrays (or CPU tensors).  Packs them into pinned memory and
        enqueues the H2D copies (on `stream` if given, there after the event `after`: whatever still reads the slot's
        device buffer).  Returns (points_dev [P,F], frame_offsets_dev [B+1], max_frame_points, bytes)."""
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        sizes = [int(fr.shape[0]) for fr in frames]
        f = int(frames[0].shape[1])
        total = int(sum(sizes))
        s = self._slot(device, total, len(frames), f, slot)
        off = 0
        hp = s["host_pts"].numpy()
        jobs = []
        for fr in frames:
            arr = fr.numpy() if isinstance(fr, torch.Tensor) else fr
            if arr.dtype != np.float32:
                raise ValueError("points must be float32")
            jobs.append((off, arr))
            off += arr.shape[0]
        if total * f * 4 >= (4 << 20) and len(jobs) > 1 and _pack_workers() > 1:
            # large batches: the copies into the pinned staging buffer run on a few threads (numpy releases the GIL
            # inside the memcpy); a Waymo batch of 4 is 14.5 MB, ~0.9 ms on one thread
            list(_pack_pool().map(lambda j: hp.__setitem__(slice(j[0], j[0] + j[1].shape[0]), j[1]), jobs))
        else:
            for o, arr in jobs:
                hp[o:o + arr.shape[0]] = arr
        s["host_off"].numpy()[:] = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            if after is not None:
                torch.cuda.current_stream(device).wait_event(after)
            s["dev_pts"][:total].copy_(s["host_pts"][:total], non_blocking=True)
            s["dev_off"].copy_(s["host_off"], non_blocking=True)
        nbytes = total * f * 4 + (len(frames) + 1) * 4
        s["total"], s["mfp"] = total, max(sizes) if sizes else 0
        return s["dev_pts"][:total], s["dev_off"], s["mfp"], nbytes

    def staged(self, slot=0):
        """(points_dev, frame_offsets_dev, max_frame_points) of the batch last uploaded into `slot`."""
        s = self._slots[slot]
        return s["dev_pts"][:s["total"]], s["dev_off"], s["mfp"]

    # ------------------------------------------------------------------ device step
    def _lane(self, lane):
        """(engine, voxelizer) of `lane`; lanes beyond the first get their own state (same network, same options)."""
        while len(self._lanes) <= lane:
            e0 = self.engine
            eng = BackboneEngine(self.backbone, precision=e0.precision, materialize_pairs=e0.materialize_pairs,
                                 use_tensor_cores=e0.use_tensor_cores, sort_rows=e0.sort_rows,
                                 concurrent=e0.concurrent, cap_growth=e0.cap_growth)
            self._lanes.append((eng, BatchVoxelizer(*self._vox_args)))
        return self._lanes[lane]

    def launch_resident(self, points_dev, frame_offsets_dev, max_frame_points=None, lane=0):
        """All kernels of the step on the current stream; no host sync.  Returns a handle for finish()."""
        batch = frame_offsets_dev.numel() - 1
        engine, voxelizer = self._lane(lane)
        # the rulebooks only read coordinates: the means are finished on a stream of their own (joined by the
        # engine before the entry layer reads them)
        fs = None
        if engine.concurrent:
            fs = self._lane_streams.get((points_dev.device, "feat", lane))
            if fs is None:
                fs = self._lane_streams[(points_dev.device, "feat", lane)] = torch.cuda.Stream(device=points_dev.device)
        # the step's clearing launch (tables, scan states, -1 fills) depends on nothing: it runs on a side stream next to
        # the voxelizer instead of between it and the first rulebook
        cap = voxelizer.capacity(points_dev.device, points_dev.shape[0], batch, points_dev.shape[1])
        prefilled = engine.prefill(points_dev.device, cap, batch, table0_external=True)
        # the voxelizer also builds the convolutions' level-0 coordinate table while it assigns the voxel rows
        vox = voxelizer(points_dev, frame_offsets_dev, max_frame_points, features_stream=fs,
                        level0_table=engine.level0_table(points_dev.device, cap, batch))
        n0 = vox["voxel_offsets"][batch:batch + 1]
        arena = engine.launch(vox["voxel_features"], vox["voxel_coords"], batch, n0_dev=n0, cap0=vox["cap"],
                              features_ready=vox["features_ready"], prefilled=prefilled, table0_built=True)
        handle = dict(vox=vox, arena=arena, batch=batch)
        if self.bev:
            handle["spatial_features"] = self._height_compression(engine, arena, batch, lane)
        return handle

    def _height_compression(self, engine, arena, batch, lane):
        """HeightCompression of the stride-8 output into this lane's BEV buffer, row count read on the device."""
        from .height_compression import height_compression
        out_step = [st for st in engine.steps if st.export == "out"][0]
        feats = arena["bufs"][out_step.out_buf]
        last = len(arena["caps"]) - 1
        shape = [int(v) for v in engine.level_shapes[last]]
        key = (lane, batch, feats.shape[1], tuple(shape), feats.dtype, str(feats.device))
        buf = self._bev_bufs.get(key)
        if buf is None:
            ws_bytes = _lib.load().fv2p_height_compression_workspace_bytes(batch, _lib.i32x3(shape))
            buf = (torch.empty((batch, feats.shape[1] * shape[0], shape[1], shape[2]), dtype=feats.dtype,
                               device=feats.device),
                   torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device))
            self._bev_bufs = {k: v for k, v in self._bev_bufs.items() if k[0] != lane}
            self._bev_bufs[key] = buf
            engine.arena_gen += 1  # new buffers behind captured addresses
        return height_compression(feats, arena["indices"][last], shape, batch, out=buf[0],
                                  n_dev=arena["counts"][last:last + 1], workspace=buf[1])

    def launch_graph(self, slot=0, lane=0):
        """Same step as launch_resident over the WHOLE staging buffer of `slot`, replayed from a CUDA graph.

        The step has no host synchronisation and every buffer (points staging, voxel outputs, rulebooks, feature
        arena) has a fixed address, so it is captured once per staging capacity and replayed with one launch.
        Row counts are read from device memory by every kernel, so the same graph serves batches with different
        point counts up to the staging capacity.  Call upload() first; returns the handle for finish()."""
        s = self._slots.get(slot)
        if s is None:
            raise RuntimeError("launch_graph: call upload() first")
        key = (s["dev_pts"].data_ptr(), s["dev_off"].data_ptr(), s["pcap"], s["batch"], lane)
        engine, voxelizer = self._lane(lane)
        entry = self._graphs.get(key)
        # the parameter identity is recomputed here, not read back from the engine: a replay never reaches
        # _prepare_params, so weights updated after the capture would otherwise keep their stale packed copies
        if entry is not None and entry[2] != (engine.arena_gen, voxelizer.gen, engine.param_key(s["device"])):
            entry = None  # buffers or parameters behind the captured addresses changed: capture again
        if entry is None:
            cur = torch.cuda.current_stream(s["device"])
            side = torch.cuda.Stream(device=s["device"])
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up: allocates arenas, packs weights, sets kernel attributes
                for _ in range(2):
                    handle = self.launch_resident(s["dev_pts"], s["dev_off"], s["pcap"], lane)
            cur.wait_stream(side)
            torch.cuda.synchronize(s["device"])
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                handle = self.launch_resident(s["dev_pts"], s["dev_off"], s["pcap"], lane)
            entry = (graph, handle, (engine.arena_gen, voxelizer.gen, engine._param_key))
            self._graphs[key] = entry
        entry[0].replay()
        return entry[1]

    def finish(self, handle, fetch="counts", sync=True):
        """D2H of the row counts (always) and, with fetch='encoded', of the stride-8 output features and
        indices.  Returns (outputs dict of SparseConvTensor, info dict)."""
        vox, arena, batch = handle["vox"], handle["arena"], handle["batch"]
        eng = self.engine
        arena["counts_host"].copy_(arena["counts"], non_blocking=True)
        d2h = arena["counts_host"].numel() * 4
        stream = torch.cuda.current_stream(arena["device"])
        if fetch == "encoded":
            # sized by capacity on the host side; only the live rows are copied after the counts are known
            stream.synchronize()
            outs, n = eng.views(arena, vox["voxel_coords"], batch)
            enc = outs["out"]
            rows = enc.features.shape[0]
            hb = self._result_host(0, rows, enc.features.shape[1], enc.features.dtype)
            hb[0][:rows].copy_(enc.features, non_blocking=True)
            hb[1][:rows].copy_(enc.indices, non_blocking=True)
            stream.synchronize()
            d2h += enc.features.numel() * enc.features.element_size() + enc.indices.numel() * 4
            info = dict(counts=n, d2h_bytes=d2h, encoded_features_host=hb[0][:rows], encoded_indices_host=hb[1][:rows])
            return outs, info
        if sync:
            stream.synchronize()
        outs, n = eng.views(arena, vox["voxel_coords"], batch)
        return outs, dict(counts=n, d2h_bytes=d2h)

    def _result_host(self, slot, rows, cols, dtype):
        """Pinned host staging for the stride-8 result of `slot` (grow-only)."""
        if self._enc_host is None:
            self._enc_host = {}
        hb = self._enc_host.get(slot)
        if hb is None or hb[0].shape[0] < rows or hb[0].shape[1] != cols or hb[0].dtype != dtype:
            cap = int(rows * 1.25) + 64
            hb = (torch.empty((cap, cols), dtype=dtype).pin_memory(), torch.empty((cap, 4), dtype=torch.int32).pin_memory())
            self._enc_host[slot] = hb
        return hb

    def __call__(self, frames, device="cuda", fetch="counts"):
        pts, off, mfp, h2d = self.upload(frames, device)
        while True:
            handle = self.launch_graph() if self.use_graph else self.launch_resident(pts, off, mfp)
            try:
                outs, info = self.finish(handle, fetch)
                break
            except CapacityOverflow:  # more rows than the arena bound: enlarge every lane's bounds and run again
                if not any([eng.grow() for eng, _ in self._lanes]):
                    raise
        info["h2d_bytes"] = h2d
        batch_dict = {
            'batch_size': len(frames),
            'voxel_features': handle["vox"]["voxel_features"][:info["counts"][0]],
            'voxel_coords': handle["vox"]["voxel_coords"][:info["counts"][0]],
            'voxel_num_points': handle["vox"]["voxel_num_points"][:info["counts"][0]],
            'encoded_spconv_tensor': outs['out'], 'encoded_spconv_tensor_stride': 8,
            'multi_scale_3d_features': {k: outs[k] for k in ('x_conv1', 'x_conv2', 'x_conv3', 'x_conv4')},
            'multi_scale_3d_strides': {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8},
        }
        if "spatial_features" in handle:  # HeightCompression.forward's keys (height_compression.py:23-24)
            batch_dict['spatial_features'] = handle["spatial_features"]
            batch_dict['spatial_features_stride'] = 8
        return batch_dict, info

    # ------------------------------------------------------------------ pipelined throughput mode
    def run_stream(self, batches, device="cuda", depth=2):
        """Generator over an iterable of batches (each a list of host point arrays).  Yields, in order, one dict per
        batch with HOST results: 'counts' (rows per level), 'encoded_features' [N,C] and 'encoded_indices' [N,4]
        (pinned tensors, valid until `depth` further batches have been yielded), 'h2d_bytes', 'd2h_bytes'.

        Same work per batch as __call__(fetch='encoded'), but pipelined: `depth` batches are in flight (staging
        slots), their kernels alternate between the engine lanes, the H2D copy of batch i+1 and the D2H copy of
        batch i-depth+1 run on two copy streams meanwhile, and batch i+1 is packed into pinned memory by a helper
        thread while this one waits for a result.  Deeper pipelines (3, 4, 6 slots) measured the same throughput on
        one B200: the kernels set the pace (profiles/r2_notes.md).  The result of a step is snapshotted on the device
        (live rows only) right after the step, so the arena can be reused at once.
        """
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._copy_stream is None or self._copy_stream.device != device:
            self._copy_stream = torch.cuda.Stream(device=device)
            self._copy_in_stream = torch.cuda.Stream(device=device)
        # results leave on `copy`, points arrive on `copy_in`: PCIe is full duplex, one stream would serialise them
        copy, copy_in, main = self._copy_stream, self._copy_in_stream, torch.cuda.current_stream(device)
        lib = _lib.load()
        out_step = [st for st in self.engine.steps if st.export == "out"][0]
        n_lanes = min(self.lanes, depth)
        streams = []
        for lane in range(n_lanes):
            if lane == 0:
                streams.append(main)
                continue
            st = self._lane_streams.get((device, lane))
            if st is None:
                st = self._lane_streams[(device, lane)] = torch.cuda.Stream(device=device)
            st.wait_stream(main)
            streams.append(st)
        copy_in.wait_stream(main)
        # Staging runs one batch ahead on a helper thread: while this thread waits for the kernels and the D2H copy of
        # batch i-1, batch i+1 is packed into its slot's pinned buffer (the memcpy and the CUDA waits release the GIL)
        # and its H2D copy is enqueued behind the last kernels that read the slot's device buffer.  Without it the
        # host's serial work per batch (pack + D2H wait) was as long as the kernels and set the pace.
        slot_busy = {}  # slot -> (H2D finished: the pinned buffer is free, kernels finished: the device buffer is free)

        def stage(i, frames):
            slot = i % depth
            with torch.cuda.device(device):
                h2d_done, kernels_done = slot_busy.get(slot, (None, None))
                if h2d_done is not None:
                    h2d_done.synchronize()
                pts, off, mfp, h2d = self.upload(frames, device, slot=slot, stream=copy_in, after=kernels_done)
                up = torch.cuda.Event()
                up.record(copy_in)
            return slot, pts, off, mfp, h2d, up

        it = iter(batches)
        first = next(it, None)
        fut = _stage_pool().submit(stage, 0, first) if first is not None else None
        pending = []
        i = 0
        while fut is not None:
            slot, pts, off, mfp, h2d, up = fut.result()
            lane = i % n_lanes
            stream = streams[lane]
            stream.wait_event(up)
            with torch.cuda.stream(stream):
                handle = self.launch_graph(slot, lane) if self.use_graph else self.launch_resident(pts, off, mfp, lane)
                arena = handle["arena"]
                last = len(arena["caps"]) - 1
                src_f = arena["bufs"][out_step.out_buf]
                src_i = arena["indices"][last]
                snap = self._snapshot(slot, src_f, src_i, arena["counts"])
                n_ptr = _lib.ctypes.c_void_p(arena["counts"].data_ptr() + 4 * last)
                with torch.cuda.device(device):
                    _lib.check(lib.fv2p_copy_rows(_lib.ptr(src_f), _lib.ptr(snap["feat"]),
                                                  src_f.shape[1] * src_f.element_size(), src_f.shape[0], n_ptr,
                                                  _lib.stream_ptr(device)), "copy_rows")
                    _lib.check(lib.fv2p_copy_rows(_lib.ptr(src_i), _lib.ptr(snap["ind"]), 16, src_i.shape[0], n_ptr,
                                                  _lib.stream_ptr(device)), "copy_rows")
                    snap["counts"].copy_(arena["counts"], non_blocking=True)
                done = torch.cuda.Event()
                done.record(stream)
            slot_busy[slot] = (up, done)
            pending.append((slot, snap, done, h2d, src_f.shape[1], src_f.dtype))
            # the next batch's staging starts now (its slot was last used by batch i + 1 - depth, already launched)
            nxt = next(it, None)
            i += 1
            fut = _stage_pool().submit(stage, i, nxt) if nxt is not None else None
            if len(pending) >= depth:
                yield self._collect(pending.pop(0), copy)
        while pending:
            yield self._collect(pending.pop(0), copy)
        for st in streams[1:]:
            main.wait_stream(st)
        main.wait_stream(copy_in)

    def _snapshot(self, slot, feat, ind, counts):
        key = ("snap", slot)
        s = self._slots.get(key)
        if s is None or s["feat"].shape != feat.shape or s["feat"].dtype != feat.dtype:
            s = dict(feat=torch.empty_like(feat), ind=torch.empty_like(ind), counts=torch.empty_like(counts),
                     counts_host=torch.zeros(counts.shape, dtype=counts.dtype).pin_memory())
            self._slots[key] = s
        return s

    def _collect(self, item, copy):
        slot, snap, done, h2d, cols, dtype = item
        copy.wait_event(done)
        with torch.cuda.stream(copy):
            snap["counts_host"].copy_(snap["counts"], non_blocking=True)
        copy.synchronize()
        host = snap["counts_host"].tolist()
        n_levels = len(host) - 1
        if host[n_levels]:
            for eng, _ in self._lanes:
                eng.grow()
            raise CapacityOverflow("fv2p_b200: capacity overflow (status %d) in run_stream; the arena bounds have been "
                                   "enlarged, restart the stream from this batch (or build the backbone with "
                                   "CAP_GROWTH: None for the hard bounds)" % host[n_levels])
        rows = host[n_levels - 1]
        hb = self._result_host(("stream", slot), rows, cols, dtype)
        with torch.cuda.stream(copy):
            hb[0][:rows].copy_(snap["feat"][:rows], non_blocking=True)
            hb[1][:rows].copy_(snap["ind"][:rows], non_blocking=True)
        copy.synchronize()
        d2h = len(host) * 4 + rows * cols * hb[0].element_size() + rows * 16
        return dict(counts=host[:n_levels], encoded_features=hb[0][:rows], encoded_indices=hb[1][:rows],
                    h2d_bytes=h2d, d2h_bytes=d2h)


_POOL = None
_STAGE_POOL = None


def _stage_pool():
    """One helper thread that packs and uploads the next batch of run_stream."""
    global _STAGE_POOL
    if _STAGE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _STAGE_POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="fv2p-stage")
    return _STAGE_POOL


def _pack_workers():
    """Threads for packing frames into pinned memory: up to 4, but no more than this process's share of the host
    cores when several ranks run on one box (8 ranks x (4 pack threads + staging thread + a spinning main thread) on a
    16-core host made the end-to-end path host-bound: 7.3 k frames/s against 12.6 k device-timed on 8 GPUs)."""
    import os
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 4
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(4, cores // ranks - 1))


def _pack_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=_pack_workers(), thread_name_prefix="fv2p-pack")
    return _POOL


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

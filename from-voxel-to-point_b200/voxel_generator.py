"""VoxelGenerator with the reference's constructor / generate() contract
(pcdet/datasets/processor/voxel_generator.py:5-72), computed by the sm_100a voxelize+mean kernels, and
BatchVoxelizer, the batched device-resident form the fused hot path uses."""
import numpy as np
import torch

from . import _lib


class VoxelGenerator(object):
    """Args as the reference: voxel_size [3], point_cloud_range [6], max_num_points, max_voxels=20000."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, device=None):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = max_num_points
        self._max_voxels = max_voxels
        self._grid_size = grid_size
        self._device = device
        self.last_voxel_features = None  # mean features of the last generate() call (device tensor)

    def generate(self, points, max_voxels=None):
        """points [P,F] float32 (numpy or CUDA tensor) -> (voxels [M,T,F], coors [M,3] zyx int32, num [M] int32),
        the same types as the input (numpy in -> numpy out, like the reference)."""
        max_voxels = int(max_voxels or self._max_voxels)
        as_numpy = isinstance(points, np.ndarray)
        if as_numpy:
            if points.dtype != np.float32:
                raise ValueError("points must be float32 (fp64 would change the reference's arithmetic, SURVEY A.4)")
            device = torch.device(self._device or "cuda")
            pts = torch.from_numpy(np.ascontiguousarray(points)).to(device)
        else:
            pts = points.contiguous()
            if pts.dtype != torch.float32:
                raise ValueError("points must be float32")
        dev = _lib.require_device(pts)
        lib = _lib.load()
        p, f = pts.shape
        t = int(self._max_num_points)
        cap = max(1, min(p, max_voxels))
        device = pts.device
        coords = torch.empty((cap, 4), dtype=torch.int32, device=device)
        feats = torch.empty((cap, f), dtype=torch.float32, device=device)
        num = torch.empty((cap,), dtype=torch.int32, device=device)
        voxels = torch.empty((cap, t, f), dtype=torch.float32, device=device)
        ws_bytes = lib.fv2p_voxelize_workspace_bytes(p, 1, p, t, cap) + 1024
        ws = _lib.Workspace.get(device, ws_bytes, "voxelize")
        m_host = (_lib.ctypes.c_int32 * 1)(0)
        with torch.cuda.device(dev):
            st = lib.fv2p_voxel_generate(_lib.ptr(pts), p, f, _lib.f32arr(self._point_cloud_range),
                                         _lib.f32arr(self._voxel_size), t, max_voxels, _lib.ptr(coords),
                                         _lib.ptr(feats), _lib.ptr(num), _lib.ptr(voxels), cap, m_host, _lib.ptr(ws),
                                         ws.numel(), _lib.stream_ptr(device))
        _lib.check(st, "voxel_generate")
        m = int(m_host[0])
        self.last_voxel_features = feats[:m]
        voxels, coors, num = voxels[:m], coords[:m, 1:], num[:m]
        if as_numpy:
            return voxels.cpu().numpy(), np.ascontiguousarray(coors.cpu().numpy()), num.cpu().numpy()
        return voxels, coors.contiguous(), num

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size

    def __repr__(self):
        return ("VoxelGenerator(voxel_size=%s, point_cloud_range=%s, max_num_points=%s, max_voxels=%s, grid_size=%s)"
                % (self._voxel_size, self._point_cloud_range.tolist(), self._max_num_points, self._max_voxels,
                   self._grid_size.tolist()))


class BatchVoxelizer(object):
    """Voxelize + MeanVFE for a whole batch in one stream-ordered call, everything device-resident.

    Produces the collate_batch layout (pcdet/datasets/dataset.py:162-169): coords [M,4] = (b,z,y,x) with frames
    contiguous, voxel_features [M,F] (what MeanVFE would emit), num_points [M]; M stays in a device scalar
    (``voxel_offsets[batch]``).  Does not synchronise.
    """

    LAUNCHES = 5  # kernels per call: clear, insert, rank (+ level-0 table), select, reduce

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels, want_voxels=False):
        self.vg = VoxelGenerator(voxel_size, point_cloud_range, max_num_points, max_voxels)
        self.want_voxels = want_voxels
        self._bufs = None
        self.gen = 0  # bumped whenever the output buffers are reallocated (captured graphs hold their addresses)

    def _ensure(self, device, total_points, batch, f):
        t = int(self.vg._max_num_points)
        cap = max(1, min(total_points, batch * int(self.vg._max_voxels)))
        key = (str(device), batch, f)
        b = self._bufs
        if b is None or b["key"] != key or b["cap"] < cap or b["pcap"] < total_points:
            pcap = int(total_points * 1.1) + 1024
            cap = max(1, min(pcap, batch * int(self.vg._max_voxels)))
            lib = _lib.load()
            b = dict(key=key, cap=cap, pcap=pcap,
                     coords=torch.empty((cap, 4), dtype=torch.int32, device=device),
                     feats=torch.empty((cap, f), dtype=torch.float32, device=device),
                     num=torch.empty((cap,), dtype=torch.int32, device=device),
                     voxels=torch.empty((cap, t, f), dtype=torch.float32, device=device) if self.want_voxels else None,
                     voff=torch.zeros((batch + 1,), dtype=torch.int32, device=device),
                     status=torch.zeros((1,), dtype=torch.int32, device=device),
                     ws=torch.empty(lib.fv2p_voxelize_workspace_bytes(pcap, batch, pcap, t, cap) + 1024,
                                    dtype=torch.uint8, device=device))
            self._bufs = b
            self.gen += 1
        return b

    def capacity(self, device, total_points, batch, f):
        """Row capacity of the output buffers a call with these sizes will use (allocates them if needed)."""
        return self._ensure(device, total_points, batch, f)["cap"]

    def __call__(self, points, frame_offsets, max_frame_points=None, features_stream=None, level0_table=None):
        """points [P,F] fp32 CUDA, frame_offsets [B+1] int32 CUDA.  Returns a dict of device buffers sized at
        capacity plus 'voxel_offsets' [B+1]; live rows are [:voxel_offsets[B]].

        With ``features_stream`` the coordinates are complete on the current stream and the means / point counts on
        that stream; 'features_ready' is then the event to wait for before reading them.

        ``level0_table`` = (table buffer, its row capacity, spatial shape [D,H,W]): the call also builds the sparse
        convolutions' level-0 coordinate table (fv2p_voxelize_mean_table) - BackboneEngine.level0_table() hands it out
        and launch(table0_built=True) then skips fv2p_table_build."""
        dev = _lib.require_device(points)
        assert points.dtype == torch.float32 and points.is_contiguous()
        assert frame_offsets.dtype == torch.int32 and frame_offsets.is_cuda
        batch = frame_offsets.numel() - 1
        p, f = points.shape
        b = self._ensure(points.device, p, batch, f)
        vg = self.vg
        with torch.cuda.device(dev):
            table, table_cap, shape3 = level0_table if level0_table is not None else (None, 0, None)
            st = _lib.load().fv2p_voxelize_mean_table(
                _lib.ptr(points), _lib.ptr(frame_offsets), p, batch, int(max_frame_points or p), f,
                _lib.f32arr(vg._point_cloud_range), _lib.f32arr(vg._voxel_size), int(vg._max_num_points),
                int(vg._max_voxels), _lib.ptr(b["coords"]), _lib.ptr(b["feats"]), _lib.ptr(b["num"]),
                _lib.ptr(b["voxels"]), _lib.ptr(b["voff"]), b["cap"], _lib.ptr(b["status"]), _lib.ptr(b["ws"]),
                b["ws"].numel(), _lib.stream_ptr(points.device),
                _lib.ctypes.c_void_p(features_stream.cuda_stream) if features_stream is not None else None,
                _lib.ptr(table), int(table_cap), _lib.i32x3(shape3) if shape3 is not None else None)
        _lib.check(st, "voxelize_mean")
        ready = None
        if features_stream is not None:
            ready = torch.cuda.Event()
            ready.record(features_stream)
        return dict(voxel_coords=b["coords"], voxel_features=b["feats"], voxel_num_points=b["num"],
                    voxels=b["voxels"], voxel_offsets=b["voff"], status=b["status"], cap=b["cap"],
                    features_ready=ready)

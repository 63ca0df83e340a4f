"""Deterministic LiDAR-like synthetic frames (the reference ships no data).

A frame is a spinning 64-beam sensor over a ground plane, piecewise facades ("walls") and a few
car-sized boxes, with 2 cm range noise and 10 % dropout, cropped by the reference's xy range mask
(pcdet/utils/common_utils.py:59-62).  Surface-like clouds matter: uniformly random points have almost
no occupied neighbours and would misrepresent the gather/GEMM load of the backbone.

Dataset presets carry the config keys the path reads (tools/cfgs/dataset_configs/kitti_dataset.yaml:4,
65-70 and waymo_dataset.yaml:5,62-67).
"""
import numpy as np

DATASETS = {
    "kitti": dict(
        point_cloud_range=[0.0, -40.0, -3.0, 70.4, 40.0, 1.0], voxel_size=[0.05, 0.05, 0.1],
        max_points_per_voxel=5, max_voxels={"train": 16000, "test": 40000}, num_point_features=4,
        elev_deg=(-24.8, 2.0), fov_deg=45.0, az_steps=330, sensor_h=1.73, ground_z=-1.73, r_max=70.0),
    "waymo": dict(
        point_cloud_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15],
        max_points_per_voxel=5, max_voxels={"train": 80000, "test": 90000}, num_point_features=5,
        elev_deg=(-17.6, 2.4), fov_deg=180.0, az_steps=3150, sensor_h=2.0, ground_z=0.0, r_max=75.0),
}


def grid_size(cfg):
    """(x,y,z) grid like DataProcessor (data_processor.py:59-60); fp64 there, same integers."""
    r = np.asarray(cfg["point_cloud_range"], np.float64)
    v = np.asarray(cfg["voxel_size"], np.float64)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def _ray_boxes(dirs, origin_z, boxes):
    """Nearest hit distance of rays (from (0,0,origin_z)) with axis-aligned boxes, inf if none."""
    t_best = np.full(dirs.shape[0], np.inf)
    o = np.array([0.0, 0.0, origin_z])
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        for lo, hi in boxes:
            t0 = (lo - o) * inv
            t1 = (hi - o) * inv
            tn = np.minimum(t0, t1).max(axis=1)
            tf = np.maximum(t0, t1).min(axis=1)
            hit = (tf >= tn) & (tn > 0.5)
            t_best = np.where(hit & (tn < t_best), tn, t_best)
    return t_best


def lidar_frame(dataset="kitti", seed=0, shuffle=False, az_steps=None, beams=64):
    """Returns points [P,F] float32: x,y,z,intensity(,elongation)."""
    cfg = DATASETS[dataset]
    rng = np.random.default_rng(seed)
    az_steps = int(az_steps or cfg["az_steps"])
    fov = np.deg2rad(cfg["fov_deg"])
    az = np.linspace(-fov, fov, az_steps, endpoint=cfg["fov_deg"] < 180.0)
    elev = np.deg2rad(np.linspace(cfg["elev_deg"][0], cfg["elev_deg"][1], beams))
    A, E = np.meshgrid(az, elev, indexing="ij")  # scan order: azimuth-major like a spinning sensor
    A, E = A.ravel(), E.ravel()
    dirs = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], 1)
    sensor_z = cfg["ground_z"] + cfg["sensor_h"]
    # ground plane
    with np.errstate(divide="ignore"):
        t_ground = np.where(dirs[:, 2] < -1e-6, -cfg["sensor_h"] / dirs[:, 2], np.inf)
    # facades: piecewise-constant range over ~3 degree azimuth segments
    seg = max(1, int(round(np.deg2rad(3.0) / (2 * fov / az_steps))))
    n_seg = (az_steps + seg - 1) // seg
    wall_r = rng.uniform(8.0, cfg["r_max"], n_seg)
    wall_t = np.repeat(np.repeat(wall_r, seg)[:az_steps], beams) / np.maximum(np.cos(E), 1e-3)
    # a handful of 4 x 2 x 1.6 m boxes on the ground
    boxes = []
    for _ in range(12):
        r, a = rng.uniform(6.0, 45.0), rng.uniform(-fov, fov)
        cx, cy = r * np.cos(a), r * np.sin(a)
        lo = np.array([cx - 2.0, cy - 1.0, cfg["ground_z"]])
        boxes.append((lo, lo + np.array([4.0, 2.0, 1.6])))
    t_box = _ray_boxes(dirs, sensor_z, boxes)
    t = np.minimum(np.minimum(t_ground, wall_t), t_box)
    t = t + rng.normal(0.0, 0.02, t.shape)
    keep = np.isfinite(t) & (t > 1.0) & (rng.random(t.shape) > 0.10)
    xyz = dirs[keep] * t[keep, None]
    xyz[:, 2] += sensor_z
    F = cfg["num_point_features"]
    extra = rng.random((xyz.shape[0], F - 3))
    pts = np.concatenate([xyz, extra], 1).astype(np.float32)
    lo_hi = cfg["point_cloud_range"]
    m = (pts[:, 0] >= lo_hi[0]) & (pts[:, 0] <= lo_hi[3]) & (pts[:, 1] >= lo_hi[1]) & (pts[:, 1] <= lo_hi[4])
    pts = pts[m]
    if shuffle:  # DataProcessor.shuffle_points (data_processor.py:31-41), seeded here
        pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts)


def random_voxels(shape, n, batch, seed=0):
    """Random DISTINCT active coordinates per batch element, batch-contiguous (the recipe of the
    reference's pcdet/ops/spconv/test_utils.py:144-193 generate_sparse_data)."""
    rng = np.random.default_rng(seed)
    vol = int(np.prod(shape))
    rows = []
    for b in range(batch):
        lin = rng.choice(vol, size=min(n, vol), replace=False)
        z, rem = np.divmod(lin, shape[1] * shape[2])
        y, x = np.divmod(rem, shape[2])
        rows.append(np.stack([np.full_like(z, b), z, y, x], 1))
    return np.concatenate(rows, 0).astype(np.int32)


def randomize_state(state_dict, seed=0):
    """Deterministic, framework-independent parameter fill for a backbone ``state_dict``.

    Works on any mapping name -> tensor/ndarray with the reference's key names (conv ``weight``
    [kD,kH,kW,Cin,Cout] / ``bias``, BatchNorm ``weight bias running_mean running_var``), so the
    reference modules, this package's modules and any CPU checker can all be loaded with the SAME
    numbers without shipping multi-megabyte fixtures.  Conv init follows conv.py:106-111
    (kaiming-uniform, a=sqrt(5) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in))); BatchNorm statistics follow
    SURVEY.md section 8d (gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5)).
    Returns {name: float32 ndarray}; integer buffers (num_batches_tracked) are skipped.
    """
    import zlib

    out = {}
    for name in state_dict:
        shape = tuple(int(s) for s in state_dict[name].shape)
        if name.endswith("num_batches_tracked"):
            continue
        rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
        leaf = name.rsplit(".", 1)[-1]
        owner = name.rsplit(".", 1)[0]
        is_conv_w = leaf == "weight" and len(shape) == 5
        conv_w_name = owner + ".weight"
        if is_conv_w:
            fan_in = int(np.prod(shape[:4]))
            b = 1.0 / np.sqrt(fan_in)
            val = rng.uniform(-b, b, shape)
        elif leaf == "bias" and conv_w_name in state_dict and len(state_dict[conv_w_name].shape) == 5:
            w_shape = state_dict[conv_w_name].shape
            b = 1.0 / np.sqrt(float(np.prod([int(s) for s in w_shape[:4]])))
            val = rng.uniform(-b, b, shape)
        elif leaf == "weight":
            val = rng.uniform(0.5, 1.5, shape)
        elif leaf == "running_var":
            val = rng.uniform(0.5, 1.5, shape)
        else:  # BN bias, running_mean
            val = rng.normal(0.0, 0.1, shape)
        out[name] = val.astype(np.float32)
    return out

"""Generates the committed golden fixtures by running the REAL reference in the build container.

    python tests/golden/make_golden.py

Needs /root/reference (imported in place, see oracle/ref.py) and oracle/_ref/sparse_conv_ext.so
(python oracle/build_ref.py).  Neither exists on the GPU box: tests only read the .npz files written
here.  Every array stored is an INPUT drawn from a seeded numpy generator or an OUTPUT of reference
code: numba VoxelGenerator (voxel_generator.py), MeanVFE (mean_vfe.py), sparse_conv_ext CPU path
(get_indice_pairs_3d / indice_conv_fp32) and the reference backbones (spconv_backbone.py).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

spec = importlib.util.spec_from_file_location("fv2p_synth", os.path.join(ROOT, "from-voxel-to-point_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

R = ref.load_reference_python()
torch.manual_seed(0)
torch.set_num_threads(4)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", name, len(arrays), "arrays, %.1f KB" % (os.path.getsize(path) / 1024))


# ---------------------------------------------------------------------------------- voxelizer
def voxel_cases():
    cases = []
    for ds, az, seed, shuffle, max_vox in (("kitti", 48, 1, False, 40000), ("kitti", 48, 2, True, 1500),
                                           ("waymo", 120, 3, False, 90000), ("waymo", 120, 4, True, 2500)):
        cfg = synth.DATASETS[ds]
        pts = synth.lidar_frame(ds, seed, shuffle=shuffle, az_steps=az)
        cases.append((ds, pts, cfg["voxel_size"], cfg["point_cloud_range"], 5, max_vox))
    # boundary stress: points sitting on / next to voxel faces, out of range in z, duplicates
    rng = np.random.default_rng(5)
    cfg = synth.DATASETS["kitti"]
    vs, r = np.float32(cfg["voxel_size"]), np.float32(cfg["point_cloud_range"])
    ijk = rng.integers(0, [200, 200, 40], size=(4000, 3)).astype(np.float32)
    eps = rng.choice(np.float32([0.0, 1e-7, -1e-7, 1e-4, -1e-4, 0.5]), size=(4000, 3))
    xyz = r[:3] + (ijk + eps) * vs
    xyz[:50, 2] = 7.0
    xyz[50:100, 0] = -1.0
    pts = np.concatenate([xyz, rng.random((4000, 1), dtype=np.float32)], 1).astype(np.float32)
    pts = np.concatenate([pts, pts[:500]], 0)
    cases.append(("edge", pts, cfg["voxel_size"], cfg["point_cloud_range"], 3, 3000))
    return cases


arrs = {}
for i, (tag, pts, vs, rng_, T, mv) in enumerate(voxel_cases()):
    gen = R.VoxelGenerator(voxel_size=vs, point_cloud_range=rng_, max_num_points=T, max_voxels=mv)
    voxels, coors, num = gen.generate(pts)
    vfe = R.MeanVFE(model_cfg={}, num_point_features=pts.shape[1])
    bd = vfe({"voxels": torch.from_numpy(voxels), "voxel_num_points": torch.from_numpy(num).float()})
    arrs.update({f"c{i}_points": pts, f"c{i}_voxel_size": np.float32(vs), f"c{i}_range": np.float32(rng_),
                 f"c{i}_T": np.int32(T), f"c{i}_max_voxels": np.int32(mv), f"c{i}_voxels": voxels,
                 f"c{i}_coors": coors, f"c{i}_num": num, f"c{i}_mean": bd["voxel_features"].numpy()})
    print(" voxel case", tag, pts.shape, "->", voxels.shape, "grid", gen.grid_size)
arrs["n_cases"] = np.int32(i + 1)
save("voxelize", **arrs)

# ---------------------------------------------------------------------------------- rulebooks + conv
GEOMS = [  # (tag, subm, ksize, stride, pad)
    ("subm3", True, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ("s2p1", False, [3, 3, 3], [2, 2, 2], [1, 1, 1]),
    ("s2p011", False, [3, 3, 3], [2, 2, 2], [0, 1, 1]),
    ("down311", False, [3, 1, 1], [2, 1, 1], [0, 0, 0]),
    ("s1p1", False, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ("k2s2", False, [2, 2, 2], [2, 2, 2], [0, 0, 0]),
]
arrs = {}
rng = np.random.default_rng(7)
shape = [11, 40, 36]
coords = synth.random_voxels(shape, 700, 3, seed=11)
# partly shuffled, non batch-contiguous input (the reference accepts any row order)
perm = np.arange(coords.shape[0])
sel = rng.choice(coords.shape[0], 600, replace=False)
perm[np.sort(sel)] = sel
coords_shuf = coords[perm]
# dense blob: many neighbours
zz, yy, xx = np.meshgrid(np.arange(2, 8), np.arange(5, 17), np.arange(3, 14), indexing="ij")
blob = np.stack([np.zeros(zz.size, np.int64), zz.ravel(), yy.ravel(), xx.ravel()], 1)
blob = blob[rng.random(blob.shape[0]) < 0.7].astype(np.int32)
for cname, cset, batch in (("rand", coords, 3), ("shuf", coords_shuf, 3), ("blob", blob, 1)):
    arrs[f"{cname}_indices"] = cset
    arrs[f"{cname}_batch"] = np.int32(batch)
    for tag, subm, ks, st, pd in GEOMS:
        t_ind = torch.from_numpy(cset)
        outids, pairs, num = R.spconv.ops.get_indice_pairs(t_ind, batch, shape, ks, st, pd, 1, 0, subm, False)
        arrs[f"{cname}_{tag}_outids"] = outids.numpy()
        arrs[f"{cname}_{tag}_pairs"] = pairs.numpy()
        arrs[f"{cname}_{tag}_num"] = num.numpy()
        if cname == "blob":
            for cin, cout in ((4, 16), (16, 32), (5, 16)):
                feats = rng.standard_normal((cset.shape[0], cin)).astype(np.float32)
                w = (rng.standard_normal((*ks, cin, cout)) / np.sqrt(cin * np.prod(ks))).astype(np.float32)
                out = R.ext.indice_conv_fp32(torch.from_numpy(feats), torch.from_numpy(w), pairs, num,
                                             outids.shape[0], 0, int(subm))
                arrs[f"conv_{tag}_{cin}_{cout}_feats"] = feats
                arrs[f"conv_{tag}_{cin}_{cout}_w"] = w
                arrs[f"conv_{tag}_{cin}_{cout}_out"] = out.numpy()
arrs["shape"] = np.int32(shape)
arrs["geoms"] = np.array([g[0] for g in GEOMS])
save("rulebook_conv", **arrs)

# ---------------------------------------------------------------------------------- backbones
for ds, az, nframes in (("kitti", 12, 2), ("waymo", 20, 1)):
    cfg = synth.DATASETS[ds]
    g = synth.grid_size(cfg)
    feats_l, coords_l, pts_l = [], [], []
    for b in range(nframes):
        pts = synth.lidar_frame(ds, 100 + b, az_steps=az)
        gen = R.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], 5, cfg["max_voxels"]["test"])
        voxels, coors, num = gen.generate(pts)
        mean = R.MeanVFE({}, pts.shape[1])({"voxels": torch.from_numpy(voxels),
                                            "voxel_num_points": torch.from_numpy(num).float()})["voxel_features"]
        feats_l.append(mean.numpy())
        coords_l.append(np.concatenate([np.full((coors.shape[0], 1), b, np.int32), coors], 1))
        pts_l.append(pts)
    feats = np.concatenate(feats_l, 0)
    coords = np.concatenate(coords_l, 0)
    for name in ("VoxelBackBone8x", "VoxelResBackBone8x"):
        net = getattr(R, name)(model_cfg={}, input_channels=feats.shape[1], grid_size=np.array(g)).eval()
        state = synth.randomize_state(net.state_dict(), seed=3)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=False)
        with torch.no_grad():
            bd = net({"voxel_features": torch.from_numpy(feats), "voxel_coords": torch.from_numpy(coords).float(),
                      "batch_size": nframes})
        arrs = {"points_offsets": np.cumsum([0] + [p.shape[0] for p in pts_l]).astype(np.int64),
                "points": np.concatenate(pts_l, 0), "voxel_features": feats, "voxel_coords": coords,
                "batch_size": np.int32(nframes), "grid_size": np.int64(g), "seed": np.int32(3)}
        for k, t in bd["multi_scale_3d_features"].items():
            arrs[k + "_features"] = t.features.numpy()
            arrs[k + "_indices"] = t.indices.numpy()
        out = bd["encoded_spconv_tensor"]
        arrs["out_features"] = out.features.numpy()
        arrs["out_indices"] = out.indices.numpy()
        arrs["out_shape"] = np.int32(out.spatial_shape)
        for key, (outids, _ind, pairs, num, _shp) in out.indice_dict.items():
            arrs[f"rb_{key}_num"] = num.numpy()
            arrs[f"rb_{key}_nout"] = np.int32(outids.shape[0])
        save(f"backbone_{ds}_{name}", **arrs)

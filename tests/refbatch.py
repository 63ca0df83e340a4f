"""Expected results for a whole batch, assembled from per-frame runs of the reference (test infrastructure).

The reference's dense int32 grid limits one call to 23 frames and costs 370 MB of memset per frame and rulebook
(SURVEY.md section 0), so the expectation for a batch is built frame by frame - frames are independent
(geometry.h:179-180) and `collate_batch` keeps them contiguous (dataset.py:162-169), hence:

  * rows of every level are the per-frame rows concatenated, batch index rewritten;
  * a submanifold / strided rulebook of the batch is, per kernel offset, the per-frame pair lists concatenated
    with the frames' input / output row offsets added, then the -1 tail (both reference loops walk the input rows
    in order, so frame b's pairs and first-touch outputs all precede frame b+1's).

`backend='ext'` drives the REAL reference extension (oracle/_ref, built from /root/reference; it travels to the GPU
box) through oracle/ref.py; `backend='oracle'` uses the C restatement (slower: ~3 s per KITTI frame, ~24 s per
Waymo frame).  'auto' prefers the extension.
"""
import numpy as np

from fv2p_b200 import synth
from oracle import oracle as O
from oracle import ref as R

EXPORTS = ("x_conv1", "x_conv2", "x_conv3", "x_conv4", "out")


def voxelize_frames(ds, frames, split="test", max_voxels=None):
    cfg = synth.DATASETS[ds]
    mv = int(max_voxels or cfg["max_voxels"][split])
    per = []
    for f in frames:
        v, c, n = O.voxelize(f, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points_per_voxel"], mv)
        per.append((O.mean_vfe(v, n), c, n))
    return per


def _frame_forward(name, state, feats, coords, shape, backend):
    if backend == "ext":
        import torch
        params = {k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}
        r = R.ext_backbone_forward(name, params, feats, coords, 1, shape)
        out = {k: (r[k][0].numpy(), r[k][1].numpy(), r[k][2]) for k in EXPORTS}
        out["rulebooks"] = {k: tuple(t.numpy() for t in v) for k, v in r["rulebooks"].items()}
        return out
    return O.backbone_forward(name, state, feats, coords, 1, shape)


def expected_batch(ds, name, state, frames, split="test", backend="auto", max_voxels=None):
    """Returns dict: 'voxel_features' [M,F], 'voxel_coords' [M,4], 'voxel_num_points' [M], per export
    (features, indices), and 'rulebooks' {key: (outids, pairs [K,2,N_in], num [K])} for the whole batch."""
    if backend == "auto":
        backend = "ext" if R.have_ext() else "oracle"
    cfg = synth.DATASETS[ds]
    gs = synth.grid_size(cfg)
    shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
    per = voxelize_frames(ds, frames, split, max_voxels)
    res = [_frame_forward(name, state, f, O.collate([c]), shape, backend) if c.shape[0] else None for f, c, n in per]
    out = {"voxel_features": np.concatenate([p[0] for p in per]),
           "voxel_coords": np.concatenate([np.concatenate([np.full((p[1].shape[0], 1), b, np.int32), p[1]], 1)
                                           for b, p in enumerate(per)]),
           "voxel_num_points": np.concatenate([p[2] for p in per])}
    live = [(b, r) for b, r in enumerate(res) if r is not None]
    for k in EXPORTS:
        feats = [r[k][0] for _, r in live]
        inds = []
        for b, r in live:
            i = np.array(r[k][1], np.int32, copy=True)
            i[:, 0] = b
            inds.append(i)
        out[k] = (np.concatenate(feats), np.concatenate(inds))
    books = {}
    keys = list(live[0][1]["rulebooks"].keys()) if live else []
    for key in keys:
        kvol = live[0][1]["rulebooks"][key][1].shape[0]
        n_in = sum(r["rulebooks"][key][1].shape[2] for _, r in live)
        pairs = np.full((kvol, 2, n_in), -1, np.int32)
        num = np.zeros((kvol,), np.int32)
        outids = []
        in_off = out_off = 0
        for b, r in live:
            o, p, n = r["rulebooks"][key]
            o = np.array(o, np.int32, copy=True)
            o[:, 0] = b
            outids.append(o)
            for k in range(kvol):
                h = int(n[k])
                pairs[k, 0, num[k]:num[k] + h] = p[k, 0, :h] + in_off
                pairs[k, 1, num[k]:num[k] + h] = p[k, 1, :h] + out_off
                num[k] += h
            in_off += p.shape[2]
            out_off += o.shape[0]
        books[key] = (np.concatenate(outids), pairs, num)
    out["rulebooks"] = books
    out["backend"] = backend
    return out

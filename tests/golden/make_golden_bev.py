"""Generates tests/golden/height_compression.npz with the REFERENCE's own HeightCompression module and
SparseConvTensor (imported in place from /root/reference; only possible in the build container).
    python tests/golden/make_golden_bev.py
Kept apart from make_golden.py so that the large backbone fixtures need not be regenerated for it.
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
hc_path = os.path.join(ref.REF_ROOT, "pcdet", "models", "backbones_2d", "map_to_bev", "height_compression.py")
spec_ = importlib.util.spec_from_file_location("ref_height_compression", hc_path)
mod = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(mod)

# small shape: the KITTI BEV map is 36 MB per frame
rng = np.random.default_rng(11)
shape, batch, c = [2, 25, 22], 3, 16
ind = synth.random_voxels(shape, 400, batch, seed=12).astype(np.int32)
feats = rng.standard_normal((ind.shape[0], c)).astype(np.float32)
x = R.spconv.SparseConvTensor(torch.from_numpy(feats), torch.from_numpy(ind), shape, batch)


class Cfg:
    NUM_BEV_FEATURES = c * shape[0]


bd = mod.HeightCompression(Cfg())({"encoded_spconv_tensor": x, "encoded_spconv_tensor_stride": 8})
path = os.path.join(HERE, "height_compression.npz")
np.savez_compressed(path, features=feats, indices=ind, spatial_shape=np.int32(shape), batch_size=np.int32(batch),
                    spatial_features=bd["spatial_features"].numpy(), stride=np.int32(bd["spatial_features_stride"]))
print("wrote", path, bd["spatial_features"].shape, "%.1f KB" % (os.path.getsize(path) / 1024))

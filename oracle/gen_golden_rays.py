"""Golden vectors for row f3 (dataset rays): calls the reference's own
libs/datasets/data_utils.get_rays / get_near_far (numpy) on the cameras of
two seeded synthetic scenes and stores inputs + outputs in
tests/golden/dataset_rays.npz.  Build container only (needs /root/reference).

cv2 / trimesh are absent here and not used by these two functions: empty
stand-in modules let data_utils import; `np.int` (removed in numpy >= 1.24,
used at data_utils.py:121-124) is aliased to int for the call."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
REF = os.environ.get("GPNERF_REFERENCE", "/root/reference")

for name in ("cv2", "trimesh"):
    sys.modules.setdefault(name, types.ModuleType(name))
if not hasattr(np, "int"):
    np.int = int          # noqa: NPY001 – what the reference was written against
sys.path.insert(0, os.path.join(REF, "libs", "datasets"))
import data_utils as du  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
import gpnerf_oracle as orc  # noqa: E402

out = {}
stats = {}
for tag, (H, seed, angle) in {"a": (48, 3, 45.0), "b": (64, 11, 200.0)}.items():
    sc = synth.retarget(synth.make_scene("zju", H=H, W=H, V=3, seed=seed), angle)
    K = sc["target_K"][0].numpy().astype(np.float64)
    pose = sc["target_pose"][0].numpy().astype(np.float64)
    R, T = pose[:, :3].copy(), pose[:, 3].copy()      # sample_ray is called with T[..., 0] (ZjumocapDataset.py:410)
    bounds = sc["can_bounds"][0].numpy()               # fp32 world-frame box, as the dataset passes it
    ray_o, ray_d = du.get_rays(H, H, K, R, T)          # reference
    ray_o = ray_o.reshape(-1, 3).astype(np.float32)    # sample_ray, test split (data_utils.py:333-334)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    near, far, at_box = du.get_near_far(bounds, ray_o, ray_d)
    near, far = near.astype(np.float32), far.astype(np.float32)
    ray_o, ray_d = ray_o[at_box], ray_d[at_box]
    o2, d2, n2, f2, m2 = orc.dataset_rays(H, H, K, R, T, bounds)
    stats[tag] = dict(rays=int(at_box.sum()), mask_ne=int((m2 != at_box).sum()), ray_o_ne=int((o2 != ray_o).sum()),
                      ray_d_ne=int((d2 != ray_d).sum()), near_ne=int((n2 != near).sum()), far_ne=int((f2 != far).sum()))
    for k, v in dict(H=np.int32(H), K=K, R=R, T=T, bounds=bounds, ray_o=ray_o, ray_d=ray_d, near=near, far=far,
                     mask_at_box=at_box).items():
        out[f"{tag}.{k}"] = v
print(stats)
assert all(v == 0 for s in stats.values() for k, v in s.items() if k.endswith("_ne")), "oracle restatement differs"
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataset_rays.npz"), **out)
print("wrote tests/golden/dataset_rays.npz")

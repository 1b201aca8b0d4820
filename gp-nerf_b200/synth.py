"""Seeded synthetic ZJU-Mocap-shaped inputs for the GP-NeRF render hot path.

Everything here is host-side numpy/torch-CPU and deterministic in `seed`.  The
dict returned by :func:`make_scene` follows the reference's batch schema
(libs/datasets/ZjumocapDataset.py:464-517) and adds the two upstream products
the hot path consumes but this repo does not build:

* ``levels``   – the 4 dense geometry-volume levels, NCDHW fp32, exactly what
  ``SparseConvTensor.dense()`` hands to ``grid_sample`` in
  libs/nerfheads/networks/SparseConvNet.py:105-124;
* ``featmaps`` – the image-encoder output ``[V, C, H/4, W/4]``
  (libs/encoders/UNet.py:217-234).

Two scene kinds:

``"zju"``    realistic sparsity: 6890 surface points on a capsule body, level-k
             active set = vertex voxels max-pooled k times (the site growth of
             ``SparseConv3d(3, 2, padding=1)``, SparseConvNet.py:78-87).
``"dense"``  worst case: every voxel active, zoomed camera so the SMPL box
             fills the frame (BaseRender semantics, no compaction).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

N_SMPL = 6890
VOXEL = 0.005


def _rodrigues(axis, angle):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle) * kx + (1 - math.cos(angle)) * (kx @ kx)


def _capsule_points(rng, a, b, r, n):
    """n points on the surface of the capsule with axis a→b and radius r."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    ax = b - a
    length = np.linalg.norm(ax)
    if length < 1e-9:
        u = rng.normal(size=(n, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        return a + r * u
    ax /= length
    # orthonormal frame around the axis
    tmp = np.array([1.0, 0, 0]) if abs(ax[0]) < 0.9 else np.array([0, 1.0, 0])
    e1 = np.cross(ax, tmp)
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(ax, e1)
    side_area = 2 * math.pi * r * length
    cap_area = 4 * math.pi * r * r
    n_side = int(round(n * side_area / (side_area + cap_area)))
    t = rng.uniform(0, length, size=n_side)
    phi = rng.uniform(0, 2 * math.pi, size=n_side)
    side = a + t[:, None] * ax + r * (np.cos(phi)[:, None] * e1 + np.sin(phi)[:, None] * e2)
    u = rng.normal(size=(n - n_side, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    along = u @ ax
    centre = np.where(along[:, None] > 0, b, a)
    caps = centre + r * u
    return np.concatenate([side, caps], 0)


def smpl_standin(rng, n=N_SMPL):
    """6890 surface points of a capsule body in the SMPL frame (x right, y up)."""
    parts = [  # (a, b, radius)
        ((0, 0.02, 0), (0, 0.50, 0), 0.145),        # torso
        ((0, 0.70, 0), (0, 0.70, 0), 0.105),        # head
        ((0.16, 0.48, 0), (0.385, 0.48, 0), 0.045),  # left arm
        ((-0.16, 0.48, 0), (-0.385, 0.48, 0), 0.045),
        ((0.09, -0.08, 0), (0.09, -0.815, 0), 0.07),  # legs
        ((-0.09, -0.08, 0), (-0.09, -0.815, 0), 0.07),
    ]
    areas = []
    for a, b, r in parts:
        length = np.linalg.norm(np.asarray(b, float) - np.asarray(a, float))
        areas.append(2 * math.pi * r * length + 4 * math.pi * r * r)
    areas = np.asarray(areas)
    counts = np.floor(n * areas / areas.sum()).astype(int)
    counts[0] += n - counts.sum()
    pts = [_capsule_points(rng, a, b, r, c) for (a, b, r), c in zip(parts, counts)]
    return np.concatenate(pts, 0).astype(np.float32)


def _look_at(pos, centre, up):
    z = centre - pos
    z = z / np.linalg.norm(z)
    down = -up
    x = np.cross(down, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    rot = np.stack([x, y, z], 0)          # world -> camera rows
    t = -rot @ pos
    return np.concatenate([rot, t[:, None]], 1).astype(np.float32)  # [3,4]


def _ring_camera(centre, up, radius, angle_deg):
    tmp = np.array([1.0, 0, 0]) if abs(up[0]) < 0.9 else np.array([0, 0, 1.0])
    e1 = np.cross(up, tmp)
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(up, e1)
    a = math.radians(angle_deg)
    pos = centre + radius * (math.cos(a) * e1 + math.sin(a) * e2)
    return _look_at(pos, centre, up)


def dataset_rays(H, W, K, pose, bounds_world):
    """Rays + near/far for every pixel, semantics of the reference's CPU loader
    (libs/datasets/data_utils.py:47-63 get_rays, :96-130 get_near_far)."""
    rot = pose[:, :3].astype(np.float64)
    t = pose[:, 3].astype(np.float64)
    r_inv = np.linalg.inv(rot)
    origin = -r_inv @ t
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    xy1 = np.stack([i, j, np.ones_like(i)], 2)
    cam = xy1 @ np.linalg.inv(K.astype(np.float64)).T
    world = cam @ r_inv.T + origin
    d = (world - origin).reshape(-1, 3).astype(np.float32)
    o = np.broadcast_to(origin.astype(np.float32), d.shape).copy()
    b = bounds_world.astype(np.float64) + np.array([-0.01, 0.01])[:, None]
    d = d.copy()
    d[np.abs(d) < 1e-5] = 1e-5
    tt = ((b[None] - o[:, None]) / d[:, None]).reshape(-1, 6)
    p = tt[..., None] * d[:, None] + o[:, None]
    lo, hi = b[0] - 1e-6, b[1] + 1e-6
    inside = np.all((p >= lo) & (p <= hi), -1)
    at_box = inside.sum(-1) == 2
    hits = p[at_box][inside[at_box]].reshape(-1, 2, 3)
    oo, dd = o[at_box], d[at_box]
    nrm = np.linalg.norm(dd, axis=1)
    sign = np.where(((hits[:, 0] - oo) * dd).sum(1) < 0, -1.0, 1.0)
    d0 = np.linalg.norm(hits[:, 0] - oo, axis=1) / nrm * sign
    d1 = np.linalg.norm(hits[:, 1] - oo, axis=1) / nrm * sign
    near = np.minimum(d0, d1).astype(np.float32)
    far = np.maximum(d0, d1).astype(np.float32)
    return oo.astype(np.float32), dd.astype(np.float32), near, far, at_box


def make_scene(kind="zju", H=512, W=512, V=3, C=32, seed=42, smooth_images=True,
               with_rays=False, level_scale=1.0):
    """Build one seeded frame.  Returns a dict of CPU tensors (batch dim 1 where
    the reference's DataLoader adds one)."""
    assert kind in ("zju", "dense")
    rng = np.random.RandomState(seed)
    gen = torch.Generator().manual_seed(seed)

    xyz_s = smpl_standin(rng)                                   # SMPL frame
    rot = _rodrigues((0.2, 1.0, 0.1), 0.3).astype(np.float32)   # batch['R'] == batch['Rh']
    th = np.array([[0.1, 0.2, 1.0]], np.float32)
    xyz_w = (xyz_s @ rot.T + th).astype(np.float32)             # p·Rᵀ + Th  (BaseRender.py:131)

    def _bounds(p):
        lo, hi = p.min(0).copy(), p.max(0).copy()
        lo[2] -= 0.05
        hi[2] += 0.05
        return np.stack([lo, hi], 0).astype(np.float32)

    can_bounds = _bounds(xyz_w)
    bounds = _bounds(xyz_s)
    # voxel coordinates and padded shape (ZjumocapDataset.py:245-254)
    dhw = xyz_s[:, [2, 1, 0]]
    min_dhw = bounds[0, [2, 1, 0]]
    max_dhw = bounds[1, [2, 1, 0]]
    coord = np.round((dhw - min_dhw) / VOXEL).astype(np.int32)
    out_sh = np.ceil((max_dhw - min_dhw) / VOXEL).astype(np.int32)
    out_sh = (out_sh | 31) + 1
    D, Hh, Ww = [int(v) for v in out_sh]

    # ---- geometry volume: 4 dense levels, NCDHW fp32 ----
    occ = torch.zeros(1, 1, D, Hh, Ww)
    cz = np.clip(coord[:, 0], 0, D - 1)
    cy = np.clip(coord[:, 1], 0, Hh - 1)
    cx = np.clip(coord[:, 2], 0, Ww - 1)
    occ[0, 0, cz, cy, cx] = 1.0
    levels = []
    for _k in range(4):
        occ = F.max_pool3d(occ, kernel_size=3, stride=2, padding=1)
        shape = occ.shape[-3:]
        vals = torch.randn((1, C) + tuple(shape), generator=gen).clamp_(min=0) * level_scale
        if kind == "zju":
            vals = vals * occ
        levels.append(vals.contiguous())

    # ---- cameras ----
    centre = xyz_w.mean(0).astype(np.float64)
    up = rot.astype(np.float64)[:, 1]
    focal = 537.0 * (W / 512.0)
    tgt_focal = focal if kind == "zju" else focal * 4.0
    K_src = np.array([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1]], np.float32)
    K_tgt = np.array([[tgt_focal, 0, W / 2.0], [0, tgt_focal, H / 2.0], [0, 0, 1]], np.float32)
    src_poses = np.stack([_ring_camera(centre, up, 3.0, 360.0 * v / V) for v in range(V)], 0)
    tgt_pose = _ring_camera(centre, up, 3.0, 45.0)

    # ---- images and encoder feature maps ----
    if smooth_images:
        small = torch.rand((V, 3, H // 8, W // 8), generator=gen) * 2 - 1
        src_imgs = F.interpolate(small, size=(H, W), mode="bilinear", align_corners=True)
    else:
        src_imgs = torch.rand((V, 3, H, W), generator=gen) * 2 - 1
    featmaps = torch.randn((V, C, H // 4, W // 4), generator=gen)

    scene = {
        "kind": kind, "H": H, "W": W, "V": V, "C": C, "seed": seed,
        "feature": torch.from_numpy(np.concatenate([xyz_s, np.zeros_like(xyz_s)], 1))[None],
        "coord": torch.from_numpy(coord)[None],
        "out_sh": torch.from_numpy(out_sh.astype(np.int32))[None],
        "bounds": torch.from_numpy(bounds)[None],
        "can_bounds": torch.from_numpy(can_bounds)[None],
        "R": torch.from_numpy(rot)[None],
        "Rh": torch.from_numpy(rot)[None],
        "Th": torch.from_numpy(th)[None],
        "src_imgs": src_imgs[None].contiguous(),
        "src_poses": torch.from_numpy(src_poses)[None],
        "src_Ks": torch.from_numpy(np.stack([K_src] * V, 0))[None],
        "target_pose": torch.from_numpy(tgt_pose)[None],
        "target_K": torch.from_numpy(K_tgt)[None],
        "target_K_inv": torch.from_numpy(np.linalg.inv(K_tgt.astype(np.float64)).astype(np.float32))[None],
        "levels": levels,
        "featmaps": featmaps.contiguous(),
        "frame_index": 0, "cam_ind": 0,
        "_ring": (centre, up),
    }
    if with_rays or kind == "dense":
        o, d, near, far, at_box = dataset_rays(H, W, K_tgt, tgt_pose, can_bounds)
        scene.update({
            "ray_o": torch.from_numpy(o)[None], "ray_d": torch.from_numpy(d)[None],
            "near": torch.from_numpy(near)[None], "far": torch.from_numpy(far)[None],
            "mask_at_box": torch.from_numpy(at_box)[None],
            "body_msk": torch.ones(1, int(at_box.sum())),
            "rgb": torch.zeros(1, int(at_box.sum()), 3),
        })
    return scene


def sparsify_levels(levels):
    """The dense levels [1,32,D,H,W] as the sparse-conv network holds them before
    `.dense()` (SparseConvNet.py:110): per level (features [N,32] fp32, indices
    [N,4] int32 = (batch, d, h, w)) of the sites with any non-zero channel, in
    row-major site order, plus the 4 (D,H,W)."""
    out, dims = [], []
    for t in levels:
        vol = t[0]                                   # [32,D,H,W]
        active = (vol != 0).any(0)
        idx = active.nonzero()                       # [N,3] (d,h,w)
        feat = vol[:, idx[:, 0], idx[:, 1], idx[:, 2]].t().contiguous()
        idx4 = torch.cat([torch.zeros_like(idx[:, :1]), idx], 1).to(torch.int32).contiguous()
        out.append((feat, idx4))
        dims.append(tuple(int(v) for v in vol.shape[-3:]))
    return out, dims


def flip_cameras(scene):
    """The same scene in the camera convention the reference's `neg_ray` switch exists for (THuman,
    BaseRender.py:165-168, 319-323; demo_render.py:236-237): every camera looks down its -z axis, i.e. visible
    points have negative depth and sit at negative ray parameters.  Built by negating the third row of every
    pose [R|t] (x_cam' = diag(1,1,-1) x_cam); intrinsics unchanged."""
    out = dict(scene)
    flip = torch.tensor([1.0, 1.0, -1.0]).view(3, 1)
    out["src_poses"] = scene["src_poses"] * flip
    out["target_pose"] = scene["target_pose"] * flip
    return out


def retarget(scene, angle_deg):
    """The same scene seen from another novel view on the camera ring (a sweep:
    only target_pose changes; the volume, the source views and K stay)."""
    centre, up = scene["_ring"]
    out = dict(scene)
    out["target_pose"] = torch.from_numpy(_ring_camera(centre, up, 3.0, float(angle_deg)))[None]
    return out


def make_head_weights(V=3, C=32, seed=42, random_bias=False):
    """Random-init hot-path head weights keyed like the reference state_dict
    (kaiming-normal weights, zero bias: libs/nerfheads/trainhead.py:13-17;
    layer shapes :39-41, :85-110).  `rgb_fc.0` takes 32·V inputs (the reference
    bakes V=3, trainhead.py:96)."""
    gen = torch.Generator().manual_seed(seed)
    cf = C + 3
    shapes = {
        "sigmahead.out_geometry_fc.0": (64, 4 * 32),
        "rgbhead.base_fc.0": (64, 3 * cf), "rgbhead.base_fc.2": (32, 64),
        "rgbhead.vis_fc.0": (32, 32), "rgbhead.vis_fc.2": (32, 32),
        "rgbhead.rgb_fc.0": (32, 32 * V), "rgbhead.rgb_fc.2": (16, 32), "rgbhead.rgb_fc.4": (3, 16),
        "rgbhead.out_geometry_fc.0": (64, 64 + 2 * cf), "rgbhead.out_geometry_fc.2": (32, 64),
        "rgbhead.out_geometry_fc.4": (16, 32), "rgbhead.out_geometry_fc.6": (1, 16),
    }
    w = {}
    for name, (o, i) in shapes.items():
        w[name + ".weight"] = torch.randn((o, i), generator=gen) * math.sqrt(2.0 / i)
        w[name + ".bias"] = (torch.randn((o,), generator=gen) * 0.1) if random_bias else torch.zeros(o)
    return w


def fill_encoder_params(module, seed=42):
    """Seeded parameters for an image-encoder module (reference ResUNet or its mirror): walks the state_dict
    in order, so two modules end up with identical values only if their keys and shapes agree.  Convolutions
    He-normal, norm scales 1 + 0.1·N(0,1), every bias 0.1·N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for name, v in sd.items():
        r = torch.randn(v.shape, generator=g)
        if v.dim() == 4:
            v.copy_(r * (2.0 / (v.shape[1] * v.shape[2] * v.shape[3])) ** 0.5)
        elif name.endswith("weight"):
            v.copy_(1.0 + 0.1 * r)
        else:
            v.copy_(0.1 * r)
    module.load_state_dict(sd)
    return module
